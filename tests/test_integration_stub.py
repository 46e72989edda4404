"""The reference-side ctypes stub printed in INTEGRATION.md section 2 is executable documentation: this file extracts
that code block, executes it against the in-tree library and checks it against include/aspire_b200.h (CPU: argument
counts of every prototype it binds; GPU: the stub's ``ot_distance_b200`` against the package's own binding)."""
import ctypes
import os
import re
import types

import pytest

from conftest import ROOT

LIB = os.path.join(ROOT, "aspire_b200", "libaspire_b200.so")


def _stub_source():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    sec = text[text.index("## 2. The ctypes stub"):]
    m = re.search(r"```python\n(.*?)```", sec, re.S)
    assert m, "INTEGRATION.md section 2 lost its python block"
    return m.group(1)


def _exec_stub():
    """exec the stub with ``ctypes.CDLL("libaspire_b200.so")`` resolved to the in-tree build."""
    shim = types.ModuleType("ctypes")
    shim.__dict__.update(ctypes.__dict__)
    shim.CDLL = lambda name, *a, **k: ctypes.CDLL(LIB if os.path.basename(name) == "libaspire_b200.so" else name, *a, **k)
    ns = {"__name__": "pair_distances_b200"}
    import builtins
    real_import = builtins.__import__

    def fake_import(name, *a, **k):
        return shim if name == "ctypes" else real_import(name, *a, **k)

    ns["__builtins__"] = dict(vars(builtins), __import__=fake_import)
    exec(compile(_stub_source(), "INTEGRATION.md#stub", "exec"), ns)
    return ns


def _header_arg_count(fn):
    hdr = open(os.path.join(ROOT, "include", "aspire_b200.h")).read()
    m = re.search(r"\b(?:int|size_t|long long)\s+" + fn + r"\s*\(([^;]*?)\)\s*;", hdr, re.S)
    assert m, fn + " not declared in include/aspire_b200.h"
    args = m.group(1).strip()
    return 0 if args in ("", "void") else args.count(",") + 1


def test_stub_argtypes_match_header():
    ns = _exec_stub()
    L = ns["_L"]
    bound = [n for n in ("asp_ot_score", "asp_ot_score_workspace_bytes") if getattr(L, n).argtypes is not None]
    assert bound == ["asp_ot_score", "asp_ot_score_workspace_bytes"]
    for fn in bound:
        assert len(getattr(L, fn).argtypes) == _header_arg_count(fn), fn
    # the output struct of the stub has the header's nine pointers, in order
    hdr = open(os.path.join(ROOT, "include", "aspire_b200.h")).read()
    body = re.search(r"typedef struct asp_ot_outputs \{(.*?)\}", hdr, re.S).group(1)
    assert [f for f, _ in ns["_OtOut"]._fields_] == re.findall(r"float\*\s*(\w+);", body)


@pytest.mark.gpu
def test_stub_runs_on_gpu_and_matches_package_binding():
    import torch
    from aspire_b200 import AllPairMaskedWasserstein, rep_len_tup
    ns = _exec_stub()
    g = torch.Generator().manual_seed(0)
    for S in (10, 24):
        B, D = 37, 768
        q = 0.3 * torch.randn(B, S, D, generator=g)
        c = 0.3 * torch.randn(B, S, D, generator=g)
        ql = torch.randint(1, S + 1, (B,), generator=g).tolist()
        cl = torch.randint(1, S + 1, (B,), generator=g).tolist()
        for b in range(B):
            q[b, ql[b]:] = 0
            c[b, cl[b]:] = 0
        qt = rep_len_tup(embed=q.permute(0, 2, 1), abs_lens=ql)
        ct = rep_len_tup(embed=c.permute(0, 2, 1), abs_lens=cl)
        got = ns["ot_distance_b200"](qt, ct)
        want = AllPairMaskedWasserstein({}).compute_distance(query=qt, cand=ct)
        assert torch.allclose(got.cpu(), want.cpu(), rtol=1e-5, atol=1e-5), S
