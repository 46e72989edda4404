"""CPU: the C-ABI shared library loads and exports every symbol include/aspire_b200.h declares."""
import ctypes
import os

import pytest

from aspire_b200 import _abi


def test_header_declares_expected_entry_points():
    names = _abi.declared_symbols()
    for must in ("asp_version", "asp_last_error", "asp_span_mean_pool", "asp_pair_cost", "asp_l2max",
                 "asp_ot_sinkhorn", "asp_ot_sinkhorn_from_cost", "asp_bbox_diameter", "asp_topk", "asp_topk_merge"):
        assert must in names


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_abi.LIB_PATH):
        from aspire_b200 import build
        build.build()
    L = ctypes.CDLL(_abi.LIB_PATH)
    for name in _abi.declared_symbols():
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    L.asp_version.restype = ctypes.c_int
    assert L.asp_version() >= 100
    L.asp_last_error.restype = ctypes.c_char_p
    assert L.asp_last_error() is not None


def test_argument_validation_without_gpu():
    """Bad arguments are rejected before any CUDA call (so this runs on the CPU box)."""
    L = _abi.lib()
    rc = L.asp_topk(None, 1, 10, 5, 0, None, None, None)
    assert rc == -1 and b"NULL" in L.asp_last_error()
    rc = L.asp_span_mean_pool(None, None, 1, 1, 768, 1, None, None, None)
    assert rc == -1
    rc = L.asp_set_option(b"no_such_option", 1)
    assert rc == -1 and b"unknown key" in L.asp_last_error()
    # every documented switch accepts its documented values (and its default leaves the process as it was) and rejects the next one
    for key, ok, bad, default in ((b"attn_tc", (0, 1, 2, 3, 4, 5), 6, 5), (b"ln_on_read", (0, 1), 2, 1), (b"gemm_pair", (-1, 0, 1, 2), 3, -1),
                                  (b"span_tma", (0, 1), 2, 1), (b"oa_warps", (8, 12), 10, 12)):
        for v in ok:
            assert L.asp_set_option(key, v) == 0, (key, v)
        assert L.asp_set_option(key, bad) == -1, (key, bad)
        assert L.asp_set_option(key, default) == 0


def test_no_cpu_fallback():
    import torch
    from aspire_b200 import distances
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    q = distances.rep_len_tup(embed=torch.zeros(1, 8, 2), abs_lens=[2])
    with pytest.raises(_abi.AspireB200Error):
        distances.allpair_masked_dist_l2max(q, q)
    with pytest.raises(_abi.AspireB200Error):
        distances.AllPairMaskedWasserstein({}).compute_distance(q, q)
