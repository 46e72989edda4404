"""ctypes binding of libaspire_b200.so (the C ABI declared in include/aspire_b200.h).

There is no CPU fallback: importing the library works without a GPU (so symbol/loader tests run on CPU),
but every compute wrapper requires CUDA tensors and raises if the shared library is missing.
"""
import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# ASPIRE_B200_LIB: developer override (A/B builds of the same ABI under tools/); the product path is the in-tree library
LIB_PATH = os.environ.get("ASPIRE_B200_LIB") or os.path.join(_HERE, "libaspire_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "aspire_b200.h")

_lib = None

c_float_p = ctypes.POINTER(ctypes.c_float)
c_int_p = ctypes.POINTER(ctypes.c_int32)
c_ll_p = ctypes.POINTER(ctypes.c_longlong)


class AspOtOutputs(ctypes.Structure):
    """Mirror of ``asp_ot_outputs`` (include/aspire_b200.h)."""
    _fields_ = [(n, ctypes.c_void_p) for n in
                ("dual", "primal", "f", "g", "alpha", "beta", "neg_cost", "plan", "weighted")]


class AspireB200Error(RuntimeError):
    pass


def declared_symbols():
    """Names of every function declared in include/aspire_b200.h (used by the loader test)."""
    with open(HEADER_PATH) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(asp_[a-z0-9_]+)\s*\(", text)))


def lib():
    """Load (once) and return the shared library; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AspireB200Error(
            f"{LIB_PATH} not found: build the CUDA extension first (python -m aspire_b200.build or "
            f"__graft_entry__.build()). aspire_b200 has no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, ci, cf, cll = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_longlong
    L.asp_version.restype = ci
    L.asp_last_error.restype = ctypes.c_char_p
    L.asp_sm_count.restype = ci
    L.asp_set_option.argtypes = [ctypes.c_char_p, ci]
    L.asp_span_mean_pool.argtypes = [vp, vp, ci, ci, ci, ci, vp, vp, vp]
    L.asp_pair_cost.argtypes = [vp, vp, ci, vp, vp, ci, ci, ci, ci, vp, vp]
    L.asp_l2max.argtypes = [vp, vp, ci, vp, vp, ci, ci, ci, ci, vp, vp, vp, vp]
    L.asp_l2max_ws.argtypes = [vp, vp, ci, vp, vp, ci, ci, ci, ci, vp, vp, vp, vp, ctypes.c_size_t, vp]
    L.asp_l2max_workspace_bytes.argtypes = [ci, ci, ci, ci]
    L.asp_pair_heads.argtypes = [vp, vp, ci, vp, ci, ci, ci, cf, vp, vp, vp, vp]
    L.asp_mix_cls_scores.argtypes = [vp, vp, ci, vp, ci, ci, cf, cf, vp]
    L.asp_ot_sinkhorn.argtypes = [vp, vp, ci, vp, vp, ci, ci, ci, ci, c_float_p, ci, cf, vp,
                                  ctypes.POINTER(AspOtOutputs), vp]
    L.asp_ot_score.argtypes = [vp, vp, ci, vp, vp, ci, ci, ci, ci, c_float_p, ci, cf, ctypes.POINTER(AspOtOutputs), vp,
                               ctypes.c_size_t, vp]
    L.asp_ot_score_workspace_bytes.argtypes = [ci, ci, ci, ci]
    L.asp_ot_score_indexed.argtypes = [vp, vp, ci, vp, vp, vp, ci, ci, ci, ci, c_float_p, ci, cf, ctypes.POINTER(AspOtOutputs), vp]
    L.asp_ot_score_allpairs.argtypes = [vp, vp, ci, vp, vp, ci, ci, ci, ci, c_float_p, ci, cf, vp, vp, ctypes.c_size_t, vp]
    L.asp_ot_score_allpairs_workspace_bytes.argtypes = [ci, ci, ci, ci, ci]
    L.asp_ot_sinkhorn_from_cost.argtypes = [vp, vp, ci, vp, ci, ci, ci, c_float_p, ci, cf,
                                            ctypes.POINTER(AspOtOutputs), vp]
    L.asp_gemm_bf16_tn.argtypes = [vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, vp, vp, vp, vp]
    L.asp_l2max_allpairs.argtypes = [vp, vp, ci, vp, vp, ci, ci, ci, vp, vp, vp, ctypes.c_size_t, vp]
    L.asp_l2max_allpairs_workspace_bytes.argtypes = [ci, ci, ci, ci]
    L.asp_bbox_diameter.argtypes = [vp, cll, vp, cll, ci, vp, vp, vp]
    L.asp_topk.argtypes = [vp, ci, cll, ci, cll, vp, vp, vp]
    L.asp_topk_merge.argtypes = [vp, vp, ci, ci, ci, vp, vp, vp]
    L.asp_topk_workspace_bytes.argtypes = [ci, cll, ci]
    L.asp_topk_ws.argtypes = [vp, ci, cll, ci, cll, ci, vp, vp, vp, vp, ctypes.c_size_t, vp]
    L.asp_topk_merge_packed.argtypes = [vp, ci, ci, ci, vp, vp, vp]
    for name in declared_symbols():
        fn = getattr(L, name)  # AttributeError here == header/library mismatch
        if name not in ("asp_last_error", "asp_launch_count", "asp_ot_score_workspace_bytes", "asp_l2max_allpairs_workspace_bytes",
                        "asp_bert_workspace_bytes", "asp_wordpiece_create", "asp_wordpiece_destroy"):
            fn.restype = ci
    L.asp_launch_count.restype = cll
    L.asp_wordpiece_create.restype = vp
    L.asp_wordpiece_create.argtypes = [vp, vp, ci, ci, ci, vp, ci]
    L.asp_wordpiece_destroy.restype = None
    L.asp_wordpiece_destroy.argtypes = [vp]
    L.asp_wordpiece_set_unicode.argtypes = [vp, vp, vp, vp, vp]
    L.asp_wordpiece_encode.argtypes = [vp, vp, vp, ci, ci, ci, vp, vp, vp]
    L.asp_pack_pool.argtypes = [vp, vp, ci, ci, ci, vp, ci]
    L.asp_abstracts_plan.argtypes = [vp, vp, ci, ci, vp, vp]
    L.asp_abstracts_fill.argtypes = [vp, vp, vp, ci, ci, ci, ci, cll, ci, ci, vp, vp, vp, vp]
    L.asp_ot_score_workspace_bytes.restype = ctypes.c_size_t
    L.asp_l2max_allpairs_workspace_bytes.restype = ctypes.c_size_t
    L.asp_l2max_workspace_bytes.restype = ctypes.c_size_t
    L.asp_topk_workspace_bytes.restype = ctypes.c_size_t
    L.asp_ot_score_allpairs_workspace_bytes.restype = ctypes.c_size_t
    L.asp_bert_workspace_bytes.restype = ctypes.c_size_t
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        msg = lib().asp_last_error().decode("utf8", "replace")
        raise AspireB200Error(f"{what} failed (code {rc}): {msg}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_of(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise AspireB200Error("aspire_b200 kernels need CUDA tensors; there is no CPU fallback")


def launch_count():
    return int(lib().asp_launch_count())


def set_option(key, value):
    check(lib().asp_set_option(key.encode(), int(value)), "asp_set_option")
