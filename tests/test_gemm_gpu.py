"""GPU parity of the tcgen05 GEMM (asp_gemm_bf16_tn) against plain PyTorch fp32/fp64 matmuls of the same operands.

bf16 mode: operands are bf16, products are exact in fp32, so the only difference to ``A.float() @ W.float().T`` is
the accumulation order -> 2e-3 relative to the row scale is generous.  bf16x3 mode: operands are (hi, lo) splits of
fp32 values; against an fp64 matmul of the original fp32 values the error must be ~2^-16 relative per product
(tolerance 3e-5 of the |A||W| scale), i.e. ~100x tighter than plain bf16.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

EPI_BF16, EPI_GELU, EPI_RESID, EPI_F32 = 0, 1, 2, 3


def _gemm(a_hi, a_lo, w_hi, w_lo, bias, resid, epi, want_lo=False):
    from aspire_b200 import _abi
    M, K = a_hi.shape
    N = w_hi.shape[0]
    dev = a_hi.device
    out_hi = torch.empty((M, N), dtype=torch.bfloat16, device=dev) if epi in (EPI_BF16, EPI_GELU) else None
    out_lo = torch.empty((M, N), dtype=torch.bfloat16, device=dev) if (out_hi is not None and want_lo) else None
    out_f = torch.empty((M, N), dtype=torch.float32, device=dev) if epi in (EPI_RESID, EPI_F32) else None
    _abi.check(_abi.lib().asp_gemm_bf16_tn(_abi.ptr(a_hi), _abi.ptr(a_lo), _abi.ptr(w_hi), _abi.ptr(w_lo), _abi.ptr(bias),
                                          _abi.ptr(resid), M, N, K, epi, _abi.ptr(out_hi), _abi.ptr(out_lo),
                                          _abi.ptr(out_f), _abi.stream_of(dev)), "asp_gemm_bf16_tn")
    torch.cuda.synchronize()
    return out_hi, out_lo, out_f


def _split(x):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi.contiguous(), lo.contiguous()


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 128, 128), (128, 256, 768), (300, 768, 768), (1000, 2304, 768),
                                   (777, 768, 3072)])
def test_gemm_bf16_fp32_out(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    _, _, out = _gemm(a, None, w, None, bias, None, EPI_F32)
    ref = a.float() @ w.float().T + bias
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2e-3 * scale, f"max err {err} vs scale {scale}"


def test_gemm_epilogues():
    M, N, K = 260, 768, 768
    g = torch.Generator(device="cuda").manual_seed(7)
    a = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g)
    ref = a.float() @ w.float().T + bias
    hi, lo, _ = _gemm(a, None, w, None, bias, None, EPI_BF16, want_lo=True)
    assert (hi.float() + lo.float() - ref).abs().max().item() <= 2e-3 * ref.abs().max().item()
    assert (hi.float() - ref).abs().max().item() <= 8e-3 * ref.abs().max().item()  # bf16 rounding of the output
    hi, _, _ = _gemm(a, None, w, None, bias, None, EPI_GELU)
    refg = torch.nn.functional.gelu(ref)
    assert (hi.float() - refg).abs().max().item() <= 8e-3 * refg.abs().max().item()
    _, _, out = _gemm(a, None, w, None, bias, resid, EPI_RESID)
    assert (out - (ref + resid)).abs().max().item() <= 2e-3 * ref.abs().max().item()


@pytest.mark.parametrize("M,N,K", [(256, 768, 768), (500, 768, 3072)])
def test_gemm_bf16x3_is_fp32_equivalent(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(11)
    a = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    a_hi, a_lo = _split(a)
    w_hi, w_lo = _split(w)
    _, _, out = _gemm(a_hi, a_lo, w_hi, w_lo, None, None, EPI_F32)
    ref = (a.double() @ w.double().T)
    scale = (a.abs().double() @ w.abs().double().T)  # |A||W|: the natural error scale of a dot product
    rel = ((out.double() - ref).abs() / scale).max().item()
    _, _, out1 = _gemm(a_hi, None, w_hi, None, None, None, EPI_F32)
    rel1 = ((out1.double() - ref).abs() / scale).max().item()
    assert rel <= 3e-5, f"bf16x3 rel err {rel}"
    assert rel1 > 10 * rel, f"plain bf16 ({rel1}) should be far less accurate than bf16x3 ({rel})"


@pytest.mark.parametrize("epi", [EPI_BF16, EPI_GELU, EPI_RESID, EPI_F32])
def test_gemm_kernel_variants_agree(epi):
    """Every kernel choice (one tile per CTA, persistent 128- / 192- / 256-wide, 1/2/4-CTA clusters sharing W by TMA
    multicast, CTA pairs on tcgen05 cta_group::2)
    runs the same K order into the same fp32 accumulator, so the results must agree to the last bit -- including the
    ragged last row block (M = 1000) and the cluster ranks whose row block lies entirely past M."""
    from aspire_b200 import _abi
    M, N, K = 1000, 768, 768
    g = torch.Generator(device="cuda").manual_seed(23)
    a = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    a_hi, a_lo = _split(a)
    w_hi, w_lo = _split(w)
    bias = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g) if epi == EPI_RESID else None
    try:
        base = None
        for mode, cluster, pair in [(0, 1, 0), (1, 1, 0), (2, 1, 0), (4, 1, 0), (1, 2, 0), (2, 2, 0), (1, 4, 0), (2, 4, 0),
                                    (3, 1, 0), (3, 1, 1), (3, 1, 2)]:
            _abi.set_option("gemm_kernel", mode)
            _abi.set_option("gemm_cluster", cluster)
            _abi.set_option("gemm_pair", pair)  # 1 / 2: CTA pairs (tcgen05 cta_group::2), 128- / 256-wide pair tiles
            for lo in (False, True):
                out = _gemm(a_hi, a_lo if lo else None, w_hi, w_lo if lo else None, bias, resid, epi, want_lo=lo)
                got = [t.clone() for t in out if t is not None]
                key = (lo,)
                if base is None:
                    base = {}
                if key not in base:
                    base[key] = got
                else:
                    for x, y in zip(base[key], got):
                        assert torch.equal(x, y), (f"mode {mode} cluster {cluster} pair {pair} lo={lo} differs from the "
                                                   "one-tile kernel")
    finally:
        _abi.set_option("gemm_kernel", 3)
        _abi.set_option("gemm_cluster", 1)
        _abi.set_option("gemm_pair", -1)


@pytest.mark.parametrize("epi,K", [(EPI_GELU, 256), (EPI_RESID, 2048)])
def test_gemm_pair_auto_rule_large_m(epi, K):
    """gemm_pair = -1 (default) sends the bf16-output GEMMs and the K >= 2048 residual GEMM of >= 16384 rows to the CTA-pair kernel
    (256-wide pair tiles); an odd number of 128-row blocks (16500 rows = 129 blocks) leaves the last pair half empty.
    Bit-identical to the one-CTA kernel."""
    from aspire_b200 import _abi
    M, N = 16500, 768
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randn(M, K, device="cuda", generator=g)
    w = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    a_hi, _ = _split(a)
    w_hi, _ = _split(w)
    bias = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g) if epi == EPI_RESID else None
    out = {}
    try:
        for pair in (-1, 0):
            _abi.set_option("gemm_pair", pair)
            out[pair] = [t.clone() for t in _gemm(a_hi, None, w_hi, None, bias, resid, epi, want_lo=False) if t is not None]
    finally:
        _abi.set_option("gemm_pair", -1)
    for x, y in zip(out[-1], out[0]):
        assert torch.equal(x, y)
