"""Times asp_gemm_bf16_tn on the four BERT-base projection shapes for each kernel choice (asp_set_option
"gemm_kernel": 0 one tile per CTA, 1 persistent 128-wide, 2 persistent 256-wide, 3 default pick) and checks every
result against torch.  usage: python tools/gemm_bench.py [--tokens 8192] [--x3]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aspire_b200 import _abi  # noqa: E402

EPI_BF16, EPI_GELU, EPI_RESID, EPI_F32 = 0, 1, 2, 3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tokens", type=int, nargs="+", default=[8192, 2048])
    ap.add_argument("--x3", action="store_true")
    ap.add_argument("--modes", type=int, nargs="+", default=[0, 1, 2, 3])
    ap.add_argument("--shapes", nargs="+", default=None, help="custom shapes as N,K,epilogue (instead of the BERT set)")
    ap.add_argument("--clusters", type=int, nargs="+", default=[1], help="CTAs per cluster sharing W by TMA multicast")
    ap.add_argument("--pairs", type=int, nargs="+", default=[0], help="CTA-pair kernel: 0 off, 1 128-wide, 2 256-wide")
    ap.add_argument("--reps", type=int, default=8, help="timed repetitions (0: one checked launch only, for ncu)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1)
    shapes = [("qkv", 2304, 768, EPI_BF16), ("attn_out", 768, 768, EPI_RESID), ("ffn1", 3072, 768, EPI_GELU),
              ("ffn2", 768, 3072, EPI_RESID)]
    if a.shapes:
        shapes = [(f"n{n}k{k}e{e}", n, k, e) for n, k, e in (map(int, x.split(",")) for x in a.shapes)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for M in a.tokens:
        for name, N, K, epi in shapes:
            x = torch.randn(M, K, device=dev, generator=g)
            w = torch.randn(N, K, device=dev, generator=g) / K ** 0.5
            bias = torch.randn(N, device=dev, generator=g)
            resid = torch.randn(M, N, device=dev, generator=g)
            xh, wh = x.bfloat16().contiguous(), w.bfloat16().contiguous()
            xl = (x - xh.float()).bfloat16().contiguous() if a.x3 else None
            wl = (w - wh.float()).bfloat16().contiguous() if a.x3 else None
            out_hi = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
            out_f = torch.empty(M, N, dtype=torch.float32, device=dev)
            if a.x3:
                ref = x.double() @ w.double().T + bias.double()
            else:
                ref = xh.double() @ wh.double().T + bias.double()
            if epi == EPI_GELU:
                ref = torch.nn.functional.gelu(ref)
            elif epi == EPI_RESID:
                ref = ref + resid.double()
            line = f"M={M:5d} {name:8s} N={N:4d} K={K:4d}"
            for mode, cl, pair in [(m, c, p) for m in a.modes for c in (a.clusters if m else [1])
                                   for p in (a.pairs if m else [0])]:
                _abi.set_option("gemm_kernel", mode)
                _abi.set_option("gemm_cluster", cl)
                _abi.set_option("gemm_pair", pair)

                def run():
                    _abi.check(_abi.lib().asp_gemm_bf16_tn(
                        _abi.ptr(xh), _abi.ptr(xl), _abi.ptr(wh), _abi.ptr(wl), _abi.ptr(bias),
                        _abi.ptr(resid) if epi == EPI_RESID else None, M, N, K, epi,
                        _abi.ptr(out_hi) if epi in (EPI_BF16, EPI_GELU) else None, None,
                        _abi.ptr(out_f) if epi in (EPI_RESID, EPI_F32) else None, _abi.stream_of(dev)), "gemm")
                out_hi.zero_(); out_f.zero_()
                run()
                torch.cuda.synchronize()
                got = (out_hi if epi in (EPI_BF16, EPI_GELU) else out_f).double()
                err = ((got - ref).abs().max() / ref.abs().max()).item()
                if a.reps == 0:
                    line += f" | m{mode}c{cl}p{pair}: err {err:.1e}"
                    continue
                ts = []
                for _ in range(a.reps):
                    flush.fill_(1)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); run(); e1.record(); torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1) * 1e3)
                ts.sort()
                us = ts[len(ts) // 2]
                tf = 2.0 * M * N * K * (3 if a.x3 else 1) / us / 1e6
                line += f" | m{mode}c{cl}p{pair}: {us:6.1f} us {tf:5.0f} TF err {err:.0e}"
            print(line, flush=True)
    _abi.set_option("gemm_kernel", 3)
    _abi.set_option("gemm_cluster", 1)
    _abi.set_option("gemm_pair", 0)


if __name__ == "__main__":
    main()
