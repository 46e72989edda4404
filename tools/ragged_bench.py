"""Developer benchmark: the fused otAspire kernel on ragged documents (3-10 sentences) vs the all-10 bench shape."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from aspire_b200 import _abi, ot_scores, epsilon_schedule


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(1)
    eps = epsilon_schedule(65.0, 0.05, 0.9)
    N, POOL = 256000, 1000
    q = 0.3 * torch.randn(N // POOL, 10, 768, device=dev, generator=g)
    cs = [0.3 * torch.randn(N, 10, 768, device=dev, generator=g) for _ in range(2)]
    out = {"dual": torch.empty(N, device=dev)}
    for name, lo in (("all 10 sentences", 10), ("3-10 sentences (uniform)", 3), ("7-10 sentences", 7)):
        ql = torch.randint(lo, 11, (N // POOL,), device=dev, generator=g).int()
        cl = torch.randint(lo, 11, (N,), device=dev, generator=g).int()
        for c in cs:
            c.mul_((torch.arange(10, device=dev)[None, :] < cl[:, None])[:, :, None])  # zero the padded rows
        st = {"i": 0}
        def call():
            st["i"] += 1
            ot_scores(q, ql, cs[st["i"] % 2], cl, eps, q_group=POOL, out=out)
        t = timeit(call)
        by = float(cl.sum().item()) * 768 * 4 + 12 * N
        print(f"{name:28s}: {t:.3f} ms  {N / t * 1e3:.3e} pairs/s  {by / t / 1e6:.0f} GB/s of valid rows", flush=True)
        cs = [0.3 * torch.randn(N, 10, 768, device=dev, generator=g) for _ in range(2)]


if __name__ == "__main__":
    main()
