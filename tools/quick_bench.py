"""Developer micro-benchmark (not the judged bench): fused otAspire kernel vs the two-kernel path."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from aspire_b200 import _abi, ot_scores, l2max_scores, epsilon_schedule


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
    print("n_eps", len(eps))
    sizes = [int(a) for a in sys.argv[1:]] or [1000, 8000, 64000, 256000]
    for N in sizes:
        nq = max(N // 1000, 1)
        q = 0.3 * torch.randn(nq, 10, 768, device=dev, generator=g)
        ql = torch.full((nq,), 10, dtype=torch.int32, device=dev)
        # two buffers > L2 so that consecutive timed calls do not hit in L2
        cs = [0.3 * torch.randn(N, 10, 768, device=dev, generator=g) for _ in range(2 if N * 30720 > 64e6 else 8)]
        cl = torch.full((N,), 10, dtype=torch.int32, device=dev)
        out = {"dual": torch.empty(N, device=dev)}
        ws = torch.empty(N * 100, device=dev)
        state = {"i": 0}

        def call():
            state["i"] += 1
            ot_scores(q, ql, cs[state["i"] % len(cs)], cl, eps, q_group=1000 if N >= 1000 else N, out=out, cost_workspace=ws)
        res = {}
        for k, name in ((0, "fused"), (1, "cost+warp")):
            if name == "cost+warp" and N > 300000:
                continue
            _abi.set_option("ot_kernel", k)
            res[name] = timeit(call)
        _abi.set_option("ot_kernel", 0)
        t_l2 = timeit(lambda: l2max_scores(q[:1], ql[:1], cs[0], cl, broadcast_query=True))
        gb = N * 30732 / 1e9
        print(f"N={N:8d} " + "  ".join(f"{n} {t:8.3f} ms ({N / t * 1e3:.3e} pairs/s, {gb / t * 1e3:7.1f} GB/s)"
                                        for n, t in res.items()) + f"   l2max {t_l2:8.3f} ms ({gb / t_l2 * 1e3:7.1f} GB/s)",
              flush=True)
        del cs


if __name__ == "__main__":
    main()
