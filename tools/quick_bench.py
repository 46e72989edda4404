"""Developer micro-benchmark (not the judged bench): times the scoring kernels at a few batch sizes."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from aspire_b200 import _abi, ot_scores, l2max_scores, epsilon_schedule
import ctypes

def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
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
    q = 0.3 * torch.randn(1, 10, 768, device=dev, generator=g)
    ql = torch.tensor([10], dtype=torch.int32, device=dev)
    eps = epsilon_schedule(65.0, 0.05, 0.9)
    print("n_eps", len(eps))
    L = _abi.lib()
    for N in [1000, 10000, 100000, 1000000]:
        c = 0.3 * torch.randn(N, 10, 768, device=dev, generator=g)
        cl = torch.full((N,), 10, dtype=torch.int32, device=dev)
        cost = torch.empty(N, 10, 10, device=dev)
        dual = torch.empty(N, device=dev)
        st = _abi.stream_of(dev)
        t_cost = timeit(lambda: _abi.check(L.asp_pair_cost(_abi.ptr(q), _abi.ptr(ql), 1, _abi.ptr(c), _abi.ptr(cl), N, 10, 10, 768, _abi.ptr(cost), st), "cost"))
        eps32 = np.asarray(eps, dtype=np.float32)
        outs = _abi.AspOtOutputs(dual=dual.data_ptr())
        res = {}
        for k in (1, 2):
            _abi.set_option("ot_kernel", k)
            res[k] = timeit(lambda: _abi.check(L.asp_ot_sinkhorn_from_cost(_abi.ptr(cost), _abi.ptr(ql), 1, _abi.ptr(cl), N, 10, 10, eps32.ctypes.data_as(_abi.c_float_p), len(eps32), 1.0, ctypes.byref(outs), st), "sink"))
        _abi.set_option("ot_kernel", 0)
        t_l2 = timeit(lambda: l2max_scores(q, ql, c, cl, broadcast_query=True))
        gb = N * 30732 / 1e9
        print(f"N={N:8d} cost {t_cost:8.3f} ms ({gb/t_cost*1e3:7.1f} GB/s)  sink_warp {res[1]:8.3f} ms  sink_thread {res[2]:8.3f} ms "
              f" best-total {(t_cost+min(res.values())):8.3f} ms -> {N/(t_cost+min(res.values()))*1e3:.3e} pairs/s   l2max {t_l2:8.3f} ms")
        del c, cost

if __name__ == "__main__":
    main()
