"""Developer A/B under the power cap: both fused kernels run back to back for ~3 s each, twice, same data."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from aspire_b200 import _abi, ot_scores, epsilon_schedule

dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(1)
eps = epsilon_schedule(65.0, 0.05, 0.9)
N, POOL = 256000, 1000
q = 0.3 * torch.randn(N // POOL, 10, 768, device=dev, generator=g)
cs = [0.3 * torch.randn(N, 10, 768, device=dev, generator=g) for _ in range(3)]
ql = torch.full((N // POOL,), 10, dtype=torch.int32, device=dev)
cl = torch.full((N,), 10, dtype=torch.int32, device=dev)
out = {"dual": torch.empty(N, device=dev)}
for rep in range(2):
    for mode in (1, 0):
        _abi.set_option("ot_fused_mode", mode)
        t_end = time.time() + 1.0
        i = 0
        while time.time() < t_end:  # settle
            ot_scores(q, ql, cs[i % 3], cl, eps, q_group=POOL, out=out); i += 1
            torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 1000
        a.record()
        for k in range(n):
            ot_scores(q, ql, cs[k % 3], cl, eps, q_group=POOL, out=out)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / n
        print(f"rep {rep} mode {'v7' if mode else 'v6'}: {ms:.3f} ms/step  {N / ms * 1e3:.3e} pairs/s sustained over {n} steps", flush=True)
_abi.set_option("ot_fused_mode", 1)
