"""Developer timer under the power cap: the fused 10-sentence kernel back to back for ~1000 steps, twice, same data.
A/B two builds by running it once per library (ASPIRE_B200_LIB=experiments/lib/libaspire_b200_<name>.so, see
tools/build_variant.sh)."""
import os, sys, time
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
name = os.path.basename(os.environ.get("ASPIRE_B200_LIB", "in-tree"))
if "ASP_TC" in os.environ:
    _abi.set_option("ot_fused_tc", int(os.environ["ASP_TC"]))
    name += f" ot_fused_tc={os.environ['ASP_TC']}"
for rep in range(2):
    t_end = time.time() + 1.0
    i = 0
    while time.time() < t_end:  # settle
        ot_scores(q, ql, cs[i % 3], cl, eps, q_group=POOL, out=out); i += 1
        torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = int(os.environ.get("ASP_STEPS", 1000))
    a.record()
    for k in range(n):
        ot_scores(q, ql, cs[k % 3], cl, eps, q_group=POOL, out=out)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    print(f"{name} rep {rep}: {ms:.3f} ms/step  {N / ms * 1e3:.3e} pairs/s sustained over {n} steps", flush=True)
