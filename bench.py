#!/usr/bin/env python
"""Headline benchmark: OT-scored doc-pairs/sec (10 sentences/doc, 768-d) -- BASELINE.json metric.

A step = the hot path over one batch: ONE query scored against ITS pool of 1k candidates
(BASELINE configs[1]) with otAspire (pair cost + softmax marginals + masked epsilon-scaling Sinkhorn, dual
value).  Pools rotate over a resident corpus larger than L2, so every step streams its candidates from HBM.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N>1 (launched by torchrun, one rank per GPU): weak scaling -- every rank holds its own shard of each pool
(1k candidates per rank), scores the step's query against it, keeps a local top-100 with global ids, and the
ranks exchange them with ONE NCCL all-gather + merge per step (the path's only collective).

Prints ONE JSON line (rank 0).  ``value`` = device-resident throughput (CUDA events, max over ranks);
``e2e`` = the same workload through the public host-buffer API (pinned H2D of the pool + D2H of the scores
inside the timed region); ``roofline`` = HBM roofline of the dominant kernel; ``cpu_baseline`` = the oracle
port of the reference's CPU path (torch, all host threads) on a bounded sample.
``--impl reference`` times that CPU path as its own arm.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SENTS, DIM, POOL = 10, 768, 1000
BLUR, SCALING, TEMP = 0.05, 0.9, 1.0
DIAMETER = 65.0          # explicit bounding-box diameter shared by all steps/ranks
TOPK = 100
BYTES_PER_PAIR = SENTS * DIM * 4 + 12  # candidate reps once + lens + score = 30 732 B
METRIC = "OT-scored doc-pairs/sec (10 sents, 768-d)"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi sampler running during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().strip().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.f.name)
        busy = [s for s in sm if s > 0.5 * max(mx + [1.0])] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_corpus(n_pools, device, seed):
    """Synthetic abstracts: reps = 0.3*randn, all 10 sentences valid."""
    g = torch.Generator(device=device).manual_seed(seed)
    pools = [0.3 * torch.randn(POOL, SENTS, DIM, device=device, generator=g) for _ in range(n_pools)]
    queries = [0.3 * torch.randn(1, SENTS, DIM, device=device, generator=g) for _ in range(n_pools)]
    return queries, pools


# ------------------------------------------------------------------------------------------ CPU reference
def cpu_reference_step(ar, q, c, lens_q, lens_c, threads):
    """One step of the reference's CPU path (oracle port): compute_distance on [B,S,D] with the query replicated
    B times, as caching_score does (disent_models.py:274-297), same explicit diameter as the GPU arm."""
    torch.set_num_threads(threads)
    t0 = time.perf_counter()
    d = ar.ot_distance(q.expand(c.shape[0], -1, -1), lens_q, c, lens_c, blur=BLUR, scaling=SCALING, temp=TEMP,
                       diameter=DIAMETER)
    return time.perf_counter() - t0, d


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import aspire_ref as ar
    threads = os.cpu_count() or 1
    g = torch.Generator().manual_seed(1234)
    q = 0.3 * torch.randn(1, SENTS, DIM, generator=g)
    pools = [0.3 * torch.randn(POOL, SENTS, DIM, generator=g) for _ in range(4)]
    lq, lc = [SENTS] * POOL, [SENTS] * POOL
    for i in range(args.warmup):
        cpu_reference_step(ar, q, pools[i % 4], lq, lc, threads)
    total = 0.0
    for i in range(args.steps):
        dt, _ = cpu_reference_step(ar, q, pools[i % 4], lq, lc, threads)
        total += dt
    value = POOL * args.steps / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(1),
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port",
                             "sample": f"{args.steps} steps x {POOL} pairs, oracle/aspire_ref.py (torch CPU restatement "
                                       f"of pair_distances.py:21-92 + geomloss 0.2.4), one call per step"},
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(n_gpus):
    return {"workload": "otAspire OT scoring: 1 query x 1k candidates per step (per GPU), 10 sents/doc, 768-d, "
                        "blur 0.05, scaling 0.9, temp 1.0 (BASELINE configs[1])",
            "pairs_per_step": POOL * n_gpus, "n_eps": None, "diameter": DIAMETER,
            "cache": "pools rotate over a resident corpus of 16 x 30.7 MB per GPU (> 126 MB L2)",
            "parallelism": f"candidate-sharded x{n_gpus}, top-{TOPK} NCCL all-gather per step" if n_gpus > 1
                           else "single GPU"}


# ------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--pools", type=int, default=16)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    from aspire_b200 import _abi, epsilon_schedule, ot_scores
    from aspire_b200.ranking import gather_topk, topk

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (native arm) needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _abi.lib()  # fail loudly if the CUDA library is missing

    eps = epsilon_schedule(DIAMETER, BLUR, SCALING)
    queries, pools = make_corpus(args.pools, dev, 1234 + rank)
    if world > 1:  # the step's query is the same on every rank (rank 0's), broadcast once
        for q in queries:
            dist.broadcast(q, 0)
    q_lens = torch.tensor([SENTS], dtype=torch.int32, device=dev)
    c_lens = torch.full((POOL,), SENTS, dtype=torch.int32, device=dev)
    cost_ws = torch.empty((POOL, SENTS, SENTS), dtype=torch.float32, device=dev)
    base_id = rank * POOL

    def step(i):
        p = i % args.pools
        res = ot_scores(queries[p], q_lens, pools[p], c_lens, eps, temp=TEMP, want=("dual",), broadcast_query=True,
                        cost_workspace=cost_ws)
        if world > 1:
            s, ids = topk((-res["dual"])[None], TOPK, base_id=base_id)
            return gather_topk(s, ids, TOPK)
        return res["dual"]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # clock ramp: ~0.3 s of untimed steps so the timed region runs at load clocks
    t_end = time.time() + 0.3
    i = 0
    while time.time() < t_end:
        step(i)
        i += 1
        if i % 64 == 0:
            torch.cuda.synchronize()
    for i in range(args.warmup):
        step(i)
    sync_all()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = _abi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    sync_all()
    launches = _abi.launch_count() - launches0
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())

    # ---- per-kernel timing of the same steps: which kernel dominates, and its HBM roofline ----------
    L = _abi.lib()
    st = _abi.stream_of(dev)
    import ctypes
    eps32 = np.asarray(eps, dtype=np.float32)
    dual = torch.empty(POOL, device=dev)
    outs = _abi.AspOtOutputs(dual=dual.data_ptr())
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    for i in range(args.steps):
        p = i % args.pools
        ev[i][0].record()
        _abi.check(L.asp_pair_cost(_abi.ptr(queries[p]), _abi.ptr(q_lens), 1, _abi.ptr(pools[p]), _abi.ptr(c_lens), POOL,
                                   SENTS, SENTS, DIM, _abi.ptr(cost_ws), st), "asp_pair_cost")
        ev[i][1].record()
        _abi.check(L.asp_ot_sinkhorn_from_cost(_abi.ptr(cost_ws), _abi.ptr(q_lens), 1, _abi.ptr(c_lens), POOL, SENTS, SENTS,
                                               eps32.ctypes.data_as(_abi.c_float_p), len(eps32), TEMP,
                                               ctypes.byref(outs), st), "asp_ot_sinkhorn_from_cost")
        ev[i][2].record()
    torch.cuda.synchronize()
    t_cost = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    t_sink = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    clocks = sampler.stop() if sampler else None

    # ---- e2e: host buffers in, host scores out, through the public API ------------------------------
    from aspire_b200.similarity import score_pool_tensors
    host_pools = [p.cpu().pin_memory() for p in pools[:4]]
    host_q = [q.cpu().pin_memory() for q in queries[:4]]
    host_lens = torch.full((POOL,), SENTS, dtype=torch.int32).pin_memory()
    for i in range(3):
        score_pool_tensors(host_q[i % 4], host_pools[i % 4], host_lens, diameter=DIAMETER)
    sync_all()
    e2e_steps = max(10, min(args.steps, 100))
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        out = score_pool_tensors(host_q[i % 4], host_pools[i % 4], host_lens, diameter=DIAMETER)
        if world > 1:
            s, ids = topk(out["device_scores"][None], TOPK, base_id=base_id)
            gather_topk(s, ids, TOPK)
    sync_all()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    h2d = host_pools[0].numel() * 4 + host_q[0].numel() * 4 + host_lens.numel() * 4
    d2h = POOL * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    dom_name, dom_ms = ("pair_cost_kernel", t_cost) if t_cost >= t_sink else ("sinkhorn kernel", t_sink)
    achieved = BYTES_PER_PAIR * POOL / (dom_ms * 1e-3) / 1e9
    cfg = workload_config(world)
    cfg["n_eps"] = len(eps)
    line = {
        "metric": METRIC, "value": POOL * world * args.steps / (ms * 1e-3), "unit": "pairs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "clocks": clocks,
        "e2e": {"value": POOL * world * e2e_steps / e2e_s, "unit": "pairs/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "api": "aspire_b200.similarity.score_pool_tensors (pinned host tensors in, host scores out)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "kernel": dom_name, "kernel_ms": dom_ms, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": BYTES_PER_PAIR * POOL,
                     "per_kernel_ms": {"pair_cost_kernel": t_cost, "sinkhorn": t_sink},
                     "step_hbm_frac": BYTES_PER_PAIR * POOL / (ms / args.steps * 1e-3) / 1e9 / peak},
    }
    if world == 1:
        from oracle import aspire_ref as ar
        threads = os.cpu_count() or 1
        qc = queries[0].cpu()
        pc = pools[0].cpu()
        lq, lc = [SENTS] * POOL, [SENTS] * POOL
        cpu_reference_step(ar, qc, pc, lq, lc, threads)
        n, tot = 0, 0.0
        while tot < args.cpu_seconds and n < 200:
            dt, dref = cpu_reference_step(ar, qc, pc, lq, lc, threads)
            tot += dt
            n += 1
        got = ot_scores(queries[0], q_lens, pools[0], c_lens, eps, temp=TEMP, broadcast_query=True)["dual"].cpu()
        rel = ((got - dref).abs() / dref.abs().clamp(min=1)).max().item()
        line["cpu_baseline"] = {"value": POOL * n / tot, "unit": "pairs/s", "cores": threads, "kind": "port",
                                "sample": f"{n} calls x {POOL} pairs ({tot:.1f} s) of oracle/aspire_ref.ot_distance "
                                          f"(torch CPU restatement of pair_distances.py:21-92 + geomloss 0.2.4)",
                                "parity_max_rel_err_vs_gpu": rel}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
