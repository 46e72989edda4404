#!/usr/bin/env python
"""Headline benchmark: OT-scored doc-pairs/sec (10 sentences/doc, 768-d) -- BASELINE.json metric.

Workload = BASELINE configs[1] (otAspire OT scoring: 1 query x 1k candidates, 10 sents/doc, 768-d, blur 0.05,
scaling 0.9, temp 1.0) batched the way an evaluation run presents it: a step scores ``--queries`` (default 64)
queries, EACH against ITS OWN pool of 1k candidates, in ONE launch of the fused kernel (pair cost + softmax marginals
+ masked epsilon-scaling Sinkhorn, dual value) -- 64 000 pairs and 1.97 GB of candidate reps per step.  Steps rotate
over ``--pools`` resident batches, each far larger than the 126 MB L2, so every step streams from HBM.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N>1 (launched by torchrun, one rank per GPU): weak scaling -- every rank holds its own shard of each query's pool
(1k candidates per query per rank), scores it, keeps a per-query local top-100 with global ids, and the ranks
exchange them with ONE NCCL all-gather + merge per step (the path's only collective).

Prints ONE JSON line (rank 0).  ``value`` = device-resident throughput (CUDA events, max over ranks);
``e2e`` = the same workload through the public host-buffer API (pinned H2D of the pools + D2H of the scores inside
the timed region; PCIe-bound); ``roofline`` = HBM roofline of the dominant kernel; ``cpu_baseline`` = the oracle port
of the reference's CPU path (torch, all host threads) on a bounded sample; ``latency_1x1k`` = one query x 1k
candidates per launch (the un-batched shape of configs[1]).  ``--impl reference`` times the CPU path as its own arm.
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
CPU_QUERIES = 8          # queries per step of the CPU arms (bounded sample of the same workload)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def bind_to_gpu_numa_node(device_index):
    """Pin this process to the CPUs local to its GPU (sysfs local_cpulist of the GPU's PCI function) BEFORE the pinned
    host buffers are allocated, so that with several ranks per box every rank's H2D traffic stays on its own socket."""
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as fh:
            spec = fh.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{bdf}: {spec}"
    except Exception as e:  # best effort: a container without sysfs PCI nodes simply keeps its affinity
        return f"not bound ({type(e).__name__})"
    return "not bound"


def ncu_traffic(pairs_per_launch):
    """dram__bytes_read.sum + dram__bytes_write.sum of one ot_fused_kernel launch of this size, in GB, from the
    committed `ncu --set full` capture (profiles/ncu_traffic.json); None when no capture of that launch shape exists."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            t = json.load(fh)
        e = t["ot_fused_kernel"].get(str(pairs_per_launch))
        if e:
            return e["dram_gb"], e["source"]
    except Exception:
        pass
    return None, None


class ClockSampler:
    """nvidia-smi sampler (B200_PROFILING.md clocks line).  It is started BEFORE the warm-up (the process needs a few
    hundred ms to come up) and samples every 20 ms with a timestamp; stop(t0, t1) keeps the samples that fall inside
    the measured window [t0, t1] (host wall-clock, the window is bracketed by synchronisations)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "20", "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    @staticmethod
    def _epoch(ts):
        import datetime
        return datetime.datetime.strptime(ts.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()

    def stop(self, t0, t1):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().strip().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                rows.append((self._epoch(parts[0]), float(parts[1]), float(parts[2]), float(parts[3]),
                             [n for n, v in zip(names, parts[5:9]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        os.unlink(self.f.name)
        inside = [r for r in rows if t0 <= r[0] <= t1]
        window = "measured window"
        if len(inside) < 3:  # very short run: fall back to every sample taken under load (warm-up runs the same steps)
            mx = max([r[2] for r in rows] + [1.0])
            inside = [r for r in rows if r[1] > 0.5 * mx] or rows
            window = "whole run (measured window shorter than 3 samples)"
        reasons = sorted({n for r in inside for n in r[4]})
        return {"sm_mhz": float(np.median([r[1] for r in inside])) if inside else None,
                "sm_max_mhz": max(r[2] for r in inside) if inside else None,
                "power_w": float(np.median([r[3] for r in inside])) if inside else None,
                "reasons": reasons, "samples": len(inside), "window": window}


def make_corpus(n_batches, nq, device, seed):
    """Synthetic abstracts: reps = 0.3*randn, all 10 sentences valid.
    One batch = nq queries [nq,10,768] + their pools [nq*1000,10,768] (candidates of query i at rows i*1000...)."""
    g = torch.Generator(device=device).manual_seed(seed)
    pools = [0.3 * torch.randn(nq * POOL, SENTS, DIM, device=device, generator=g) for _ in range(n_batches)]
    queries = [0.3 * torch.randn(nq, SENTS, DIM, device=device, generator=g) for _ in range(n_batches)]
    return queries, pools


# ------------------------------------------------------------------------------------------ CPU reference
def cpu_reference_step(ar, q, c, threads):
    """One step of the reference's CPU path (oracle port): for every query, compute_distance on [1000,S,D] with the
    query replicated 1000 times, as caching_score does (disent_models.py:274-297); same explicit diameter as the
    GPU arm.  q [nq,S,D], c [nq*1000,S,D]."""
    torch.set_num_threads(threads)
    lens = [SENTS] * POOL
    t0 = time.perf_counter()
    out = []
    for i in range(q.shape[0]):
        out.append(ar.ot_distance(q[i:i + 1].expand(POOL, -1, -1), lens, c[i * POOL:(i + 1) * POOL], lens, blur=BLUR,
                                  scaling=SCALING, temp=TEMP, diameter=DIAMETER))
    return time.perf_counter() - t0, torch.cat(out)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import aspire_ref as ar
    threads = os.cpu_count() or 1
    g = torch.Generator().manual_seed(1234)
    nq = CPU_QUERIES
    qs = [0.3 * torch.randn(nq, SENTS, DIM, generator=g) for _ in range(2)]
    pools = [0.3 * torch.randn(nq * POOL, SENTS, DIM, generator=g) for _ in range(2)]
    for i in range(args.warmup):
        cpu_reference_step(ar, qs[i % 2], pools[i % 2], threads)
    total = 0.0
    for i in range(args.steps):
        dt, _ = cpu_reference_step(ar, qs[i % 2], pools[i % 2], threads)
        total += dt
    value = nq * POOL * args.steps / total
    cfg = workload_config(1, args.queries)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port",
                             "sample": f"each step = {nq} of the workload's queries x {POOL} candidates ({nq * POOL} pairs), "
                                       f"oracle/aspire_ref.py (torch CPU restatement of pair_distances.py:21-92 + "
                                       f"geomloss 0.2.4), one compute_distance call per query"},
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(n_gpus, nq):
    return {"workload": f"otAspire OT scoring (BASELINE configs[1]: 1 query x 1k candidates, 10 sents/doc, 768-d, "
                        f"blur 0.05, scaling 0.9, temp 1.0), {nq} such queries per step in one fused launch (per GPU)",
            "queries_per_step": nq * n_gpus, "pairs_per_step": nq * POOL * n_gpus, "n_eps": None, "diameter": DIAMETER,
            "cache": f"steps rotate over resident batches of {nq * POOL * BYTES_PER_PAIR / 1e9:.2f} GB each (>> 126 MB L2)",
            "parallelism": f"candidate-sharded x{n_gpus}, per-query top-{TOPK} NCCL all-gather per step" if n_gpus > 1
                           else "single GPU"}


# ------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--queries", type=int, default=256, help="queries (x 1k candidates each) per step")
    ap.add_argument("--pools", type=int, default=3, help="resident corpus batches the steps rotate over")
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
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else "single rank: not bound"
    sampler = ClockSampler(local_rank) if rank == 0 else None

    NQ = args.queries
    NP = NQ * POOL  # pairs per step per GPU
    eps = epsilon_schedule(DIAMETER, BLUR, SCALING)
    queries, pools = make_corpus(args.pools, NQ, dev, 1234 + rank)
    if world > 1:  # the step's queries are the same on every rank (rank 0's), broadcast once
        for q in queries:
            dist.broadcast(q, 0)
    q_lens = torch.full((NQ,), SENTS, dtype=torch.int32, device=dev)
    c_lens = torch.full((NP,), SENTS, dtype=torch.int32, device=dev)
    out = {"dual": torch.empty(NP, dtype=torch.float32, device=dev)}
    base_id = rank * POOL

    def step(i):
        p = i % args.pools
        res = ot_scores(queries[p], q_lens, pools[p], c_lens, eps, temp=TEMP, want=("dual",), q_group=POOL, out=out)
        if world > 1:
            s, ids = topk((-res["dual"]).view(NQ, POOL), TOPK, base_id=base_id)
            return gather_topk(s, ids, TOPK)
        return res["dual"]

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # burst figure (informational): the first steps of the process, before the clocks settle under the power cap
    for i in range(3):
        step(i)
    sync_all()
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_burst = min(args.steps, 20)
    b0.record()
    for i in range(n_burst):
        step(i)
    b1.record()
    sync_all()
    burst_ms = b0.elapsed_time(b1) / n_burst
    # clock ramp: ~0.5 s of untimed steps so the timed region runs at load clocks
    t_end = time.time() + 0.5
    i = 0
    while time.time() < t_end:
        step(i)
        i += 1
        torch.cuda.synchronize()
    for i in range(args.warmup):
        step(i)
    sync_all()

    t_wall0 = time.time()
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

    # ---- per-launch timing of the dominant kernel (CUDA events on its stream, same rotation of batches) ----
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(args.steps)]
    for i in range(args.steps):
        p = i % args.pools
        ev[i][0].record()
        ot_scores(queries[p], q_lens, pools[p], c_lens, eps, temp=TEMP, want=("dual",), q_group=POOL, out=out)
        ev[i][1].record()
    torch.cuda.synchronize()
    t_fused = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    clocks = sampler.stop(t_wall0, time.time()) if sampler else None

    # ---- latency of the un-batched shape: ONE query x 1k candidates per launch ----
    lat_out = {"dual": torch.empty(POOL, dtype=torch.float32, device=dev)}
    q1, ql1, cl1 = queries[0][:1].contiguous(), q_lens[:1].contiguous(), c_lens[:POOL].contiguous()
    def one(i):
        p = i % args.pools
        off = (i % NQ) * POOL
        ot_scores(q1, ql1, pools[p][off:off + POOL], cl1, eps, temp=TEMP, want=("dual",), broadcast_query=True, out=lat_out)
    for i in range(20):
        one(i)
    torch.cuda.synchronize()
    la, lb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    la.record()
    for i in range(200):
        one(i)
    lb.record()
    torch.cuda.synchronize()
    lat_ms = la.elapsed_time(lb) / 200

    # ---- e2e: host buffers in, host scores out, through the public API ------------------------------
    from aspire_b200.similarity import score_pools_host
    n_host = 1 if NP * BYTES_PER_PAIR > 4e9 else 2

    def pinned_copy(t):  # device -> pinned host without a pageable intermediate
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t)
        return h
    host_pools = [pinned_copy(pools[k]) for k in range(n_host)]
    host_q = [pinned_copy(queries[k]) for k in range(n_host)]
    host_lens = torch.full((NP,), SENTS, dtype=torch.int32).pin_memory()
    host_qlens = torch.full((NQ,), SENTS, dtype=torch.int32).pin_memory()
    for i in range(2):
        score_pools_host(host_q[i % n_host], host_qlens, host_pools[i % n_host], host_lens, POOL, diameter=DIAMETER)
    sync_all()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        res = score_pools_host(host_q[i % n_host], host_qlens, host_pools[i % n_host], host_lens, POOL, diameter=DIAMETER)
        if world > 1:
            s, ids = topk(res["device_scores"].view(NQ, POOL), TOPK, base_id=base_id)
            gather_topk(s, ids, TOPK)
    sync_all()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    h2d = host_pools[0].numel() * 4 + host_q[0].numel() * 4 + host_lens.numel() * 4 + host_qlens.numel() * 4
    d2h = NP * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = peaks()
    traffic_gb, traffic_src = ncu_traffic(NP)
    achieved = BYTES_PER_PAIR * NP / (t_fused * 1e-3) / 1e9
    cfg = workload_config(world, NQ)
    cfg["n_eps"] = len(eps)
    line = {
        "metric": METRIC, "value": NP * world * args.steps / (ms * 1e-3), "unit": "pairs/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "clocks": clocks,
        "e2e": {"value": NP * world * e2e_steps / e2e_s, "unit": "pairs/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": e2e_steps, "cpu_affinity": numa,
                "api": "aspire_b200.similarity.score_pools_host (pinned host tensors in, chunked H2D overlapped with "
                       "scoring, host scores out)"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic_gb, "traffic_source": traffic_src, "kernel": "ot_fused_v7_kernel", "kernel_ms": t_fused, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": BYTES_PER_PAIR * NP,
                     "step_hbm_frac": BYTES_PER_PAIR * NP / (ms / args.steps * 1e-3) / 1e9 / peak},
        "burst": {"value": NP * world / (burst_ms * 1e-3), "unit": "pairs/s", "steps": n_burst, "ms_per_step": burst_ms,
                  "hbm_frac": BYTES_PER_PAIR * NP / (burst_ms * 1e-3) / 1e9 / peak,
                  "note": "first steps of the process (rank 0's clock), before the SM clock settles under the 1 kW power "
                          "cap; `value` above is the sustained figure"},
        "latency_1x1k": {"pairs_per_launch": POOL, "ms_per_launch": lat_ms, "pairs_per_s": POOL / (lat_ms * 1e-3)},
    }
    if world == 1:
        from oracle import aspire_ref as ar
        threads = os.cpu_count() or 1
        qc = queries[0][:CPU_QUERIES].cpu()
        pc = pools[0][:CPU_QUERIES * POOL].cpu()
        cpu_reference_step(ar, qc[:1], pc[:POOL], threads)
        n, tot = 0, 0.0
        while tot < args.cpu_seconds and n < 50:
            dt, dref = cpu_reference_step(ar, qc, pc, threads)
            tot += dt
            n += 1
        got = ot_scores(queries[0], q_lens, pools[0], c_lens, eps, temp=TEMP, q_group=POOL)["dual"][:CPU_QUERIES * POOL].cpu()
        rel = ((got - dref).abs() / dref.abs().clamp(min=1)).max().item()
        line["cpu_baseline"] = {"value": CPU_QUERIES * POOL * n / tot, "unit": "pairs/s", "cores": threads, "kind": "port",
                                "sample": f"{n} passes over {CPU_QUERIES} of the step's queries x {POOL} candidates "
                                          f"({tot:.1f} s) of oracle/aspire_ref.ot_distance (torch CPU restatement of "
                                          f"pair_distances.py:21-92 + geomloss 0.2.4)",
                                "parity_max_rel_err_vs_gpu": rel}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
