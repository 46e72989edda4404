"""CPU, world_size 2 over gloo: the N>1 path's host logic -- shard the pool, local top-k with global ids,
ONE all-gather, merge -- must give the single-process ranking whatever the shard count."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, scores, k, out):
    from aspire_b200.ranking import gather_topk, host_merge, shard_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        Q, N = scores.shape
        lo, hi = shard_bounds(N, world, rank)
        ids = torch.arange(lo, hi).unsqueeze(0).expand(Q, -1).contiguous()
        ls, li = host_merge(scores[:, lo:hi].contiguous(), ids, k)  # stands in for asp_topk on the shard
        ms, mi = gather_topk(ls, li, k, merge_fn=host_merge)
        if rank == 0:
            out.put((ms, mi))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_sharded_topk_matches_single_process():
    from aspire_b200.ranking import host_merge
    g = torch.Generator().manual_seed(0)
    Q, N, k = 7, 501, 20
    scores = torch.randn(Q, N, generator=g)
    scores[:, 100] = scores[:, 400]  # cross-shard exact ties: lower id must win
    scores[:, 250] = scores[:, 251]
    ids = torch.arange(N).unsqueeze(0).expand(Q, -1).contiguous()
    want_s, want_i = host_merge(scores, ids, k)
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, scores, k, out)) for r in range(2)]
    for p in procs:
        p.start()
    got_s, got_i = out.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert torch.equal(got_i, want_i) and torch.equal(got_s, want_s)
