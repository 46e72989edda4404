"""K5: per-query top-k on the GPU and the candidate-sharded multi-GPU exchange.

Replaces the head of ``sorted(..., reverse=True)`` at src/evaluation/evaluate.py:76 and
src/pre_process/pp_gen_nearest.py:339.  Order everywhere: (score descending, candidate id ascending) -- the
order a stable descending sort gives a pool listed by id -- so the merged ranking does not depend on how many
shards the pool was split into.

Multi-GPU: every rank scores all queries against its contiguous shard of the candidate pool,
keeps a local top-k with GLOBAL candidate ids, and the only collective of the whole path is one NCCL
all-gather of [Q,k] scores + ids followed by a k-way merge kernel on every rank.
"""
import torch

from . import _abi


def topk(scores, k, base_id=0):
    """scores fp32 CUDA [Q,N] (higher = better) -> (top scores [Q,k], ids int64 [Q,k] = base_id + column)."""
    _abi.require_cuda(scores)
    assert scores.dim() == 2 and scores.dtype == torch.float32
    scores = scores.contiguous()
    Q, N = scores.shape
    out_s = torch.empty((Q, k), dtype=torch.float32, device=scores.device)
    out_i = torch.empty((Q, k), dtype=torch.int64, device=scores.device)
    _abi.check(_abi.lib().asp_topk(_abi.ptr(scores), Q, N, k, int(base_id), _abi.ptr(out_s), _abi.ptr(out_i),
                                   _abi.stream_of(scores.device)), "asp_topk")
    return out_s, out_i


def topk_merge(scores, ids, k):
    """Merge candidate lists: scores [Q,M] fp32, ids [Q,M] int64 (id < 0 = filler) -> best k by (score desc, id asc)."""
    _abi.require_cuda(scores, ids)
    scores, ids = scores.contiguous(), ids.contiguous()
    Q, M = scores.shape
    assert M % k == 0, "topk_merge expects R lists of k entries per query"
    out_s = torch.empty((Q, k), dtype=torch.float32, device=scores.device)
    out_i = torch.empty((Q, k), dtype=torch.int64, device=scores.device)
    _abi.check(_abi.lib().asp_topk_merge(_abi.ptr(scores), _abi.ptr(ids), Q, M // k, k, _abi.ptr(out_s),
                                         _abi.ptr(out_i), _abi.stream_of(scores.device)), "asp_topk_merge")
    return out_s, out_i


def shard_bounds(n_items, world_size, rank):
    """Contiguous, balanced split of [0, n_items): first (n % world) shards get one extra item."""
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_topk(local_scores, local_ids, k, group=None, merge_fn=None):
    """All-gather per-rank top-k lists ([Q,k] fp32 / int64 with global ids) and merge them on every rank.

    One collective per tensor (NCCL all_gather over NVLink on the GPU box; gloo in the CPU tests, where
    ``merge_fn`` supplies a host merge because the CUDA library is absent).
    """
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_scores, local_ids
    world = dist.get_world_size(group)
    gs = [torch.empty_like(local_scores) for _ in range(world)]
    gi = [torch.empty_like(local_ids) for _ in range(world)]
    dist.all_gather(gs, local_scores.contiguous(), group=group)
    dist.all_gather(gi, local_ids.contiguous(), group=group)
    all_s, all_i = torch.cat(gs, dim=1), torch.cat(gi, dim=1)
    return (merge_fn or topk_merge)(all_s, all_i, k)


def host_merge(scores, ids, k):
    """Host restatement of the merge order (used by the gloo tests and as the checker of topk_merge)."""
    Q = scores.shape[0]
    out_s = torch.full((Q, k), float("-inf"))
    out_i = torch.full((Q, k), -1, dtype=torch.int64)
    for q in range(Q):
        rows = [(float(s), int(i)) for s, i in zip(scores[q].tolist(), ids[q].tolist()) if i >= 0]
        rows.sort(key=lambda t: (-t[0], t[1]))
        for j, (s, i) in enumerate(rows[:k]):
            out_s[q, j], out_i[q, j] = s, i
    return out_s, out_i
