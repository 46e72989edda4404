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


def topk(scores, k, base_id=0, negate=False, want_packed=False):
    """scores fp32 CUDA [Q,N] (higher = better) -> (top scores [Q,k], ids int64 [Q,k] = base_id + column).

    ``negate``: rank by -scores and return the negated values (OT distances: smaller = better) -- no separate
    negation pass.  ``want_packed``: additionally return the list as sortable 64-bit keys [Q,k] (int64 storage of the
    library's uint64: score bits << 32 | ~id), the one tensor ``gather_topk`` exchanges.  k <= 128 runs the single-pass
    kernels (the matrix is read from HBM once); larger k the radix-select kernel."""
    _abi.require_cuda(scores)
    assert scores.dim() == 2 and scores.dtype == torch.float32
    scores = scores.contiguous()
    Q, N = scores.shape
    L = _abi.lib()
    out_s = torch.empty((Q, k), dtype=torch.float32, device=scores.device)
    out_i = torch.empty((Q, k), dtype=torch.int64, device=scores.device)
    need = int(L.asp_topk_workspace_bytes(Q, N, k))
    if need == 0:
        if want_packed:
            raise _abi.AspireB200Error("topk: packed keys need k <= 128")
        src = -scores if negate else scores
        _abi.check(L.asp_topk(_abi.ptr(src), Q, N, k, int(base_id), _abi.ptr(out_s), _abi.ptr(out_i),
                              _abi.stream_of(scores.device)), "asp_topk")
        return out_s, out_i
    ws = torch.empty(need, dtype=torch.uint8, device=scores.device)
    packed = torch.empty((Q, k), dtype=torch.int64, device=scores.device) if want_packed else None
    _abi.check(L.asp_topk_ws(_abi.ptr(scores), Q, N, k, int(base_id), int(bool(negate)), _abi.ptr(out_s), _abi.ptr(out_i),
                             _abi.ptr(packed), _abi.ptr(ws), need, _abi.stream_of(scores.device)), "asp_topk_ws")
    return (out_s, out_i, packed) if want_packed else (out_s, out_i)


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


def topk_merge_packed(gathered, k):
    """gathered int64 [R,Q,k] packed keys exactly as ``all_gather_into_tensor`` lays them out -> (scores, ids) [Q,k]."""
    _abi.require_cuda(gathered)
    R, Q, kk = gathered.shape
    assert kk == k and gathered.is_contiguous()
    out_s = torch.empty((Q, k), dtype=torch.float32, device=gathered.device)
    out_i = torch.empty((Q, k), dtype=torch.int64, device=gathered.device)
    _abi.check(_abi.lib().asp_topk_merge_packed(_abi.ptr(gathered), R, Q, k, _abi.ptr(out_s), _abi.ptr(out_i),
                                                _abi.stream_of(gathered.device)), "asp_topk_merge_packed")
    return out_s, out_i


def pack_keys(scores, ids):
    """Host restatement of the library's packed key: order-preserving score bits << 32 | (2^32-1 - id); 0 = filler.
    Used by the gloo tests (the kernels produce the same keys on the device)."""
    b = scores.contiguous().view(torch.int32).to(torch.int64) & 0xffffffff
    b = torch.where(scores == 0, torch.zeros_like(b), b)                      # -0 == +0
    key = torch.where(b >= 0x80000000, (~b) & 0xffffffff, b | 0x80000000)
    key = torch.where(torch.isnan(scores), torch.zeros_like(key), key)
    packed = (key << 32) | (0xffffffff - ids)
    return torch.where(ids < 0, torch.zeros_like(packed), packed)            # int64 storage of the uint64 key


def unpack_keys(packed):
    """Inverse of ``pack_keys`` -> (scores fp32, ids int64); fillers give (-inf, -1)."""
    key = (packed >> 32) & 0xffffffff
    b = torch.where(key >= 0x80000000, key & 0x7fffffff, (~key) & 0xffffffff)
    scores = torch.where(b >= 0x80000000, b - (1 << 32), b).to(torch.int32).view(torch.float32)
    ids = 0xffffffff - (packed & 0xffffffff)
    fill = packed == 0
    return torch.where(fill, torch.full_like(scores, float("-inf")), scores), torch.where(fill, torch.full_like(ids, -1), ids)


def host_merge_packed(gathered, k):
    """Host merge of [R,Q,k] packed keys (unsigned descending order) -- the checker of asp_topk_merge_packed."""
    R, Q, _ = gathered.shape
    flat = gathered.permute(1, 0, 2).reshape(Q, R * k)
    # unsigned 64-bit descending order on int64 storage: flip the sign bit
    order = torch.argsort(flat ^ (-(1 << 63)), dim=1, descending=True, stable=True)[:, :k]
    return unpack_keys(torch.gather(flat, 1, order))


def shard_bounds(n_items, world_size, rank):
    """Contiguous, balanced split of [0, n_items): first (n % world) shards get one extra item."""
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_topk(local_scores, local_ids, k, group=None, merge_fn=None, packed=None, out=None):
    """All-gather per-rank top-k lists ([Q,k] with GLOBAL ids) and merge them on every rank.

    ONE collective: the lists travel as packed 64-bit keys (``packed`` from ``topk(..., want_packed=True)``, else built
    here) through one ``all_gather_into_tensor`` into a [R,Q,k] buffer (``out``: preallocated, optional) that
    ``asp_topk_merge_packed`` reads in place -- no second collective for the ids, no concatenation.  NCCL over NVLink on
    the GPU box; gloo in the CPU tests, where the host restatement of the merge runs (``merge_fn`` kept for callers that
    supply their own host merge of (scores, ids) lists)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_scores, local_ids
    world = dist.get_world_size(group)
    if packed is None:
        packed = pack_keys(local_scores, local_ids)
    Q = packed.shape[0]
    if out is None:
        out = torch.empty((world, Q, k), dtype=torch.int64, device=packed.device)
    if packed.is_cuda:
        dist.all_gather_into_tensor(out, packed.contiguous(), group=group)
        return topk_merge_packed(out, k)
    dist.all_gather(list(out.unbind(0)), packed.contiguous(), group=group)  # gloo: list API
    if merge_fn is not None:
        s, i = unpack_keys(out)
        return merge_fn(s.permute(1, 0, 2).reshape(Q, world * k), i.permute(1, 0, 2).reshape(Q, world * k), k)
    return host_merge_packed(out, k)


def rank_corpus_ot(q, q_lens, c, c_lens, eps_list, k, base_id=0, chunk=25000, temp=1.0, group=None):
    """Every query document against a resident corpus (shard) with otAspire, top-k per query -- the whole-corpus use of
    compute_distance behind src/pre_process/pp_gen_nearest.py:131-204 (one caching_score call per query there).

    q [NQ,Sq,D], c [NC,Sc,D] fp32 CUDA (this rank's candidates, global ids base_id + row), lens int32.  Candidates go
    through the Q x C kernel (``asp_ot_score_allpairs``) ``chunk`` documents at a time; each chunk's [NQ, chunk]
    distances are reduced at once to per-query top-k lists of -distance (packed keys), the lists are merged on the
    device, and with an initialised process group the per-rank lists are all-gathered and merged (``gather_topk``).
    Returns (similarities [NQ,k] = -OT distance, ids int64 [NQ,k]); order: similarity descending, id ascending."""
    from .distances import ot_scores_allpairs
    NQ, NC = q.shape[0], c.shape[0]
    chunk = max(1, min(chunk, NC, (2 ** 31 - 1) // max(NQ, 1)))
    starts = list(range(0, NC, chunk))
    per_merge = max(1, 8192 // k)                       # lists one merge CTA takes
    lists = torch.empty((min(len(starts), per_merge), NQ, k), dtype=torch.int64, device=q.device)
    scores = torch.empty((NQ, chunk), dtype=torch.float32, device=q.device)
    n, best = 0, None
    for s0 in starts:
        m = min(chunk, NC - s0)
        sc = scores[:, :m] if m == chunk else torch.empty((NQ, m), dtype=torch.float32, device=q.device)
        ot_scores_allpairs(q, q_lens, c[s0:s0 + m], c_lens[s0:s0 + m], eps_list, temp=temp, out=sc)
        lists[n] = topk(sc, k, base_id=base_id + s0, negate=True, want_packed=True)[2]
        n += 1
        if n == lists.shape[0] and s0 != starts[-1]:    # buffer full: fold it into its first slot
            s_, i_ = topk_merge_packed(lists[:n].contiguous(), k)
            lists[0] = pack_keys(s_, i_)
            n = 1
    best = topk_merge_packed(lists[:n].contiguous(), k)
    return gather_topk(best[0], best[1], k, group=group)


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
