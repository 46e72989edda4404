"""GPU parity for the remaining kernels: span mean-pool (K1), top-k / merge (K5), bbox diameter, and the
batched callers (caching_score, score_pool) against golden vectors from the unmodified reference."""
import os

import numpy as np
import pytest
import torch

from oracle import aspire_ref as ar
from oracle import geomloss_ref as gr

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def test_span_pool_vs_golden():
    from aspire_b200.consent import span_mean_pool
    z = np.load(os.path.join(GOLDEN, "span_pool.npz"))
    cls, reps = span_mean_pool(torch.from_numpy(z["hidden"]).cuda(), torch.from_numpy(z["spans"]).cuda())
    assert np.array_equal(cls.cpu().numpy(), z["cls"])
    np.testing.assert_allclose(reps.cpu().numpy(), z["reps"], rtol=0, atol=2e-6)
    r = reps.cpu().numpy()
    assert np.all(r[1, 1:] == 0) and np.all(r[2, 2] == 0)  # missing sentences are exact zero rows


def test_span_pool_large_random_vs_oracle():
    from aspire_b200.consent import span_mean_pool, spans_from_token_idxs
    g = torch.Generator().manual_seed(7)
    B, L, D = 32, 502, 768
    hidden = torch.randn(B, L, D, generator=g)
    idxs = []
    for b in range(B):
        cuts = sorted(set(torch.randint(5, L - 1, (int(torch.randint(1, 21, (1,), generator=g)),),
                                        generator=g).tolist()))
        cuts = [4] + cuts
        idxs.append([list(range(s, e)) for s, e in zip(cuts[:-1], cuts[1:])] or [[4, 5]])
    smax = max(len(d) for d in idxs)
    cls, reps = span_mean_pool(hidden.cuda(), spans_from_token_idxs(idxs, smax).cuda())
    rc, rr = ar.span_mean_pool(hidden, idxs, max_sents=smax)
    assert torch.equal(cls.cpu(), rc)
    np.testing.assert_allclose(reps.cpu().numpy(), rr.numpy(), rtol=0, atol=5e-6)


def test_span_pool_bulk_copy_staging_is_bit_identical_to_streaming_loads():
    """asp_span_mean_pool with the spans staged by cp.async.bulk through a shared-memory ring (default) against the
    streaming-load kernel: same summation order, so bit-identical -- empty spans, one-token spans, spans longer than the
    ring (17+ rows), spans clipped at L, D = 768 and a small D."""
    from aspire_b200 import _abi
    from aspire_b200.consent import span_mean_pool
    g = torch.Generator().manual_seed(11)
    for B, L, D, S in ((7, 502, 768, 12), (3, 64, 128, 5)):
        hidden = torch.randn(B, L, D, generator=g).cuda()
        spans = torch.zeros(B, S, 2, dtype=torch.int32)
        for b in range(B):
            for s_ in range(S):
                a = int(torch.randint(0, L, (1,), generator=g))
                n = int(torch.randint(0, 60, (1,), generator=g)) if s_ % 3 else int(torch.randint(0, 3, (1,), generator=g))
                spans[b, s_] = torch.tensor([a, a + n])          # may run past L: clipped
        spans[0, 0] = torch.tensor([-1, -1])                      # missing sentence
        out = {}
        for mode in (1, 0):
            _abi.set_option("span_tma", mode)
            try:
                out[mode] = [t.clone() for t in span_mean_pool(hidden, spans.cuda())]
            finally:
                _abi.set_option("span_tma", 1)
        assert torch.equal(out[1][0], out[0][0]) and torch.equal(out[1][1], out[0][1])
        assert (out[1][1][0, 0] == 0).all()


def test_bbox_diameter():
    from aspire_b200 import bbox_diameter
    g = torch.Generator().manual_seed(9)
    for nx, ny, D in [(10, 10000, 768), (3, 5, 64), (20000, 1, 128)]:
        x, y = torch.randn(nx, D, generator=g), 2 * torch.randn(ny, D, generator=g) + 0.5
        d = bbox_diameter(x.cuda(), y.cuda())
        assert abs(d - gr.max_diameter(x, y)) <= 1e-5 * d


@pytest.mark.parametrize("Q,N,k", [(5, 1000, 100), (3, 70000, 100), (2, 50, 100), (4, 4096, 1), (1, 300000, 1024)])
def test_topk_order_and_ties(Q, N, k):
    from aspire_b200.ranking import topk, host_merge
    g = torch.Generator().manual_seed(N + k)
    s = torch.randn(Q, N, generator=g)
    s[:, : N // 3] = torch.round(s[:, : N // 3] * 4) / 4  # many exact ties
    s[0, 1] = float("-inf")
    ts, ti = topk(s.cuda(), k, base_id=1000)
    ids = (torch.arange(N) + 1000).unsqueeze(0).expand(Q, -1)
    ws, wi = host_merge(s, ids, k)
    assert torch.equal(ti.cpu(), wi)
    assert torch.equal(ts.cpu(), ws)


@pytest.mark.parametrize("Q,N,k", [(3, 8192 * 70 + 17, 128), (2, 700000, 100), (5, 9000, 7), (2, 100, 128)])
def test_topk_single_pass_negate_packed_and_multilevel_merge(Q, N, k):
    """asp_topk_ws: ranking by -scores inside the kernel, packed keys equal to the host restatement, and more chunk lists
    per query than one merge CTA holds (8192 / k) -> a second merge level."""
    from aspire_b200.ranking import topk, host_merge, pack_keys, unpack_keys
    g = torch.Generator().manual_seed(N + k)
    s = torch.randn(Q, N, generator=g)
    s[:, ::5] = torch.round(s[:, ::5] * 2) / 2           # exact ties across chunks
    s[0, 3] = 0.0
    s[0, 4] = -0.0
    ts, ti, tp = topk(s.cuda(), k, base_id=77, negate=True, want_packed=True)
    ids = (torch.arange(N) + 77).unsqueeze(0).expand(Q, -1)
    ws, wi = host_merge(-s, ids, k)
    assert torch.equal(ti.cpu(), wi)
    assert torch.equal(ts.cpu(), ws)
    assert torch.equal(tp.cpu(), pack_keys(ws, wi))
    us, ui = unpack_keys(tp.cpu())
    assert torch.equal(us, ws) and torch.equal(ui, wi)


def test_topk_merge_packed_in_gathered_layout():
    """asp_topk_merge_packed reads the [R, Q, k] buffer an all_gather_into_tensor fills, in place."""
    from aspire_b200.ranking import topk, topk_merge_packed, host_merge, shard_bounds
    g = torch.Generator().manual_seed(4)
    Q, N, k, R = 33, 20011, 100, 8
    s = torch.round(torch.randn(Q, N, generator=g) * 16) / 16
    sd = s.cuda()
    gathered = torch.empty(R, Q, k, dtype=torch.int64, device="cuda")
    for r in range(R):
        lo, hi = shard_bounds(N, R, r)
        gathered[r] = topk(sd[:, lo:hi].contiguous(), k, base_id=lo, want_packed=True)[2]
    ms, mi = topk_merge_packed(gathered, k)
    ws, wi = host_merge(s, torch.arange(N).unsqueeze(0).expand(Q, -1), k)
    assert torch.equal(mi.cpu(), wi) and torch.equal(ms.cpu(), ws)


def test_rank_corpus_ot_chunked_equals_one_shot():
    """rank_corpus_ot (candidate chunks -> per-chunk packed top-k -> device merge, buffer folding when there are more
    chunks than one merge takes) == top-k of the negated [NQ, NC] all-pairs matrix computed in one call."""
    from aspire_b200 import ot_scores_allpairs, epsilon_schedule
    from aspire_b200.ranking import rank_corpus_ot, topk
    g = torch.Generator().manual_seed(12)
    NQ, NC, S, D, k = 30, 3001, 10, 128, 100
    q = (0.3 * torch.randn(NQ, S, D, generator=g)).cuda()
    c = (0.3 * torch.randn(NC, S, D, generator=g)).cuda()
    ql = torch.randint(1, S + 1, (NQ,), generator=g).int().cuda()
    cl = torch.randint(1, S + 1, (NC,), generator=g).int().cuda()
    eps = epsilon_schedule(20.0, 0.05, 0.9)
    want_s, want_i = topk(ot_scores_allpairs(q, ql, c, cl, eps), k, base_id=500, negate=True)
    for chunk in (1000, 32):      # 4 chunks; 94 chunks (> 81 lists per merge: the buffer is folded once)
        got_s, got_i = rank_corpus_ot(q, ql, c, cl, eps, k, base_id=500, chunk=chunk)
        assert torch.equal(got_i, want_i) and torch.equal(got_s, want_s)


def test_topk_merge_vs_host():
    from aspire_b200.ranking import topk_merge, host_merge
    g = torch.Generator().manual_seed(1)
    Q, R, k = 9, 8, 100
    s = torch.round(torch.randn(Q, R * k, generator=g) * 8) / 8
    i = torch.stack([torch.randperm(10 ** 6, generator=g)[: R * k] for _ in range(Q)])
    i[:, -5:] = -1  # fillers
    ms, mi = topk_merge(s.cuda(), i.cuda(), k)
    ws, wi = host_merge(s, i, k)
    assert torch.equal(mi.cpu(), wi) and torch.equal(ms.cpu(), ws)


def test_caching_score_vs_golden():
    """pp_gen_nearest batched path (disent_models.py:256-342): 1 query x 9 ragged candidates."""
    from aspire_b200.similarity import caching_score
    z = np.load(os.path.join(GOLDEN, "caching_score.npz"))
    lens = z["c_lens"].tolist()
    offs = np.cumsum([0] + lens)
    qd = {"sent_reps": z["q"], "doc_cls_reps": np.zeros(128)}
    cds = [{"sent_reps": z["c_cat"][offs[i]:offs[i + 1]], "doc_cls_reps": np.zeros(128)} for i in range(len(lens))]
    ret = caching_score(qd, cds, "l2wasserstein")
    ref = z["l2wasserstein_scores"]
    assert np.max(np.abs(ret["batch_scores"] - ref) / np.maximum(np.abs(ref), 1)) <= 3e-4
    plan_cat = np.concatenate([p[3].reshape(-1) for p in ret["pair_scores"]])
    assert plan_cat.shape == z["plan_cat"].shape
    assert np.max(np.abs(plan_cat - z["plan_cat"])) <= 3e-4 * z["plan_cat"].max()
    assert len(ret["pair_scores"][2]) == 5 and ret["pair_scores"][2][3].shape == (6, 1)
    ret = caching_score(qd, cds, "l2max")
    np.testing.assert_allclose(ret["batch_scores"], z["l2max_scores"], rtol=1e-5, atol=2e-5)
    sims_cat = np.concatenate([p.reshape(-1) for p in ret["pair_scores"]])
    np.testing.assert_allclose(sims_cat, z["sims_cat"], rtol=1e-5, atol=2e-5)
    with pytest.raises(ValueError):
        caching_score(qd, cds, "nope")


def test_caching_score_mixing_vs_reference_golden():
    """disent_models.py:298-307: sent_loss_prop / sentsup_loss_prop scaling and the abs_loss_prop CLS-distance term,
    for all four aggregation heads, against tests/golden/caching_score_mix.npz (written by oracle/make_golden_mix.py
    from the UNMODIFIED reference method)."""
    from aspire_b200.similarity import caching_score
    z = np.load(os.path.join(GOLDEN, "caching_score_mix.npz"))
    lens = z["c_lens"].tolist()
    offs = np.cumsum([0] + lens)
    qd = {"sent_reps": z["q"], "doc_cls_reps": z["q_cls"]}
    cds = [{"sent_reps": z["c_cat"][offs[i]:offs[i + 1]], "doc_cls_reps": z["c_cls"][i]} for i in range(len(lens))]
    for agg in ("l2wasserstein", "l2max", "l2top2", "l2attention"):
        for tag in "abc":
            sp, ssp, ap = z["mix_" + tag].tolist()
            hp = {"sent_loss_prop": sp, "abs_loss_prop": ap}
            if ssp >= 0:
                hp["sentsup_loss_prop"] = ssp
            got = caching_score(qd, cds, agg, model_hparams=hp)["batch_scores"]
            ref = z[f"{agg}_{tag}"]
            tol = 3e-4 if agg == "l2wasserstein" else 2e-5
            assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1)) <= tol, (agg, tag)
    # explicit arguments win over the hparams
    got = caching_score(qd, cds, "l2max", model_hparams={"abs_loss_prop": 9.0}, sent_loss_prop=0.6, abs_loss_prop=0.4)
    np.testing.assert_allclose(got["batch_scores"], z["l2max_a"], rtol=1e-5, atol=2e-5)


def test_ranking_parity_map():
    """SURVEY 8d ranking-parity set (scaled): rank by kernel vs oracle; MAP within 1e-3 (here: equal rankings
    up to near-ties), through topk on the device."""
    from aspire_b200 import ot_scores, epsilon_schedule
    from aspire_b200.ranking import topk
    g = torch.Generator().manual_seed(5678)
    NQ, NC, S, D = 6, 400, 10, 768
    Q = 0.3 * torch.randn(NQ, S, D, generator=g)
    C = 0.3 * torch.randn(NC, S, D, generator=g)
    rel = torch.zeros(NQ, NC, dtype=torch.bool)
    for qi in range(NQ):
        for cj in torch.randperm(NC, generator=g)[:40].tolist():
            rel[qi, cj] = True
    # relevant candidates copy 1..5 query sentences (+noise) into random slots -- per query pools
    maps_k, maps_o = [], []
    eps = epsilon_schedule(60.0, 0.05, 0.9)
    for qi in range(NQ):
        pool = C.clone()
        for cj in torch.nonzero(rel[qi]).view(-1).tolist():
            n = int(torch.randint(1, 6, (1,), generator=g))
            slots = torch.randperm(S, generator=g)[:n]
            src = torch.randperm(S, generator=g)[:n]
            pool[cj, slots] = Q[qi, src] + 0.1 * torch.randn(n, D, generator=g)
        res = ot_scores(Q[qi:qi + 1].cuda(), torch.tensor([S]).int().cuda(), pool.cuda(),
                        torch.full((NC,), S).int().cuda(), eps, broadcast_query=True)
        sims = -res["dual"]
        _, order = topk(sims[None], NC)
        ref = -ar.ot_distance(Q[qi:qi + 1].expand(NC, -1, -1), [S] * NC, pool, [S] * NC, diameter=60.0)
        order_ref = sorted(range(NC), key=lambda j: -ref[j].item())
        maps_k.append(ar.average_precision([int(rel[qi, j]) for j in order[0].cpu().tolist()]))
        maps_o.append(ar.average_precision([int(rel[qi, j]) for j in order_ref]))
    assert abs(np.mean(maps_k) - np.mean(maps_o)) <= 1e-3
    assert np.mean(maps_k) > 0.5  # the synthetic relevance signal is actually recovered


def test_ranking_parity_set_at_spec_size_ot_and_ts():
    """SURVEY 8d ranking-parity set at its stated size: 50 queries x 1000 candidates, 100 relevant candidates per query
    (each copies k in 1..5 query sentence vectors + 0.1 randn into random slots), seed 5678.  Ranked by the kernels and
    by the oracle, for BOTH otAspire (-OT_eps) and tsAspire (max -dist): mean average precision within 1e-3; the
    tsAspire flat argmax equals the float64 argmax except where the float64 top-2 gap is below 1e-5."""
    from aspire_b200 import ot_scores, epsilon_schedule
    from aspire_b200.distances import l2max_scores
    g = torch.Generator().manual_seed(5678)
    NQ, NC, NREL, S, D = 50, 1000, 100, 10, 768
    Q = 0.3 * torch.randn(NQ, S, D, generator=g)
    C = 0.3 * torch.randn(NC, S, D, generator=g)
    eps = epsilon_schedule(60.0, 0.05, 0.9)
    lens_c = torch.full((NC,), S).int().cuda()
    ap = {k: [] for k in ("ot_kernel", "ot_oracle", "ts_kernel", "ts_oracle")}
    argmax_checked = argmax_ties = 0
    for qi in range(NQ):
        rel = torch.zeros(NC, dtype=torch.bool)
        rel[torch.randperm(NC, generator=g)[:NREL]] = True
        pool = C.clone()
        for cj in torch.nonzero(rel).view(-1).tolist():
            n = int(torch.randint(1, 6, (1,), generator=g))
            pool[cj, torch.randperm(S, generator=g)[:n]] = Q[qi, torch.randperm(S, generator=g)[:n]] + \
                0.1 * torch.randn(n, D, generator=g)
        qd, pd = Q[qi:qi + 1].cuda(), pool.cuda()
        ql1 = torch.tensor([S]).int().cuda()
        ot_k = -ot_scores(qd, ql1, pd, lens_c, eps, broadcast_query=True)["dual"].cpu().numpy()
        ot_o = -ar.ot_distance(Q[qi:qi + 1].expand(NC, -1, -1), [S] * NC, pool, [S] * NC, diameter=60.0).numpy()
        ts_k, idx_k, _ = l2max_scores(qd, ql1, pd, lens_c, broadcast_query=True)
        d64 = torch.cdist(Q[qi:qi + 1].double().expand(NC, -1, -1), pool.double())      # [NC, S, S]
        flat = (-d64).flatten(1)
        top2 = torch.topk(flat, 2, dim=1)
        ts_o = top2.values[:, 0].numpy()
        relv = rel.numpy().astype(int)
        for name, sc in (("ot_kernel", ot_k), ("ot_oracle", ot_o), ("ts_kernel", ts_k.cpu().numpy()), ("ts_oracle", ts_o)):
            order = np.argsort(-sc.astype(np.float64), kind="stable")
            ap[name].append(ar.average_precision(relv[order].tolist()))
        assert np.abs(ts_k.cpu().numpy() - ts_o).max() <= 2e-5
        clear = (top2.values[:, 0] - top2.values[:, 1]) > 1e-5
        argmax_checked += int(clear.sum())
        argmax_ties += int((~clear).sum())
        assert torch.equal(idx_k.cpu().long()[clear], top2.indices[:, 0][clear])
    m = {k: float(np.mean(v)) for k, v in ap.items()}
    assert abs(m["ot_kernel"] - m["ot_oracle"]) <= 1e-3, m
    assert abs(m["ts_kernel"] - m["ts_oracle"]) <= 1e-3, m
    assert m["ot_kernel"] > 0.5 and m["ts_kernel"] > 0.5, m   # the planted relevance signal is recovered
    assert argmax_checked >= 0.99 * NQ * NC, (argmax_checked, argmax_ties)


def test_topk_shard_invariance_on_one_gpu():
    """Sharded == unsharded ON THE DEVICE: a [Q, N] score matrix is cut into 8 contiguous candidate shards (uneven sizes),
    each shard's top-k taken with its base id (what every rank does), the 8 partial lists merged by ``asp_topk_merge``
    -- the result must equal the top-k of the whole matrix bit for bit, duplicated scores included (order: score
    descending, global id ascending)."""
    from aspire_b200.ranking import topk, topk_merge, shard_bounds
    g = torch.Generator(device="cuda").manual_seed(9)
    Q, N, k, W = 64, 30011, 100, 8
    scores = torch.randn(Q, N, device="cuda", generator=g)
    scores[:, ::7] = scores[:, 3:4]            # heavy exact ties across shards
    scores[5] = 1.25                           # a fully tied row
    whole_s, whole_i = topk(scores, k)
    parts_s, parts_i = [], []
    for r in range(W):
        lo, hi = shard_bounds(N, W, r)
        s, i = topk(scores[:, lo:hi].contiguous(), k, base_id=lo)
        parts_s.append(s)
        parts_i.append(i)
    ms, mi = topk_merge(torch.cat(parts_s, dim=1), torch.cat(parts_i, dim=1), k)
    assert torch.equal(mi, whole_i) and torch.equal(ms, whole_s)
    assert torch.equal(whole_i[5].cpu(), torch.arange(k, dtype=whole_i.dtype))


def test_score_pools_host_matches_device_path_and_oracle():
    """Host-buffer API (chunked, overlapped H2D): same scores as the resident-tensor path; ragged last chunk."""
    from aspire_b200 import epsilon_schedule, ot_scores
    from aspire_b200.similarity import score_pools_host
    g = torch.Generator().manual_seed(31)
    NQ, G = 5, 40
    NP = NQ * G - 7
    q = (0.3 * torch.randn(NQ, 10, 768, generator=g)).pin_memory()
    c = (0.3 * torch.randn(NP, 10, 768, generator=g)).pin_memory()
    ql = torch.randint(2, 11, (NQ,), generator=g).int()
    cl = torch.randint(1, 11, (NP,), generator=g).int()
    res = score_pools_host(q, ql, c, cl, G, diameter=60.0, chunk_queries=2)
    dev = ot_scores(q.cuda(), ql.cuda(), c.cuda(), cl.cuda(), epsilon_schedule(60.0, 0.05, 0.9), q_group=G)["dual"]
    assert torch.equal(res["scores"], -dev.cpu())
    assert res["scores"].is_pinned() and not res["scores"].is_cuda
    ts = score_pools_host(q, ql, c, cl, G, score_aggregation='l2max', chunk_queries=2)["scores"]
    qi = torch.arange(NP) // G
    best, _, _ = ar.l2max(q[qi], ql[qi].tolist(), c, cl.tolist())
    assert np.allclose(ts.numpy(), best.numpy(), rtol=1e-5, atol=2e-5)
