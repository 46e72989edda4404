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
