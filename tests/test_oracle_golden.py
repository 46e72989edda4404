"""CPU: the oracle (oracle/) must reproduce the golden vectors generated from the UNMODIFIED reference
(oracle/make_golden.py) bit for bit, and its own fp64 variant must agree with the fp32 path."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from oracle import aspire_ref as ar
from oracle import geomloss_ref as gr

from conftest import GOLDEN

OT_FILES = sorted(glob.glob(os.path.join(GOLDEN, "ot_*.npz")))


def _hp(z):
    hp = json.loads(str(z["hparams"]))
    return dict(blur=hp.get("geoml_blur", 0.05), scaling=hp.get("geoml_scaling", 0.9),
                temp=hp.get("sent_sm_temp", 1.0))


@pytest.mark.parametrize("fn", OT_FILES, ids=[os.path.basename(f) for f in OT_FILES])
def test_ot_restatement_bit_equal(fn):
    z = np.load(fn)
    q, c = torch.from_numpy(z["q"]), torch.from_numpy(z["c"])
    ql, cl = z["q_lens"].tolist(), z["c_lens"].tolist()
    dual = ar.ot_distance(q, ql, c, cl, **_hp(z))
    assert gr.SamplesLoss.last_call[1] == int(z["n_eps"])
    assert np.array_equal(dual.numpy(), z["dual"])
    primal, ex = ar.ot_distance(q, ql, c, cl, return_pair_sims=True, **_hp(z))
    assert np.array_equal(primal.numpy(), z["primal"])
    for t, k in zip(ex, ["alpha", "beta", "negc", "plan", "weighted"]):
        assert np.array_equal(t.numpy(), z[k]), k


@pytest.mark.parametrize("fn", OT_FILES, ids=[os.path.basename(f) for f in OT_FILES])
def test_l2max_restatement_bit_equal(fn):
    z = np.load(fn)
    q, c = torch.from_numpy(z["q"]), torch.from_numpy(z["c"])
    best, idx, sims = ar.l2max(q, z["q_lens"].tolist(), c, z["c_lens"].tolist())
    assert np.array_equal(best.numpy(), z["l2max_best"])
    assert np.array_equal(idx.numpy(), z["l2max_idx"])
    assert np.array_equal(sims.numpy(), z["l2max_sims"])
    assert np.array_equal(-best.numpy(), z["l2max_dist"])


@pytest.mark.parametrize("fn", OT_FILES, ids=[os.path.basename(f) for f in OT_FILES])
def test_fp64_solver_agrees(fn):
    """numpy fp64 solver on the exact cost vs the fp32 torch path: 1e-6 relative on the dual value."""
    z = np.load(fn)
    hp = _hp(z)
    C = torch.cdist(torch.from_numpy(z["q"]).double(), torch.from_numpy(z["c"]).double()).numpy()
    eps = gr.epsilon_schedule(1, float(z["diameter"]), hp["blur"], hp["scaling"])
    assert len(eps) == int(z["n_eps"])
    f, g, dual = gr.sinkhorn_np(z["alpha"].astype(np.float64), z["beta"].astype(np.float64), C, eps)
    rel = np.abs(dual - z["dual"]) / np.maximum(np.abs(dual), 1.0)
    assert rel.max() < 1e-6


def test_schedule_formula():
    eps = gr.epsilon_schedule(1, 43.12, 0.05, 0.9)
    assert eps[0] == 43.12 and eps[-1] == 0.05
    assert len(eps) == 2 + int(np.ceil(np.log(43.12 / 0.05) / np.log(1 / 0.9)))
    assert all(a >= b for a, b in zip(eps[:-1], eps[1:]))
    fx = gr.fixed_length_schedule(80.0, 0.01, 50)
    assert len(fx) == 50 and fx[0] == 80.0 and fx[-1] == 0.01


def test_get_similarity_golden():
    z = np.load(os.path.join(GOLDEN, "get_similarity_ragged.npz"))
    q, c = torch.from_numpy(z["q"]), torch.from_numpy(z["c"])
    for i, (ql, cl) in enumerate(zip(z["q_lens"], z["c_lens"])):
        d = ar.ot_distance(q[i:i + 1, :ql], [int(ql)], c[i:i + 1, :cl], [int(cl)])
        assert -float(d.item()) == z["sims"][i]


def test_caching_score_golden():
    z = np.load(os.path.join(GOLDEN, "caching_score.npz"))
    lens = z["c_lens"].tolist()
    offs = np.cumsum([0] + lens)
    cands = [z["c_cat"][offs[i]:offs[i + 1]] for i in range(len(lens))]
    s_ot, extras = ar.caching_score(z["q"], cands, "l2wasserstein")
    assert np.array_equal(s_ot, z["l2wasserstein_scores"])
    plan_cat = np.concatenate([extras[3][i, :z["q"].shape[0], :n].reshape(-1) for i, n in enumerate(lens)])
    assert np.array_equal(plan_cat, z["plan_cat"])
    s_ts, _ = ar.caching_score(z["q"], cands, "l2max")
    assert np.array_equal(s_ts, z["l2max_scores"])


def test_span_pool_golden():
    z = np.load(os.path.join(GOLDEN, "span_pool.npz"))
    spans = z["spans"]
    idxs = [[list(range(s, e)) for s, e in doc if s >= 0] for doc in spans]
    cls, reps = ar.span_mean_pool(torch.from_numpy(z["hidden"]), idxs, max_sents=3)
    assert np.array_equal(cls.numpy(), z["cls"])
    np.testing.assert_allclose(reps.numpy(), z["reps"], rtol=0, atol=2e-6)
    assert np.all(reps.numpy()[1, 1:] == 0) and np.all(reps.numpy()[2, 2] == 0)


def test_metrics_known_answers():
    with open(os.path.join(GOLDEN, "manifest.json")) as fh:
        m = json.load(fh)
    # the reference's only doctests: metrics.py:103-108 and :129-134
    assert abs(m["average_precision"]["value"] - 0.78333333333333333) < 1e-12
    assert abs(m["mean_average_precision"]["value"] - 0.39166666666666666) < 1e-12
    assert abs(ar.average_precision(m["average_precision"]["r"]) - m["average_precision"]["value"]) < 1e-12
    assert abs(ar.mean_average_precision(m["mean_average_precision"]["rs"]) -
               m["mean_average_precision"]["value"]) < 1e-12


def test_heads_oracle_vs_reference_golden():
    """l2top2 / attention restatements (oracle/aspire_ref.py) against the unmodified reference's outputs."""
    z = np.load(os.path.join(GOLDEN, "heads.npz"))
    q, c = torch.from_numpy(z["q"]), torch.from_numpy(z["c"])
    ql, cl = z["q_lens"].tolist(), z["c_lens"].tolist()
    top2, sims = ar.l2top2(q, ql, c, cl)
    np.testing.assert_allclose(top2.numpy(), z["top2_sims"], rtol=1e-6, atol=1e-5)
    np.testing.assert_allclose(sims.numpy(), z["top2_pair"], rtol=1e-6, atol=1e-5)
    assert z["top2_sims"][0] <= -9e8  # single valid pair: the runner-up is the mask constant
    for temp in (1.0, 0.25):
        doc, probs = ar.attention_sim(q, ql, c, cl, temp)
        np.testing.assert_allclose(doc.numpy(), z[f"att_sims_t{temp}"], rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(probs.numpy(), z[f"att_softmax_t{temp}"], rtol=1e-5, atol=1e-7)
