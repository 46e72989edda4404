"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container (needs /root/reference):  ``python -m oracle.make_golden``

Every fixture stores the seeded inputs AND the outputs of the reference's own functions
(imported through ``oracle.ref_shims``; geomloss is the restatement in ``oracle/geomloss_ref.py`` -- the
Sinkhorn step is therefore "parity unpinned", everything around it is the reference's own code).
The GPU parity tests and the oracle self-check replay these files; nothing reads /root/reference at test time.
"""
import json
import os
import types

import numpy as np
import torch

from . import ref_shims

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

README_ABSTRACTS = [
    {"TITLE": "Multi-Vector Models with Textual Guidance for Fine-Grained Scientific Document Similarity",
     "ABSTRACT": ["We present a new scientific document similarity model based on matching fine-grained "
                  "aspects of texts.",
                  "To train our model, we exploit a naturally-occurring source of supervision: sentences in "
                  "the full-text of papers that cite multiple papers together (co-citations)."]},
    {"TITLE": "CSFCube -- A Test Collection of Computer Science Research Articles for Faceted Query by Example",
     "ABSTRACT": ["Query by Example is a well-known information retrieval task in which a document is chosen "
                  "by the user as the search query and the goal is to retrieve relevant documents from a "
                  "large collection.",
                  "However, a document often covers multiple aspects of a topic.",
                  "To address this scenario we introduce the task of faceted Query by Example in which users "
                  "can also specify a finer grained aspect in addition to the input query document. "]},
]


def _reps(gen, B, S, D, lens, structured=False):
    """Seeded sentence reps [B,S,D] with zero pad rows (what caching_score / forward produce)."""
    if structured:  # SURVEY 8d "structured" distribution: low-rank + noise -> wider distance spread
        W = torch.randn(32, D, generator=torch.Generator().manual_seed(99)) / np.sqrt(32) * 4
        x = (0.3 * torch.randn(B, S, 32, generator=gen)) @ W + 0.05 * torch.randn(B, S, D, generator=gen)
    else:
        x = 0.3 * torch.randn(B, S, D, generator=gen)
    for b, n in enumerate(lens):
        x[b, n:] = 0
    return x.contiguous()


def ot_cases():
    g = torch.Generator().manual_seed(1234)
    rng = np.random.default_rng(1234)
    cases = []
    # name, B, Sq, Sc, D, qlens, clens, hparams, structured
    cases.append(("ot_10x10_d768", 6, 10, 10, 768, [10] * 6, [10] * 6, {}, False))
    cases.append(("ot_struct_d768", 4, 10, 10, 768, [10] * 4, [10] * 4, {}, True))
    ql = [int(v) for v in rng.integers(1, 8, 12)]; ql[0] = 7; ql[1] = 1
    cl = [int(v) for v in rng.integers(1, 10, 12)]; cl[0] = 9; cl[2] = 1
    cases.append(("ot_ragged_d96", 12, 7, 9, 96, ql, cl, {}, False))
    cases.append(("ot_ragged_t05_b02", 12, 7, 9, 96, ql, cl, {"sent_sm_temp": 0.5, "geoml_blur": 0.2}, True))
    cases.append(("ot_ragged_t5000", 12, 7, 9, 96, ql, cl, {"sent_sm_temp": 5000.0}, False))
    ql30 = [30, 26, 2, 17]; cl30 = [30, 3, 29, 26]
    cases.append(("ot_upto30_d64", 4, 30, 30, 64, ql30, cl30, {"geoml_blur": 0.1}, False))
    cases.append(("ot_blur1_d64", 4, 30, 30, 64, ql30, cl30, {"geoml_blur": 1.0, "geoml_scaling": 0.8}, True))
    out = []
    for name, B, Sq, Sc, D, qlens, clens, hp, st in cases:
        q = _reps(g, B, Sq, D, qlens, st)
        c = _reps(g, B, Sc, D, clens, st)
        out.append((name, q, qlens, c, clens, hp))
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    ref_shims.install(bert_seed=0, bert_layers=2)
    from collections import namedtuple
    import geomloss  # the restatement, registered by the shim
    from src.learning.facetid_models import pair_distances as pd_ref
    from src.learning.facetid_models import disent_models as dm_ref
    import ex_aspire_consent as ex_ref
    import ex_aspire_consent_multimatch as exm_ref
    from src.evaluation.utils import models as ev_models
    from src.evaluation.utils import metrics as ev_metrics

    RL = namedtuple("RepLen", ["embed", "abs_lens"])
    manifest = {}

    # ---- (i)+(ii)+(iii): compute_distance both branches, l2max, on the same inputs --------------------
    for name, q, qlens, c, clens, hp in ot_cases():
        qt = RL(embed=q.permute(0, 2, 1), abs_lens=qlens)
        ct = RL(embed=c.permute(0, 2, 1), abs_lens=clens)
        solver = pd_ref.AllPairMaskedWasserstein(hp)
        dual = solver.compute_distance(query=qt, cand=ct, return_pair_sims=False)
        diam, n_eps = geomloss.SamplesLoss.last_call
        primal, (alpha, beta, negc, plan, weighted) = solver.compute_distance(query=qt, cand=ct,
                                                                               return_pair_sims=True)
        # release copy in examples/ must agree with the training-side copy
        dual_ex = exm_ref.AllPairMaskedWasserstein(hp).compute_distance(query=qt, cand=ct)
        assert torch.equal(dual, dual_ex)
        best, sims = pd_ref.allpair_masked_dist_l2max(query=qt, cand=ct, return_pair_sims=True)
        dist_pos = pd_ref.allpair_masked_dist_l2max(query=qt, cand=ct, return_pair_sims=False)
        B = q.shape[0]
        flat_idx = torch.max(sims.view(B, -1), dim=1)[1]  # pair_distances.py:176 (computed there, then dropped)
        # potentials, for a finer-grained check of the solver
        alpha_r, beta_r = alpha, beta
        f, g = geomloss.SamplesLoss("sinkhorn", p=1, blur=solver.geoml_blur, reach=None,
                                    scaling=solver.geoml_scaling, debias=False, potentials=True)(
            alpha_r, q.contiguous(), beta_r, c.contiguous())
        np.savez(os.path.join(OUT, name + ".npz"),
                 q=q.numpy(), c=c.numpy(), q_lens=np.array(qlens), c_lens=np.array(clens),
                 hparams=json.dumps(hp), diameter=np.float64(diam), n_eps=np.int64(n_eps),
                 dual=dual.numpy(), primal=primal.numpy(), alpha=alpha.numpy(), beta=beta.numpy(),
                 negc=negc.numpy(), plan=plan.numpy(), weighted=weighted.numpy(), f=f.numpy(), g=g.numpy(),
                 l2max_best=best.numpy(), l2max_idx=flat_idx.numpy(), l2max_sims=sims.numpy(),
                 l2max_dist=dist_pos.numpy())
        manifest[name] = {"B": B, "diameter": diam, "n_eps": n_eps, "hparams": hp}

    # ---- AspireModel.get_similarity (evaluate.py per-pair path) ---------------------------------------
    name, q, qlens, c, clens, hp = ot_cases()[2]
    fake = types.SimpleNamespace()
    sims = [ev_models.AspireModel.get_similarity(fake, q[i, :qlens[i]], c[i, :clens[i]]) for i in range(len(qlens))]
    np.savez(os.path.join(OUT, "get_similarity_ragged.npz"), q=q.numpy(), c=c.numpy(), q_lens=np.array(qlens),
             c_lens=np.array(clens), sims=np.array(sims, dtype=np.float64))

    # ---- caching_score (pp_gen_nearest batched path): 1 query x 9 ragged candidates -------------------
    g = torch.Generator().manual_seed(777)
    qrep = (0.3 * torch.randn(6, 128, generator=g)).numpy().astype(np.float64)
    clens = [3, 8, 1, 5, 8, 2, 7, 4, 6]
    creps = [(0.3 * torch.randn(n, 128, generator=g)).numpy().astype(np.float64) for n in clens]
    qd = {"sent_reps": qrep, "doc_cls_reps": np.zeros(128)}
    cds = [{"sent_reps": r, "doc_cls_reps": np.zeros(128)} for r in creps]
    saved = {"q": qrep, "c_lens": np.array(clens), "c_cat": np.concatenate(creps, 0)}
    for agg in ("l2wasserstein", "l2max"):
        fn = (pd_ref.AllPairMaskedWasserstein({}).compute_distance if agg == "l2wasserstein"
              else pd_ref.allpair_masked_dist_l2max)
        fake = types.SimpleNamespace(score_agg_type=agg, dist_function=fn, sent_loss_prop=1.0, abs_loss_prop=0.0)
        ret = dm_ref.WordSentAlignBiEnc.caching_score(fake, qd, cds)
        saved[agg + "_scores"] = np.asarray(ret["batch_scores"])
        if agg == "l2wasserstein":
            saved["plan_cat"] = np.concatenate([p[3].reshape(-1) for p in ret["pair_scores"]])
        else:
            saved["sims_cat"] = np.concatenate([p.reshape(-1) for p in ret["pair_scores"]])
    np.savez(os.path.join(OUT, "caching_score.npz"), **saved)

    # ---- span mean-pool (consent_reps_bert with a canned hidden state) --------------------------------
    g = torch.Generator().manual_seed(4242)
    B, L, D = 3, 37, 768
    hidden = torch.randn(B, L, D, generator=g)
    spans = [[(5, 12), (12, 13), (13, 30)], [(3, 36)], [(9, 20), (20, 35)]]
    idxs = [[list(range(s, e)) for s, e in doc] for doc in spans]
    fake = types.SimpleNamespace(bert_encoding_dim=D,
                                 bert_encoder=lambda *a, **k: types.SimpleNamespace(last_hidden_state=hidden))
    bb = {"tokid_tt": None, "seg_tt": None, "attnmask_tt": None, "seq_lens": [37, 36, 35]}
    cls, reps = ex_ref.AspireConSent.consent_reps_bert(fake, bert_batch=bb, batch_senttok_idxs=idxs,
                                                       num_sents=[3, 1, 2])
    span_arr = -np.ones((B, 3, 2), dtype=np.int32)
    for b, doc in enumerate(spans):
        for s, (st, en) in enumerate(doc):
            span_arr[b, s] = (st, en)
    np.savez(os.path.join(OUT, "span_pool.npz"), hidden=hidden.numpy(), spans=span_arr, cls=cls.numpy(),
             reps=reps.numpy())

    # ---- prepare_abstracts on the README example + a truncation case (ToyTokenizer) -------------------
    tok = ref_shims.ToyTokenizer()
    long_doc = {"TITLE": "a long paper " * 5,
                "ABSTRACT": [("sentence%d " % i) + "word " * 60 for i in range(12)]}
    batch = README_ABSTRACTS + [long_doc]
    bert_batch, abs_lens, sent_tok_idxs = ex_ref.prepare_abstracts(batch_abs=batch, pt_lm_tokenizer=tok)
    span_list = [[[s[0], s[-1] + 1] for s in doc] for doc in sent_tok_idxs]
    with open(os.path.join(OUT, "prepare_abstracts.json"), "w") as fh:
        json.dump({"abs_lens": abs_lens, "seq_lens": bert_batch["seq_lens"], "spans": span_list,
                   "tokid": bert_batch["tokid_tt"].tolist(), "seg": bert_batch["seg_tt"].tolist(),
                   "attn": bert_batch["attnmask_tt"].tolist(), "long_doc": long_doc}, fh)

    # ---- encoder end to end on the README example (2-layer seeded BERT; config 1 plumbing) ------------
    model = ex_ref.AspireConSent("allenai/aspire-contextualsentence-singlem-compsci")
    bb, al, sti = ex_ref.prepare_abstracts(batch_abs=README_ABSTRACTS, pt_lm_tokenizer=tok)
    with torch.no_grad():
        cls, reps = model.forward(bert_batch=bb, abs_lens=al, sent_tok_idxs=sti)
    qt = RL(embed=reps[0:1].permute(0, 2, 1), abs_lens=[al[0]])
    ct = RL(embed=reps[1:2].permute(0, 2, 1), abs_lens=[al[1]])
    ts_score, _ = pd_ref.allpair_masked_dist_l2max(query=qt, cand=ct, return_pair_sims=True)
    np.savez(os.path.join(OUT, "readme_encoder_2layer.npz"), cls=cls.numpy(), reps=reps.numpy(),
             abs_lens=np.array(al), ts_score=ts_score.numpy())

    # ---- rank metrics known answers (the reference's only doctests, metrics.py:103-108,129-134) --------
    r1 = [1, 1, 0, 1, 0, 1, 0, 0, 0, 1]
    rs = [[1, 1, 0, 1, 0, 1, 0, 0, 0, 1], [0]]
    manifest["average_precision"] = {"r": r1, "value": float(ev_metrics.average_precision(r1))}
    manifest["mean_average_precision"] = {"rs": rs, "value": float(ev_metrics.mean_average_precision(rs))}
    with open(os.path.join(OUT, "manifest.json"), "w") as fh:
        json.dump(manifest, fh, indent=1)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
