"""GPU: the batched evaluation callers against the reference-shaped per-pair / 64-chunk calls."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import ref_shims

pytestmark = pytest.mark.gpu


@pytest.fixture()
def patched_hf(monkeypatch):
    import transformers
    monkeypatch.setattr(transformers.AutoModel, "from_pretrained",
                        staticmethod(lambda name, *a, **k: ref_shims.seeded_bert(0, num_hidden_layers=2)))
    monkeypatch.setattr(transformers.AutoTokenizer, "from_pretrained",
                        staticmethod(lambda name, *a, **k: ref_shims.ToyTokenizer()))


def _dataset(root, n=24):
    rng = np.random.RandomState(1)
    words = ["graph", "neural", "protein", "retrieval", "sentence", "transport", "kernel", "citation", "facet", "model"]
    with open(os.path.join(root, "abstracts-toyds.jsonl"), "w") as fh:
        for i in range(n):
            sents = [" ".join(rng.choice(words, size=rng.randint(4, 12))) for _ in range(rng.randint(2, 8))]
            fh.write(json.dumps({"paper_id": str(i), "title": " ".join(rng.choice(words, size=5)), "abstract": sents}) + "\n")
    pool = {"0": {"cands": [str(i) for i in range(1, n)], "relevance_adju": rng.randint(0, 4, n - 1).tolist()},
            "1": {"cands": [str(i) for i in range(2, n)], "relevance_adju": rng.randint(0, 4, n - 2).tolist()}}
    with open(os.path.join(root, "test-pid2anns-toyds.json"), "w") as fh:
        json.dump(pool, fh)


def test_batched_score_matches_per_pair_reference_calls(tmp_path, patched_hf):
    from aspire_b200.evaluation import EvalDataset, score
    from aspire_b200.similarity import get_model
    _dataset(str(tmp_path))
    ds = EvalDataset("toyds", str(tmp_path))
    model = get_model("aspire_compsci")
    model.set_encodings_cache(None)
    batched = score(model, ds, None, str(tmp_path / "scores.json"), mode="batched")
    exact = score(model, ds, None, None, mode="reference")
    for q in exact:
        a, b = dict(batched[q]), dict(exact[q])
        assert set(a) == set(b)
        for c in a:  # one schedule per pool vs one per pair: <= 1e-4 relative
            assert abs(a[c] - b[c]) <= 1e-4 * max(abs(b[c]), 1.0)
        ra, rb = [c for c, _ in batched[q]], [c for c, _ in exact[q]]
        swaps = [i for i, (x, y) in enumerate(zip(ra, rb)) if x != y]
        for i in swaps:  # any disagreement in the order is between scores closer than the tolerance
            assert abs(b[ra[i]] - b[rb[i]]) <= 2e-4 * max(abs(b[rb[i]]), 1.0)


def test_caching_scoring_model_predict_and_rank(tmp_path, patched_hf):
    from aspire_b200.consent import AspireConSent
    from aspire_b200.evaluation import CachingScoringModel
    from aspire_b200.similarity import caching_score
    _dataset(str(tmp_path))
    pid2abstract = {}
    with open(os.path.join(str(tmp_path), "abstracts-toyds.jsonl")) as fh:
        for line in fh:
            r = json.loads(line)
            pid2abstract[r["paper_id"]] = r
    for agg in ("l2wasserstein", "l2max"):
        scorer = CachingScoringModel(AspireConSent("x"), ref_shims.ToyTokenizer(), score_agg_type=agg, score_batch_size=8)
        cands = [str(i) for i in range(1, 24)]
        out = scorer.predict("0", cands, pid2abstract)
        assert len(out["cand_scores"]) == len(cands) == len(out["pair_scores"])
        # same numbers as calling caching_score on the same 8-candidate chunks by hand
        q = scorer.pid2model_reps["0"]
        manual = []
        for s in range(0, len(cands), 8):
            manual += caching_score(q, [scorer.pid2model_reps[c] for c in cands[s:s + 8]], score_agg_type=agg)["batch_scores"].tolist()
        assert np.allclose(out["cand_scores"], manual, rtol=0, atol=0)
        ranked = scorer.rank_pool("0", cands, pid2abstract)
        assert [s for _, s in ranked] == sorted(out["cand_scores"], reverse=True)
        assert q["sent_reps"].shape == (len(pid2abstract["0"]["abstract"]), 768) and q["doc_cls_reps"].shape == (768,)


def test_rank_pool_sent_from_npy_reps_vs_float64_numpy_path(tmp_path):
    """pp_gen_nearest.rank_pool_sent (:863-985): sentence vectors in ``{ds}-sent.npy`` + ``pid2idx-{ds}-sent.json``;
    scores against the float64 ``-cdist`` + per-candidate max / top-2 restatement, ranking order and output file."""
    import json
    from aspire_b200.evaluation import rank_pool_sent
    rng = np.random.RandomState(5)
    ds, n_docs, D = "toyds", 40, 768
    root, reps = str(tmp_path), str(tmp_path / "myreps")
    os.makedirs(reps)
    lens = rng.randint(1, 9, n_docs)
    lens[1] = 1
    pid2idx, rows = {}, []
    with open(os.path.join(root, f"abstracts-{ds}.jsonl"), "w") as fh:
        for d in range(n_docs):
            pid = str(500 + d)
            fh.write(json.dumps({"paper_id": pid, "title": f"t{d}", "abstract": [f"s{j}" for j in range(lens[d])]}) + "\n")
            for j in range(lens[d]):
                pid2idx[f"{pid}-{j}"] = len(rows)
                rows.append(0.3 * rng.randn(D))
    allreps = np.asarray(rows, dtype=np.float32)
    allreps[3, 5] = np.nan  # read as 0 (pp_gen_nearest.py:905)
    np.save(os.path.join(reps, f"{ds}-sent.npy"), allreps)
    with open(os.path.join(reps, f"pid2idx-{ds}-sent.json"), "w") as fh:
        json.dump(pid2idx, fh)
    pool = {"500": {"cands": [str(500 + d) for d in range(1, n_docs)], "relevance_adju": [0] * (n_docs - 1)},
            "501": {"cands": [str(500 + d) for d in range(2, n_docs)], "relevance_adju": [0] * (n_docs - 2)}}
    with open(os.path.join(root, f"test-pid2anns-{ds}.json"), "w") as fh:
        json.dump(pool, fh)
    clean = np.nan_to_num(allreps).astype(np.float64)

    def doc(pid):
        return clean[[pid2idx[f"{pid}-{j}"] for j in range(lens[int(pid) - 500])]]
    from scipy.spatial.distance import cdist
    for score_type in ("l2max", "l2top2"):
        got = rank_pool_sent(root, reps, ds, score_type=score_type)
        with open(os.path.join(reps, f"test-pid2pool-{ds}-myreps-ranked.json")) as fh:
            on_disk = json.load(fh)
        for qpid, p in pool.items():
            want = {}
            for cpid in p["cands"]:
                sims = -cdist(doc(qpid), doc(cpid)).flatten()
                if score_type == "l2max":
                    want[cpid] = sims.max()
                else:  # np.partition(kth=2) needs three entries; fewer are summed as they are (:949-957)
                    want[cpid] = np.sort(sims)[::-1][:2].sum() if sims.size >= 3 else sims.sum()
            ranked = got[qpid]
            assert [c for c, _ in ranked] == [c for c, _ in on_disk[qpid]]
            assert sorted(c for c, _ in ranked) == sorted(p["cands"])
            for cpid, neg_sim in ranked:
                assert abs(-neg_sim - want[cpid]) <= 3e-5 * max(1.0, abs(want[cpid])), (score_type, qpid, cpid)
            vals = [v for _, v in ranked]
            assert vals == sorted(vals)
    # the all-queries x all-documents caller (pp_gen_nearest.py:729-860) on the tensor-core all-pairs kernel: documents
    # of 1..8 sentences (S < 10), a candidate listed twice is ranked once, same scores and order as the per-pool path
    from aspire_b200.evaluation import rank_pool_sent_treccovid
    pool["501"]["cands"].append("505")
    pool["501"]["relevance_adju"].append(0)
    with open(os.path.join(root, f"test-pid2anns-{ds}.json"), "w") as fh:
        json.dump(pool, fh)
    per_pool = rank_pool_sent(root, reps, ds, score_type="l2max", write=False)
    deep = rank_pool_sent_treccovid(root, reps, ds, score_type="l2max", cand_chunk=16)
    with open(os.path.join(reps, f"test-pid2pool-{ds}-myreps-ranked.json")) as fh:
        assert [c for c, _ in json.load(fh)["501"]] == [c for c, _ in deep["501"]]
    for qpid in pool:
        assert len(deep[qpid]) == len(set(pool[qpid]["cands"])) == len(per_pool[qpid])
        a, b = dict(per_pool[qpid]), dict(deep[qpid])
        assert max(abs(a[c] - b[c]) for c in a) <= 3e-5
        gap = np.diff([v for _, v in per_pool[qpid]]).min()
        if gap > 1e-4:
            assert [c for c, _ in per_pool[qpid]] == [c for c, _ in deep[qpid]]


def test_resident_corpus_indexed_pools_match_packed_pools():
    """Pools as index lists into a corpus kept in HBM (asp_ot_score_indexed): identical scores to packing each pool and
    calling the plain entry point with the same schedule; ragged documents, repeated and out-of-order candidates."""
    from aspire_b200 import ot_scores, epsilon_schedule
    from aspire_b200.similarity import ResidentCorpus, pack_pool
    rng = np.random.RandomState(11)
    D = 768
    pid2enc = {f"p{i}": (0.3 * rng.randn(rng.randint(1, 11), D)).astype(np.float32) for i in range(300)}
    corpus = ResidentCorpus(pid2enc)
    hp = {"geoml_diameter": 50.0}
    for qpid, n in (("p7", 250), ("p123", 1), ("p0", 1500)):
        cands = [f"p{j}" for j in rng.randint(0, 300, n)]
        got = corpus.score_pool(qpid, cands, model_hparams=hp)
        c, cl = pack_pool([pid2enc[p] for p in cands], corpus.device, max_sents=corpus.reps.shape[1])
        q, ql = pack_pool([pid2enc[qpid]], corpus.device, max_sents=corpus.reps.shape[1])
        eps = epsilon_schedule(50.0, 0.05, 0.9)
        want = -ot_scores(q, ql, c, cl, eps, broadcast_query=True)["dual"].cpu()
        assert torch.equal(got, want)
        ts = corpus.score_pool(qpid, cands, score_aggregation='l2max')
        assert ts.shape == (n,) and torch.isfinite(ts).all()
