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
