"""CPU: file formats and ranking order of the batched evaluation callers (aspire_b200/evaluation.py) against the
formats the reference reads/writes (src/evaluation/utils/datasets.py:24-94, evaluate.py:76-82)."""
import json
import os

import numpy as np
import pytest

from aspire_b200.evaluation import EvalDataset, rank_candidates, score
from aspire_b200.similarity import SimilarityModel


def _write_dataset(root, name="toyds", n=12):
    rng = np.random.RandomState(0)
    with open(os.path.join(root, f"abstracts-{name}.jsonl"), "w") as fh:
        for i in range(n):
            rec = {"paper_id": str(1000 + i), "title": f"title {i}", "abstract": [f"sent {i} {j}" for j in range(1 + i % 4)]}
            if i % 2 == 0:
                rec["pred_labels"] = ["background_label"] * (1 + i % 4)
            fh.write(json.dumps(rec) + "\n")
    pool = {"1000": {"cands": [str(1000 + i) for i in range(1, n)], "relevance_adju": rng.randint(0, 4, n - 1).tolist()},
            "1001": {"cands": [str(1000 + i) for i in range(2, n)], "relevance_adju": rng.randint(0, 4, n - 2).tolist()}}
    with open(os.path.join(root, f"test-pid2anns-{name}.json"), "w") as fh:
        json.dump(pool, fh)
    return name, pool


class _LenModel(SimilarityModel):
    """similarity = -|#sentences difference| -> many exact ties, which must keep pool order (stable sort)."""

    def encode(self, batch_papers):
        return [np.full((len(p["ABSTRACT"]), 4), float(len(p["ABSTRACT"])), dtype=np.float32) for p in batch_papers]

    def get_similarity(self, x, y):
        return -abs(float(len(x)) - float(len(y)))


def test_dataset_reader_and_scores_file(tmp_path):
    name, pool = _write_dataset(str(tmp_path))
    ds = EvalDataset(name, str(tmp_path))
    assert ds.get("1000") == {"TITLE": "title 0", "ABSTRACT": ["sent 0 0"], "FACETS": ["background_label"]}
    assert "FACETS" not in ds.get("1001")
    assert ds.get_test_pool() == pool
    assert ds.get_gold_test_data()["1000"]["1001"] == pool["1000"]["relevance_adju"][0]
    assert ds.get_threshold_grade() == 2 and EvalDataset.__new__(EvalDataset).__class__ is EvalDataset
    model = _LenModel(name="len", encoding_type="sentence")
    out = str(tmp_path / "scores.json")
    res = score(model, ds, None, out)
    with open(out) as fh:
        on_disk = json.load(fh)
    assert list(on_disk) == ["1000", "1001"]
    for q, ranked in on_disk.items():
        assert [c for c, _ in ranked] == [c for c, _ in res[q]]
        vals = [v for _, v in ranked]
        assert vals == sorted(vals)  # stored value = -similarity, best (smallest) first (evaluate.py:77)
        # exact ties keep the candidate-pool order
        for (c1, v1), (c2, v2) in zip(ranked, ranked[1:]):
            if v1 == v2:
                assert pool[q]["cands"].index(c1) < pool[q]["cands"].index(c2)


def test_rank_candidates_is_stable_descending():
    r = rank_candidates(["a", "b", "c", "d"], [1.0, 3.0, 1.0, 3.0])
    assert r == [("b", 3.0), ("d", 3.0), ("a", 1.0), ("c", 1.0)]


def test_evaluate_writes_reference_csvs(tmp_path):
    """score -> evaluate: per-query metrics equal compute_metrics on the gold relevances in rank order, the aggregate
    is the rounded mean, and the two CSV files carry the reference's names and columns (evaluate.py:118-157)."""
    import pandas as pd
    from aspire_b200.evaluation import evaluate, sorted_relevancies
    from aspire_b200 import metrics as M
    name, pool = _write_dataset(str(tmp_path), n=30)
    ds = EvalDataset(name, str(tmp_path))
    res = score(_LenModel(name="len", encoding_type="sentence"), ds, None, None)
    rels = sorted_relevancies(res, ds)
    gold = ds.get_gold_test_data()
    assert rels["1000"] == [gold["1000"][c] for c, _ in res["1000"]]
    out_dir = str(tmp_path / "results")
    per_query, agg = evaluate(res, ds, None, results_dir=out_dir)
    assert [m["paper_id"] for m in per_query] == ["1000", "1001"]
    want = M.compute_metrics(rels["1001"], [5, 10, 20], ds.get_threshold_grade())
    for k, v in want.items():
        assert per_query[1][k] == v
    assert per_query[0]["title"] == "title 0" and per_query[0]["facet"] == "unfaceted" and per_query[0]["split"] == "test"
    assert len(agg) == 1 and agg[0]["split"] == "test"
    assert agg[0]["av_precision"] == round((per_query[0]["av_precision"] + per_query[1]["av_precision"]) / 2, 4)
    q = pd.read_csv(os.path.join(out_dir, "query-evaluations.csv"))
    a = pd.read_csv(os.path.join(out_dir, "aggregated-evaluations.csv"))
    assert {"av_precision", "ndcg%20", "precision@5", "recall@20", "r_precision", "facet", "split", "paper_id", "title"} <= set(q.columns)
    assert len(q) == 2 and len(a) == 1 and abs(a["ndcg%20"][0] - agg[0]["ndcg%20"]) < 1e-12


def test_evaluate_reads_dev_test_split_and_aggregates_all_facets(tmp_path):
    """evaluate.py:103-154: {ds}-evaluation_splits.json assigns every query to dev or test (separate aggregate rows, in
    first-seen order); facet='all' takes one result set per facet and adds the per-split aggregate over all of them."""
    from aspire_b200.evaluation import evaluate
    name, pool = _write_dataset(str(tmp_path), n=30)
    with open(os.path.join(str(tmp_path), f"{name}-evaluation_splits.json"), "w") as fh:
        json.dump({"1000": "dev", "1001": "test"}, fh)
    with open(os.path.join(str(tmp_path), f"{name}-queries-release.csv"), "w") as fh:
        fh.write("pid,title\n1000,Query zero\n1001,Query one\n")
    ds = EvalDataset(name, str(tmp_path))
    assert ds.get_test_dev_split() == {"1000": "dev", "1001": "test"}
    res = score(_LenModel(name="len", encoding_type="sentence"), ds, None, None)
    per_query, agg = evaluate(res, ds, None)
    assert [m["split"] for m in per_query] == ["dev", "test"] and per_query[0]["title"] == "Query zero"
    assert [(a["facet"], a["split"]) for a in agg] == [("unfaceted", "dev"), ("unfaceted", "test")]
    assert agg[0]["av_precision"] == round(per_query[0]["av_precision"], 4)
    # facet = 'all': the same pool file serves as every facet's pool here
    for f in ("background", "method", "result"):
        with open(os.path.join(str(tmp_path), f"test-pid2anns-{name}-{f}.json"), "w") as fh:
            json.dump(pool, fh)
    per_query, agg = evaluate({f: res for f in ("background", "method", "result")}, ds, "all", results_dir=str(tmp_path / "r"))
    assert len(per_query) == 6
    assert [(a["facet"], a["split"]) for a in agg] == [("background", "dev"), ("background", "test"), ("method", "dev"),
                                                       ("method", "test"), ("result", "dev"), ("result", "test"),
                                                       ("all", "dev"), ("all", "test")]
    assert os.path.exists(os.path.join(str(tmp_path / "r"), "aggregated-evaluations-all.csv"))
    # csfcube has no split file by definition: every query is 'test'
    assert EvalDataset.get_test_dev_split(type("D", (), {"name": "csfcube", "root_path": str(tmp_path)})()) is None
