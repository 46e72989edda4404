"""aspire_b200.metrics vs golden values produced by the unmodified reference module (oracle/make_golden_metrics.py)."""
import json
import os

import numpy as np
import pytest

from aspire_b200 import metrics as M
from conftest import GOLDEN


def test_metrics_match_reference_golden():
    with open(os.path.join(GOLDEN, "metrics.json")) as fh:
        cases = json.load(fh)["cases"]
    assert len(cases) >= 10
    for c in cases:
        got = M.compute_metrics(c["graded"], c["pr_atks"], c["threshold"])
        assert set(got) == set(c["metrics"])
        for k, v in c["metrics"].items():
            assert abs(got[k] - v) <= 1e-12, (k, got[k], v)
        g = c["graded"]
        assert abs(M.dcg_at_k(g, 10, 0) - c["dcg0"]) <= 1e-12
        assert abs(M.dcg_at_k(g, 10, 1) - c["dcg1"]) <= 1e-12
        assert abs(M.ndcg_at_k(g, 10, 1) - c["ndcg1"]) <= 1e-12
        assert abs(M.mean_reciprocal_rank([g, g[::-1]]) - c["mrr"]) <= 1e-12
        assert abs(M.mean_average_precision([g, g[::-1]]) - c["map"]) <= 1e-12


def test_docstring_values_and_errors():
    # the doctest values of the reference (metrics.py:103-108,129-134,60-62)
    r = [1, 1, 0, 1, 0, 1, 0, 0, 0, 1]
    assert abs(M.average_precision(r) - 0.78333333333333333) < 1e-15
    assert abs(M.mean_average_precision([r, [0]]) - 0.39166666666666666) < 1e-15
    assert M.average_precision([0, 0]) == 0.
    assert abs(M.precision_at_k([0, 0, 1], 3) - 1 / 3) < 1e-15
    with pytest.raises(ValueError):
        M.precision_at_k([0, 0, 1], 4)
    with pytest.raises(ValueError):
        M.dcg_at_k([1, 2], 2, method=2)
    assert M.ndcg_at_k([0], 1) == 0.
    assert M.recall_at_k([0, 0], 1, 0) == 0.0
    agg = M.aggregate_metrics([{"a": 1.0, "b": 2.0}, {"a": 3.0}])
    assert agg == {"a": 2.0, "b": 2.0}
