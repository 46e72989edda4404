"""TEST INFRASTRUCTURE ONLY -- golden values for aspire_b200/metrics.py from the UNMODIFIED reference module
src/evaluation/utils/metrics.py (run in the build container; writes tests/golden/metrics.json).

The reference calls ``np.asfarray`` (removed in NumPy 2); the alias below restores exactly what NumPy 1.x did.
"""
import importlib.util
import json
import os

import numpy as np

REF = os.environ.get("ASPIRE_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "metrics.json")


def main():
    if not hasattr(np, "asfarray"):
        np.asfarray = lambda a, dtype=np.float64: np.asarray(a, dtype=dtype)
    spec = importlib.util.spec_from_file_location("ref_metrics", os.path.join(REF, "src/evaluation/utils/metrics.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = np.random.default_rng(20261017)
    cases = []
    for n, p_rel in ((5, 0.5), (30, 0.2), (100, 0.1), (100, 0.0), (250, 0.3), (60, 1.0)):
        graded = [int(g) for g in (rng.integers(1, 4, size=n) * (rng.random(n) < p_rel))]
        for thr in (1, 2):
            atks = [k for k in (5, 10, 20) if k <= n]
            cases.append({"graded": graded, "threshold": thr, "pr_atks": atks,
                          "metrics": ref.compute_metrics(graded, atks, thr),
                          "dcg0": float(ref.dcg_at_k(graded, 10, 0)), "dcg1": float(ref.dcg_at_k(graded, 10, 1)),
                          "ndcg1": float(ref.ndcg_at_k(graded, 10, 1)),
                          "mrr": float(ref.mean_reciprocal_rank([graded, graded[::-1]])),
                          "map": float(ref.mean_average_precision([graded, graded[::-1]]))})
    with open(OUT, "w") as fh:
        json.dump({"generator": "oracle/make_golden_metrics.py", "reference": "src/evaluation/utils/metrics.py",
                   "cases": cases}, fh)
    print("wrote", OUT, len(cases), "cases")


if __name__ == "__main__":
    main()
