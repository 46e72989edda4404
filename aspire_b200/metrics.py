"""Ranking metrics of the evaluation path (host side; numpy, float64).

Same names, arguments, return values and error behaviour as the reference's ``src/evaluation/utils/metrics.py``
(``precision_at_k`` :64-95, ``average_precision`` :98-121, ``mean_average_precision`` :124-143, ``dcg_at_k``
:146-188, ``ndcg_at_k`` :191-225, ``recall_at_k`` :227-243, ``compute_metrics`` :244-281, ``r_precision`` :31-54,
``mean_reciprocal_rank`` :7-28), which ``evaluate.py`` applies to the ranked relevance lists that the scoring kernels
produce.  Written as prefix-sum formulas instead of per-rank loops: a ranked list of n judgements costs O(n).
"""
import numpy as np


def _binary(r):
    return np.asarray(r) != 0


def mean_reciprocal_rank(rs):
    """Mean over queries of 1 / (rank of the first relevant item); 0 for a query without one (metrics.py:7-28)."""
    vals = []
    for r in rs:
        hits = np.flatnonzero(np.asarray(r))
        vals.append(1.0 / (hits[0] + 1) if hits.size else 0.0)
    return np.mean(vals)


def r_precision(r):
    """Precision at the rank of the last relevant item (metrics.py:31-54)."""
    rel = _binary(r)
    hits = np.flatnonzero(rel)
    if not hits.size:
        return 0.
    return np.mean(rel[:hits[-1] + 1])


def precision_at_k(r, k):
    """Fraction of relevant items among the first k; ValueError when fewer than k were ranked (metrics.py:64-95)."""
    assert k >= 1
    head = _binary(r)[:k]
    if head.size != k:
        raise ValueError('Relevance score length < k')
    return np.mean(head)


def average_precision(r):
    """Mean of precision@rank over the ranks of the relevant items; 0 without relevant items (metrics.py:98-121)."""
    rel = _binary(r)
    if not rel.any():
        return 0.
    prec = np.cumsum(rel) / np.arange(1, rel.size + 1)
    return np.mean(prec[rel])


def mean_average_precision(rs):
    """Mean of average_precision over queries (metrics.py:124-143)."""
    return np.mean([average_precision(r) for r in rs])


def dcg_at_k(r, k, method=1):
    """Discounted cumulative gain of the first k graded judgements (metrics.py:146-188).

    method 0: weights 1, 1, 1/log2(3), ...; method 1: weights 1/log2(2), 1/log2(3), ...
    """
    gains = np.asarray(r, dtype=np.float64)[:k]
    if not gains.size:
        return 0.
    if method == 0:
        return gains[0] + np.sum(gains[1:] / np.log2(np.arange(2, gains.size + 1)))
    if method == 1:
        return np.sum(gains / np.log2(np.arange(2, gains.size + 2)))
    raise ValueError('method must be 0 or 1.')


def ndcg_at_k(r, k, method=0):
    """dcg_at_k normalised by the DCG of the ideal (descending) order; 0 when that is 0 (metrics.py:191-225)."""
    ideal = dcg_at_k(sorted(r, reverse=True), k, method)
    if not ideal:
        return 0.
    return dcg_at_k(r, k, method) / ideal


def recall_at_k(ranked_rel, atk, max_total_relevant):
    """Relevant items in the first atk over min(total relevant, max_total_relevant) (metrics.py:227-243)."""
    total = min(max_total_relevant, sum(ranked_rel))
    if total <= 0:
        return 0.0
    return float(sum(ranked_rel[:atk])) / total


def compute_metrics(ranked_judgements, pr_atks, threshold_grade):
    """All per-query metrics evaluate.py writes: graded NDCG (full list, @20, @50, @p% of the pool) and, after
    thresholding the grades at ``threshold_grade``, precision/recall/F1 at each of ``pr_atks``, R-precision,
    average precision and reciprocal rank (metrics.py:244-281; same dict keys)."""
    graded = list(ranked_judgements)
    binary = [1 if g >= threshold_grade else 0 for g in graded]
    n = len(binary)
    metrics = {}
    for pct in (5, 10, 15, 20, 25):
        metrics[f'ndcg%{pct}'] = float(ndcg_at_k(graded, int((pct / 100) * n)))
    n_rel = sum(binary)
    for atk in pr_atks:
        rec = recall_at_k(ranked_rel=binary, atk=atk, max_total_relevant=n_rel)
        prec = precision_at_k(r=binary, k=atk)
        f1 = 2 * prec * rec / (prec + rec) if (prec + rec) > 0 else 0.0
        metrics[f'precision@{atk}'] = float(prec)
        metrics[f'recall@{atk}'] = float(rec)
        metrics[f'f1@{atk}'] = float(f1)
    metrics['r_precision'] = float(r_precision(r=binary))
    metrics['av_precision'] = float(average_precision(r=binary))
    metrics['reciprocal_rank'] = float(mean_reciprocal_rank(rs=[binary]))
    metrics['ndcg'] = float(ndcg_at_k(graded, n))
    metrics['ndcg@20'] = float(ndcg_at_k(graded, 20))
    metrics['ndcg@50'] = float(ndcg_at_k(graded, 50))
    return metrics


def aggregate_metrics(query_metrics):
    """Mean of every metric over queries (evaluate.py:97-104 aggregates the per-query CSV the same way)."""
    keys = sorted({k for m in query_metrics for k in m})
    return {k: float(np.mean([m[k] for m in query_metrics if k in m])) for k in keys}
