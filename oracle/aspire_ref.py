"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch/numpy) of Aspire's pair-scoring hot path.

Each function cites the reference lines it follows (paths relative to /root/reference).
It is validated against the unmodified reference, imported through ``ref_shims`` in the
build container, by ``oracle/make_golden.py`` (the outputs are committed under
``tests/golden``).  The Sinkhorn step itself is third-party (geomloss 0.2.4) and restated in
``geomloss_ref.py`` -- PARITY UNPINNED for that step, see the header there.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (cpu_baseline /
``--impl reference``) may import this module; the product never does.
"""
from collections import namedtuple

import numpy as np
import torch

from . import geomloss_ref

RepLen = namedtuple("RepLen", ["embed", "abs_lens"])
PAD_NEG = -10e8  # pair_distances.py:39 (== -1e9, exact in fp32)


def _pad_mask(q_lens, c_lens, qmax, cmax):
    """pair_distances.py:39-43 -- 0 inside [:ql,:cl], -1e9 elsewhere; fp32 [B,qmax,cmax]."""
    ql = torch.as_tensor(q_lens).view(-1, 1, 1)
    cl = torch.as_tensor(c_lens).view(-1, 1, 1)
    inside = (torch.arange(qmax).view(1, -1, 1) < ql) & (torch.arange(cmax).view(1, 1, -1) < cl)
    return torch.where(inside, 0.0, PAD_NEG).float()


def neg_pair_dists(q, c):
    """pair_distances.py:49-50 -- -cdist on [B,S,D] inputs (torch.cdist, fp32)."""
    return -1 * torch.cdist(q.contiguous(), c.contiguous())


def marginals(neg_c_masked, temp):
    """pair_distances.py:57-60 -- softmax over sentences of the best match per sentence."""
    q_best = neg_c_masked.max(dim=2)[0]
    c_best = neg_c_masked.max(dim=1)[0]
    alpha = torch.log_softmax(q_best / temp, dim=1).exp()
    beta = torch.log_softmax(c_best / temp, dim=1).exp()
    return alpha, beta


def ot_distance(q, q_lens, c, c_lens, blur=0.05, scaling=0.9, temp=1.0, return_pair_sims=False,
                diameter=None, eps_list=None):
    """AllPairMaskedWasserstein.compute_distance, pair_distances.py:21-92.

    q, c: fp32 [B, S, D] (the reference's ``embed`` permuted back, :49,71).  Returns the dual value
    OT_eps [B] (False branch, :87-92) or (primal sum P*(-C) [B], [alpha, beta, negC, P, P*negC])
    (True branch, :61-86).  ``diameter`` / ``eps_list`` are oracle-side extensions to pin the schedule.
    """
    B, qmax, _ = q.shape
    cmax = c.shape[1]
    mask = _pad_mask(q_lens, c_lens, qmax, cmax)
    negc = neg_pair_dists(q, c) + mask
    alpha, beta = marginals(negc, temp)
    solver = geomloss_ref.SamplesLoss("sinkhorn", p=1, blur=blur, reach=None, scaling=scaling, debias=False,
                                      potentials=return_pair_sims, diameter=diameter, eps_list=eps_list)
    if not return_pair_sims:
        return solver(alpha, q.contiguous(), beta, c.contiguous())
    keep = (mask == 0).float()                      # :64-65
    negc = negc * keep                              # :66 pads -> 0
    f, g = solver(alpha, q.contiguous(), beta, c.contiguous())
    outer = (f.unsqueeze(2) + g.unsqueeze(1)) * keep  # :76-79
    plan = torch.exp((outer + negc) / blur) * (alpha.unsqueeze(2) * beta.unsqueeze(1))  # :80-82
    weighted = plan * negc                          # :84
    return weighted.sum(dim=1).sum(dim=1), [alpha, beta, negc, plan, weighted]  # :85 (rows first, then columns)


def l2max(q, q_lens, c, c_lens):
    """allpair_masked_dist_l2max, pair_distances.py:138-186 (tsAspire).

    Returns (best similarity [B] = max -dist, flat argmax [B] = i*cmax+j first occurrence (:176),
    masked pair sims [B,qmax,cmax]).
    """
    B, qmax, _ = q.shape
    cmax = c.shape[1]
    negc = -1 * torch.cdist(q, c) + _pad_mask(q_lens, c_lens, qmax, cmax)
    best, idx = torch.max(negc.view(B, qmax * cmax), dim=1)
    return best, idx, negc


def l2max_np64(q_sents, pool_sents, pool_lens):
    """rank_pool_sent numpy path, src/pre_process/pp_gen_nearest.py:942-961 -- float64 -cdist, per-cand max."""
    from scipy.spatial.distance import cdist
    sims = -cdist(np.asarray(q_sents, np.float64), np.asarray(pool_sents, np.float64))
    out, start = [], 0
    for n in pool_lens:
        out.append(sims[:, start:start + n].max())
        start += n
    return np.array(out)


def span_mean_pool(hidden, sent_tok_idxs, max_sents=None):
    """consent_reps_bert pooling, examples/ex_aspire_consent.py:75-100.

    hidden fp32 [B,L,D]; sent_tok_idxs list[B][S][tokens].  Returns (cls [B,D], sent_reps [B,Smax,D]):
    sum over the sentence's token rows / clamp(count, 1); zero rows for missing sentences.
    """
    B, L, D = hidden.shape
    smax = max_sents if max_sents is not None else max(len(s) for s in sent_tok_idxs)
    reps = torch.zeros(B, smax, D, dtype=hidden.dtype)
    for b, sents in enumerate(sent_tok_idxs):
        for s, toks in enumerate(sents[:smax]):
            if len(toks):
                sel = torch.zeros(L, dtype=hidden.dtype)
                sel[list(toks)] = 1.0
                # the reference multiplies by a dense 0/1 mask and sums over all L positions (:94-97)
                reps[b, s] = (hidden[b] * sel.unsqueeze(1)).sum(0) / max(int(sel.sum().item()), 1)
    return hidden[:, 0, :].clone(), reps


def caching_score(query_sent_reps, cand_sent_reps, agg="l2wasserstein", hparams=None):
    """WordSentAlignBiEnc.caching_score, disent_models.py:256-342 (sent_loss_prop=1, abs_loss_prop=0).

    1 query x B candidates: zero-pad candidates to cmax, replicate the query B times (:274-281), score with
    return_pair_sims=True.  Returns (batch_scores np[B], raw pair outputs).
    """
    hparams = hparams or {}
    B = len(cand_sent_reps)
    c_lens = [r.shape[0] for r in cand_sent_reps]
    cmax, qn, D = max(c_lens), query_sent_reps.shape[0], query_sent_reps.shape[1]
    cpad = np.zeros((B, cmax, D))
    for i, r in enumerate(cand_sent_reps):
        cpad[i, :c_lens[i]] = r
    qpad = np.broadcast_to(np.asarray(query_sent_reps)[None], (B, qn, D))
    q = torch.FloatTensor(np.ascontiguousarray(qpad))
    c = torch.FloatTensor(cpad)
    if agg == "l2max":
        best, idx, sims = l2max(q, [qn] * B, c, c_lens)
        return best.numpy(), sims.numpy()
    scores, extras = ot_distance(q, [qn] * B, c, c_lens, blur=hparams.get("geoml_blur", 0.05),
                                 scaling=hparams.get("geoml_scaling", 0.9), temp=hparams.get("sent_sm_temp", 1.0),
                                 return_pair_sims=True)
    return scores.numpy(), [t.numpy() for t in extras]


def average_precision(ranked_rel):
    """src/evaluation/utils/metrics.py:98-121 -- mean of precision@k at each relevant rank."""
    r = np.asarray(ranked_rel) != 0
    hits = np.flatnonzero(r)
    if hits.size == 0:
        return 0.0
    return float(np.mean([(r[:k + 1]).mean() for k in hits]))


def mean_average_precision(ranked_rels):
    """src/evaluation/utils/metrics.py:124-143."""
    return float(np.mean([average_precision(r) for r in ranked_rels]))


def l2top2(q, q_lens, c, c_lens):
    """allpair_masked_dist_l2topk (pair_distances.py:295-345), return_pair_sims=True branch: (sum of the two largest
    entries of -cdist + pad mask [B], the masked similarities [B,Sq,Sc]).  q [B,Sq,D], c [B,Sc,D]."""
    sims = neg_pair_dists(q, c) + _pad_mask(q_lens, c_lens, q.shape[1], c.shape[1])
    top = torch.topk(sims.reshape(sims.shape[0], -1), k=2, dim=1)[0]
    return top.sum(dim=1), sims


def attention_sim(q, q_lens, c, c_lens, temp=1.0):
    """AllPairMaskedAttention.compute_distance (pair_distances.py:95-135), return_pair_sims=True branch, with
    masked_2d_softmax (models_common/activations.py:35-61): (doc_sims [B], softmax [B,Sq,Sc])."""
    sims = neg_pair_dists(q, c)
    B, Sq, Sc = sims.shape
    mask = torch.zeros_like(sims)
    for b, (a, d) in enumerate(zip(q_lens, c_lens)):
        mask[b, a:, :] = -1e32
        mask[b, :, d:] = -1e32
    probs = torch.log_softmax((sims / temp + mask).reshape(B, -1), dim=1).reshape(B, Sq, Sc).exp()
    return (probs * sims).sum(dim=(1, 2)), probs
