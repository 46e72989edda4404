"""Pair-scoring functions of Aspire, backed by the sm_100a kernels (host-side mirror of the reference API).

Mirrors, name for name and argument for argument:
  * ``AllPairMaskedWasserstein`` / ``compute_distance``  -- src/learning/facetid_models/pair_distances.py:14-92
    (release copy examples/ex_aspire_consent_multimatch.py:111-189)                       [otAspire]
  * ``allpair_masked_dist_l2max``                         -- pair_distances.py:138-186     [tsAspire]
  * ``rep_len_tup``                                       -- disent_models.py:16 (namedtuple RepLen(embed, abs_lens))
plus array-level entry points (``ot_scores``, ``l2max_scores``) that the batched scorers and the bench use.

Inputs may live on the CPU (as in the reference, which never leaves the CPU in its release scripts) or on a
CUDA device; results come back on the device of ``query.embed``.  All arithmetic runs in the CUDA library --
there is no CPU implementation in this package.
"""
import ctypes
from collections import namedtuple

import numpy as np
import torch

from . import _abi

rep_len_tup = namedtuple("RepLen", ["embed", "abs_lens"])
RepLen = rep_len_tup


def _device():
    if not torch.cuda.is_available():
        raise _abi.AspireB200Error("aspire_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def epsilon_schedule(diameter, blur, scaling, p=1):
    """geomloss 0.2.4 ``epsilon_schedule`` in float64 (reached from pair_distances.py:68-72,88-91).

    [diam^p] + [exp(e) for e in arange(p ln diam, p ln blur, p ln scaling)] + [blur^p]
    """
    steps = np.arange(p * np.log(diameter), p * np.log(blur), p * np.log(scaling))
    return [float(diameter) ** p] + [float(np.exp(e)) for e in steps] + [float(blur) ** p]


def _as_bsd(embed, dev):
    """Reference layout [B, D, S] (pair_distances.py:28-31) -> contiguous fp32 [B, S, D] on the GPU."""
    if embed.dim() != 3:
        raise AssertionError("embed must be batch_size x encoding_dim x max_sents")
    t = embed.permute(0, 2, 1)
    return t.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()


def _lens_tensor(lens, dev):
    if isinstance(lens, torch.Tensor):
        return lens.to(device=dev, dtype=torch.int32).contiguous()
    return torch.tensor(list(lens), dtype=torch.int32).to(dev, non_blocking=True)


def bbox_diameter(x_rows, y_rows):
    """geomloss ``max_diameter`` over the rows of two fp32 CUDA tensors [..., D] (pad rows included)."""
    D = x_rows.shape[-1]
    dev = x_rows.device
    ws = torch.empty(2 * D + 1, dtype=torch.float32, device=dev)
    L = _abi.lib()
    _abi.check(L.asp_bbox_diameter(_abi.ptr(x_rows), x_rows.numel() // D, _abi.ptr(y_rows), y_rows.numel() // D, D,
                                   _abi.ptr(ws), ctypes.c_void_p(ws.data_ptr() + 8 * D), _abi.stream_of(dev)),
               "asp_bbox_diameter")
    return float(ws[2 * D].item())


_OT_FIELDS = ("dual", "primal", "f", "g", "alpha", "beta", "neg_cost", "plan", "weighted")


def ot_scores(q, q_lens, c, c_lens, eps_list, temp=1.0, want=("dual",), broadcast_query=False, cost_workspace=None,
              q_group=None, out=None, c_index=None):
    """Masked Sinkhorn OT on contiguous fp32 CUDA tensors through the fused entry point ``asp_ot_score``.

    c [B,Sc,D]; q [B,Sq,D] (paired, default), [1,Sq,D] with ``broadcast_query``, or [ceil(B/q_group),Sq,D] with
    ``q_group`` = number of consecutive candidates sharing one query (several query pools in one launch).
    lens: int32 CUDA tensors.  ``eps_list``: the epsilon schedule (python floats / float64).
    ``out``: optional dict of preallocated output tensors (reused across calls).  Returns a dict of outputs.
    ``c_index``: optional int32 CUDA tensor [B] -- pair b scores corpus document ``c_index[b]``; ``c`` / ``c_lens`` are
    then the whole resident corpus [N,Sc,D] / [N] and nothing is gathered (``asp_ot_score_indexed``; shapes the fused
    kernel does not cover are gathered here and go through ``asp_ot_score``).
    """
    _abi.require_cuda(q, c, q_lens, c_lens)
    if c_index is not None:
        _abi.require_cuda(c_index)
        assert c_index.dtype == torch.int32 and c_index.is_contiguous()
        N, Sc, D = c.shape
        if _abi.lib().asp_ot_score_workspace_bytes(1, q.shape[1], Sc, D) != 0:  # not a fused shape: gather
            idx = c_index.long()
            return ot_scores(q, q_lens, c.index_select(0, idx).contiguous(), c_lens.index_select(0, idx).contiguous(),
                             eps_list, temp=temp, want=want, broadcast_query=broadcast_query, q_group=q_group, out=out)
        B = int(c_index.numel())
    else:
        B, Sc, D = c.shape
    Sq = q.shape[1]
    if q_group is None:
        q_group = max(B, 1) if broadcast_query else 1
    assert q_group >= 1 and q.shape[0] == max(-(-B // q_group), 1 if broadcast_query else 0), \
        "query batch does not match candidates / q_group"
    assert q_lens.numel() == q.shape[0] and (c_index is not None or c_lens.numel() == B)
    assert q.is_contiguous() and c.is_contiguous() and q.dtype == c.dtype == torch.float32
    dev = c.device
    shapes = {"dual": (B,), "primal": (B,), "f": (B, Sq), "g": (B, Sc), "alpha": (B, Sq), "beta": (B, Sc),
              "neg_cost": (B, Sq, Sc), "plan": (B, Sq, Sc), "weighted": (B, Sq, Sc)}
    res = {k: (out[k] if out is not None and k in out else torch.empty(shapes[k], dtype=torch.float32, device=dev))
           for k in want}
    outs = _abi.AspOtOutputs(**{k: (res[k].data_ptr() if k in res else None) for k in _OT_FIELDS})
    L = _abi.lib()
    if c_index is not None:
        eps32 = np.asarray(eps_list, dtype=np.float32)
        _abi.check(L.asp_ot_score_indexed(_abi.ptr(q), _abi.ptr(q_lens), int(q_group), _abi.ptr(c), _abi.ptr(c_lens),
                                          _abi.ptr(c_index), B, Sq, Sc, D, eps32.ctypes.data_as(_abi.c_float_p), len(eps32),
                                          float(temp), ctypes.byref(outs), _abi.stream_of(dev)), "asp_ot_score_indexed")
        return res
    need = int(L.asp_ot_score_workspace_bytes(B, Sq, Sc, D))
    if need and (cost_workspace is None or cost_workspace.numel() * cost_workspace.element_size() < need):
        cost_workspace = torch.empty(need // 4, dtype=torch.float32, device=dev)
    ws_bytes = 0 if cost_workspace is None else cost_workspace.numel() * cost_workspace.element_size()
    eps32 = np.asarray(eps_list, dtype=np.float32)
    _abi.check(L.asp_ot_score(_abi.ptr(q), _abi.ptr(q_lens), int(q_group), _abi.ptr(c), _abi.ptr(c_lens),
                              B, Sq, Sc, D, eps32.ctypes.data_as(_abi.c_float_p), len(eps32), float(temp),
                              ctypes.byref(outs), _abi.ptr(cost_workspace), ws_bytes, _abi.stream_of(dev)),
               "asp_ot_score")
    return res


def ot_scores_allpairs(q, q_lens, c, c_lens, eps_list, temp=1.0, out=None):
    """otAspire dual values of EVERY query document against EVERY candidate document (``asp_ot_score_allpairs``).

    q [NQ,Sq,D], c [NC,Sc,D] contiguous fp32 CUDA, lens int32 CUDA.  Returns fp32 [NQ, NC] (distances; negate for
    similarities).  Documents of <= 10 sentences: ONE launch of the tcgen05 all-pairs kernel (Gram matrices on the tensor
    cores, Sinkhorn on the MUFU pipe); otherwise one fused 1 x N launch per query.  All on the current stream."""
    _abi.require_cuda(q, c, q_lens, c_lens)
    NQ, Sq, D = q.shape
    NC, Sc, _ = c.shape
    assert q.is_contiguous() and c.is_contiguous() and q.dtype == c.dtype == torch.float32 and c.shape[2] == D
    dev = c.device
    scores = out if out is not None else torch.empty((NQ, NC), dtype=torch.float32, device=dev)
    L = _abi.lib()
    need = int(L.asp_ot_score_allpairs_workspace_bytes(NQ, NC, Sq, Sc, D))
    ws = torch.empty(need, dtype=torch.uint8, device=dev) if need else None
    eps32 = np.asarray(eps_list, dtype=np.float32)
    _abi.check(L.asp_ot_score_allpairs(_abi.ptr(q), _abi.ptr(q_lens), NQ, _abi.ptr(c), _abi.ptr(c_lens), NC, Sq, Sc, D,
                                       eps32.ctypes.data_as(_abi.c_float_p), len(eps32), float(temp), _abi.ptr(scores),
                                       _abi.ptr(ws), need, _abi.stream_of(dev)), "asp_ot_score_allpairs")
    return scores


def l2max_scores(q, q_lens, c, c_lens, broadcast_query=False, want_pair_sims=False):
    """tsAspire on contiguous fp32 CUDA tensors: (best [B], flat argmax int32 [B], pair_sims or None)."""
    _abi.require_cuda(q, c, q_lens, c_lens)
    B, Sc, D = c.shape
    Sq = q.shape[1]
    dev = c.device
    best = torch.empty(B, dtype=torch.float32, device=dev)
    idx = torch.empty(B, dtype=torch.int32, device=dev)
    sims = torch.empty((B, Sq, Sc), dtype=torch.float32, device=dev) if want_pair_sims else None
    L = _abi.lib()
    need = int(L.asp_l2max_workspace_bytes(B, Sq, Sc, D))
    ws = torch.empty(need, dtype=torch.uint8, device=dev) if need else None
    _abi.check(L.asp_l2max_ws(_abi.ptr(q), _abi.ptr(q_lens), int(broadcast_query), _abi.ptr(c), _abi.ptr(c_lens),
                              B, Sq, Sc, D, _abi.ptr(best), _abi.ptr(idx), _abi.ptr(sims), _abi.ptr(ws), need,
                              _abi.stream_of(dev)), "asp_l2max")
    return best, idx, sims


def l2max_allpairs(q, q_lens, c, c_lens, want_idx=True):
    """tsAspire for EVERY (query, candidate) pair on the tensor cores (``asp_l2max_allpairs``).

    q [NQ,S,D], c [NC,S,D] contiguous fp32 CUDA (same S, zero padded), lens int32 CUDA.
    Returns (scores fp32 [NQ,NC] = max -dist, flat argmax int32 [NQ,NC] = i*S+j or None).
    """
    _abi.require_cuda(q, c, q_lens, c_lens)
    NQ, S, D = q.shape
    NC = c.shape[0]
    assert c.shape[1] == S and c.shape[2] == D and q.is_contiguous() and c.is_contiguous()
    dev = c.device
    L = _abi.lib()
    scores = torch.empty((NQ, NC), dtype=torch.float32, device=dev)
    idx = torch.empty((NQ, NC), dtype=torch.int32, device=dev) if want_idx else None
    ws = torch.empty(int(L.asp_l2max_allpairs_workspace_bytes(NQ, NC, S, D)), dtype=torch.uint8, device=dev)
    _abi.check(L.asp_l2max_allpairs(_abi.ptr(q), _abi.ptr(q_lens), NQ, _abi.ptr(c), _abi.ptr(c_lens), NC, S, D,
                                    _abi.ptr(scores), _abi.ptr(idx), _abi.ptr(ws), ws.numel(), _abi.stream_of(dev)),
               "asp_l2max_allpairs")
    return scores, idx


class AllPairMaskedWasserstein:
    """Drop-in for pair_distances.AllPairMaskedWasserstein (same hparam keys and defaults, :15-19).

    Extension key ``geoml_diameter``: fix the bounding-box diameter instead of deriving it from the call's
    batch as geomloss does (makes scores independent of how a pool is batched / sharded).
    """

    def __init__(self, model_hparams):
        self.geoml_blur = model_hparams.get('geoml_blur', 0.05)
        self.geoml_scaling = model_hparams.get('geoml_scaling', 0.9)
        self.geoml_reach = model_hparams.get('geoml_reach', None)
        self.sent_sm_temp = model_hparams.get('sent_sm_temp', 1.0)
        self.geoml_diameter = model_hparams.get('geoml_diameter', None)
        if self.geoml_reach is not None:
            raise NotImplementedError("unbalanced OT (geoml_reach) is not used by any released Aspire config")
        self.last_schedule = None

    def compute_distance(self, query, cand, return_pair_sims=False):
        """
        :param query: namedtuple(embed: batch_size x encoding_dim x q_max_sents; abs_lens: list(int))
        :param cand: namedtuple(embed: batch_size x encoding_dim x c_max_sents; abs_lens: list(int))
        :return: return_pair_sims=False -> OT_eps distances [B] (>= 0);
                 True -> (sum P*(-C) [B], [query_distr, cand_distr, pair_sims, transport_plan, masked_sims])
        """
        out_dev = query.embed.device
        dev = query.embed.device if query.embed.is_cuda else _device()
        qef_batch_size, _, qmax_sents = query.embed.size()
        cef_batch_size, encoding_dim, cmax_sents = cand.embed.size()
        assert (qef_batch_size == cef_batch_size)
        assert len(query.abs_lens) == qef_batch_size and len(cand.abs_lens) == cef_batch_size
        q = _as_bsd(query.embed, dev)
        c = _as_bsd(cand.embed, dev)
        ql, cl = _lens_tensor(query.abs_lens, dev), _lens_tensor(cand.abs_lens, dev)
        diameter = self.geoml_diameter
        if diameter is None:
            diameter = bbox_diameter(q, c)  # whole call batch incl. pad rows, like geomloss
        eps_list = epsilon_schedule(diameter, self.geoml_blur, self.geoml_scaling)
        self.last_schedule = (diameter, len(eps_list))
        if not return_pair_sims:
            res = ot_scores(q, ql, c, cl, eps_list, temp=self.sent_sm_temp, want=("dual",))
            return res["dual"].to(out_dev)
        res = ot_scores(q, ql, c, cl, eps_list, temp=self.sent_sm_temp,
                        want=("primal", "alpha", "beta", "neg_cost", "plan", "weighted"))
        extras = [res[k].to(out_dev) for k in ("alpha", "beta", "neg_cost", "plan", "weighted")]
        return res["primal"].to(out_dev), extras


def allpair_masked_dist_l2max(query, cand, return_pair_sims=False):
    """Drop-in for pair_distances.allpair_masked_dist_l2max (:138-186).

    return_pair_sims=True -> (batch_sims [B] = max -dist, pair_sims [B,Sq,Sc] with -1e9 on padding);
    False -> positive distances -max (what the triplet loss minimises).
    """
    out_dev = query.embed.device
    dev = query.embed.device if query.embed.is_cuda else _device()
    qef_batch_size, _, qmax_sents = query.embed.size()
    cef_batch_size, encoding_dim, cmax_sents = cand.embed.size()
    assert (qef_batch_size == cef_batch_size)
    q = _as_bsd(query.embed, dev)
    c = _as_bsd(cand.embed, dev)
    ql, cl = _lens_tensor(query.abs_lens, dev), _lens_tensor(cand.abs_lens, dev)
    best, _idx, sims = l2max_scores(q, ql, c, cl, want_pair_sims=return_pair_sims)
    if return_pair_sims:
        return best.to(out_dev), sims.to(out_dev)
    return (-1 * best).to(out_dev)


def pair_heads(q, q_lens, c, c_lens, temp=1.0, want=("top2",), raw_pads=False):
    """l2top2 / attention heads on contiguous fp32 CUDA tensors (paired): one ``asp_pair_cost`` + one ``asp_pair_heads``.

    Returns a dict with any of ``top2`` [B], ``att`` [B], ``att_probs`` [B,Sq,Sc], plus ``dist`` [B,Sq,Sc] (distances;
    0 on padding, or the raw distances to the zero pad rows with ``raw_pads`` -- what the attention head's
    ``pair_sims`` holds in the reference, pair_distances.py:121-127).
    """
    _abi.require_cuda(q, c, q_lens, c_lens)
    B, Sc, D = c.shape
    Sq = q.shape[1]
    dev = c.device
    L = _abi.lib()
    dist = torch.empty((B, Sq, Sc), dtype=torch.float32, device=dev)
    if raw_pads:
        fq = torch.full_like(q_lens, Sq)
        fc = torch.full_like(c_lens, Sc)
        _abi.check(L.asp_pair_cost(_abi.ptr(q), _abi.ptr(fq), 0, _abi.ptr(c), _abi.ptr(fc), B, Sq, Sc, D, _abi.ptr(dist),
                                   _abi.stream_of(dev)), "asp_pair_cost")
    else:
        _abi.check(L.asp_pair_cost(_abi.ptr(q), _abi.ptr(q_lens), 0, _abi.ptr(c), _abi.ptr(c_lens), B, Sq, Sc, D,
                                   _abi.ptr(dist), _abi.stream_of(dev)), "asp_pair_cost")
    res = {"dist": dist}
    shapes = {"top2": (B,), "att": (B,), "att_probs": (B, Sq, Sc)}
    for k in want:
        res[k] = torch.empty(shapes[k], dtype=torch.float32, device=dev)
    _abi.check(L.asp_pair_heads(_abi.ptr(dist), _abi.ptr(q_lens), 1, _abi.ptr(c_lens), B, Sq, Sc, float(temp),
                                _abi.ptr(res.get("top2")), _abi.ptr(res.get("att")), _abi.ptr(res.get("att_probs")),
                                _abi.stream_of(dev)), "asp_pair_heads")
    return res


def _pair_inputs(query, cand):
    out_dev = query.embed.device
    dev = query.embed.device if query.embed.is_cuda else _device()
    qef_batch_size, _, qmax_sents = query.embed.size()
    cef_batch_size, encoding_dim, cmax_sents = cand.embed.size()
    assert (qef_batch_size == cef_batch_size)
    return (_as_bsd(query.embed, dev), _lens_tensor(query.abs_lens, dev), _as_bsd(cand.embed, dev),
            _lens_tensor(cand.abs_lens, dev), out_dev)


def allpair_masked_dist_l2topk(query, cand, return_pair_sims=False):
    """Drop-in for pair_distances.allpair_masked_dist_l2topk (:295-345): the two best sentence matches.

    return_pair_sims=True -> (batch_sims [B] = sum of the two largest -dist, pair_sims [B,Sq,Sc] with -1e9 on padding);
    False -> the positive distance -sum.  Fewer than two valid sentence pairs: the runner-up is a masked entry (-1e9),
    as in the reference; a 1x1 padded shape raises like ``torch.topk(k=2)`` does there.
    """
    q, ql, c, cl, out_dev = _pair_inputs(query, cand)
    if q.shape[1] * c.shape[1] < 2:
        raise RuntimeError("selected index k out of range")  # torch.topk(k=2) over a single entry (:336)
    res = pair_heads(q, ql, c, cl, want=("top2",))
    if return_pair_sims:
        B, Sq, Sc = res["dist"].shape
        valid = (torch.arange(Sq, device=q.device)[None, :, None] < ql[:, None, None]) & \
                (torch.arange(Sc, device=q.device)[None, None, :] < cl[:, None, None])
        pair_sims = torch.where(valid, -res["dist"], torch.full_like(res["dist"], -10e8))
        return res["top2"].to(out_dev), pair_sims.to(out_dev)
    return (-1 * res["top2"]).to(out_dev)


class AllPairMaskedAttention:
    """Drop-in for pair_distances.AllPairMaskedAttention (:95-135): cross-document attention over sentence pairs."""

    def __init__(self, model_hparams):
        self.cdatt_sm_temp = model_hparams.get('cdatt_sm_temp', 1.0)

    def compute_distance(self, query, cand, return_pair_sims=False):
        """
        :param query / cand: namedtuple(embed: batch_size x encoding_dim x max_sents; abs_lens: list(int))
        :return: return_pair_sims=True -> (doc_sims [B], [pair_sims, pair_softmax, masked_sims]);
                 False -> doc_dists [B] = sum softmax * dist (the training loss' distance).
        """
        q, ql, c, cl, out_dev = _pair_inputs(query, cand)
        res = pair_heads(q, ql, c, cl, temp=self.cdatt_sm_temp, want=("att", "att_probs"), raw_pads=True)
        if return_pair_sims:
            # pad-vs-pad entries: both rows are zero vectors, torch.cdist gives exactly 0 there (the cost kernel's
            # 1e-8 clamp under the square root would give 1e-4)
            Sq, Sc = res["dist"].shape[1:]
            both_pad = (torch.arange(Sq, device=q.device)[None, :, None] >= ql[:, None, None]) & \
                       (torch.arange(Sc, device=q.device)[None, None, :] >= cl[:, None, None])
            pair_sims = -torch.where(both_pad, torch.zeros_like(res["dist"]), res["dist"])
            masked_sims = res["att_probs"] * pair_sims
            return res["att"].to(out_dev), [pair_sims.to(out_dev), res["att_probs"].to(out_dev), masked_sims.to(out_dev)]
        return (-1 * res["att"]).to(out_dev)


def allpair_masked_argmax_l2max(query, cand):
    """The flat argmax ``i*cmax_sents + j`` the reference computes at pair_distances.py:176 and drops."""
    dev = query.embed.device if query.embed.is_cuda else _device()
    q = _as_bsd(query.embed, dev)
    c = _as_bsd(cand.embed, dev)
    ql, cl = _lens_tensor(query.abs_lens, dev), _lens_tensor(cand.abs_lens, dev)
    best, idx, _ = l2max_scores(q, ql, c, cl)
    return best.to(query.embed.device), idx.to(device=query.embed.device, dtype=torch.int64)
