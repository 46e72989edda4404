"""K0 on the B200: BERT-base encoder forward through the hand-written sm_100a kernels (``asp_bert_forward``).

Takes the weights of the HF ``BertModel`` that ``AspireConSent`` loads (examples/ex_aspire_consent.py:33, attribute
``bert_encoder`` -- so released checkpoints load unchanged), repacks them once into the layout
of ``include/aspire_b200.h::asp_bert_weights`` (fused Q|K|V projection, bf16 hi / lo halves of every Linear weight,
fp32 embeddings / biases / LayerNorm parameters) and replaces the ``self.bert_encoder(...)`` call of
``consent_reps_bert`` (:72).  No CPU path: the constructor needs a CUDA device.

precision:
  "bf16x3" (default) -- every GEMM / attention operand is a (hi, lo) bf16 pair, products are hi.hi + hi.lo + lo.hi:
                         fp32-equivalent, matches the reference's fp32 forward to ~1e-5;
  "bf16"             -- plain bf16 tensor-core operands with fp32 accumulation (3x fewer MMAs).
"""
import ctypes

import torch

from . import _abi

_vp, _fp = ctypes.c_void_p, ctypes.c_void_p


class AspBertLayer(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in
                ("wqkv_hi", "wqkv_lo", "bqkv", "wo_hi", "wo_lo", "bo", "ln1_g", "ln1_b",
                 "w1_hi", "w1_lo", "b1", "w2_hi", "w2_lo", "b2", "ln2_g", "ln2_b")]


class AspBertWeights(ctypes.Structure):
    _fields_ = [("hidden", ctypes.c_int), ("heads", ctypes.c_int), ("layers", ctypes.c_int),
                ("intermediate", ctypes.c_int), ("vocab", ctypes.c_int), ("max_pos", ctypes.c_int),
                ("ln_eps", ctypes.c_float),
                ("word_emb", ctypes.c_void_p), ("pos_emb", ctypes.c_void_p), ("type_emb", ctypes.c_void_p),
                ("emb_ln_g", ctypes.c_void_p), ("emb_ln_b", ctypes.c_void_p),
                ("layer", ctypes.POINTER(AspBertLayer))]


def _bind(L):
    if getattr(L, "_bert_bound", False):
        return
    L.asp_bert_workspace_bytes.argtypes = [ctypes.POINTER(AspBertWeights), ctypes.c_int, ctypes.c_int, ctypes.c_int]
    L.asp_bert_workspace_bytes.restype = ctypes.c_size_t
    L.asp_bert_forward.argtypes = [ctypes.POINTER(AspBertWeights), _vp, _vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                   _vp, _vp, ctypes.c_size_t, _vp]
    L.asp_bert_forward.restype = ctypes.c_int
    L._bert_bound = True


def split_bf16(w):
    """fp32 tensor -> (hi, lo) contiguous bf16 tensors with hi + lo == w to ~2^-17 relative."""
    w = w.detach().float()
    hi = w.to(torch.bfloat16)
    lo = (w - hi.float()).to(torch.bfloat16)
    return hi.contiguous(), lo.contiguous()


class B200BertEncoder:
    """Device-resident repack of an HF ``BertModel`` + the forward call."""

    def __init__(self, hf_bert, device=None):
        if not torch.cuda.is_available():
            raise _abi.AspireB200Error("B200BertEncoder needs a CUDA device; there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        cfg = hf_bert.config
        if cfg.hidden_size != 768 or cfg.num_attention_heads * 64 != cfg.hidden_size:
            raise NotImplementedError("the sm_100a encoder kernels are built for BERT-base (hidden 768, 12 heads of 64)")
        if getattr(cfg, "hidden_act", "gelu") != "gelu":
            raise NotImplementedError("only the exact-erf GELU of BERT is implemented")
        if getattr(cfg, "position_embedding_type", "absolute") != "absolute":
            raise NotImplementedError("only absolute position embeddings are implemented")
        sd = {k: v.detach() for k, v in hf_bert.state_dict().items()}
        dev = self.device
        self._keep = []  # owns every device tensor referenced by the C structs

        def f32(name):
            t = sd[name].to(dev, torch.float32).contiguous()
            self._keep.append(t)
            return t.data_ptr()

        def lin(weight, bias):
            hi, lo = split_bf16(weight.to(dev))
            b = bias.to(dev, torch.float32).contiguous()
            self._keep += [hi, lo, b]
            return hi.data_ptr(), lo.data_ptr(), b.data_ptr()

        n_layers = cfg.num_hidden_layers
        self._layers = (AspBertLayer * n_layers)()
        for i in range(n_layers):
            p = f"encoder.layer.{i}."
            wq, wk, wv = (sd[p + f"attention.self.{n}.weight"] for n in ("query", "key", "value"))
            bq, bk, bv = (sd[p + f"attention.self.{n}.bias"] for n in ("query", "key", "value"))
            y = self._layers[i]
            y.wqkv_hi, y.wqkv_lo, y.bqkv = lin(torch.cat([wq, wk, wv], 0), torch.cat([bq, bk, bv], 0))
            y.wo_hi, y.wo_lo, y.bo = lin(sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"])
            y.ln1_g, y.ln1_b = f32(p + "attention.output.LayerNorm.weight"), f32(p + "attention.output.LayerNorm.bias")
            y.w1_hi, y.w1_lo, y.b1 = lin(sd[p + "intermediate.dense.weight"], sd[p + "intermediate.dense.bias"])
            y.w2_hi, y.w2_lo, y.b2 = lin(sd[p + "output.dense.weight"], sd[p + "output.dense.bias"])
            y.ln2_g, y.ln2_b = f32(p + "output.LayerNorm.weight"), f32(p + "output.LayerNorm.bias")
        w = AspBertWeights()
        w.hidden, w.heads, w.layers = cfg.hidden_size, cfg.num_attention_heads, n_layers
        w.intermediate, w.vocab, w.max_pos = cfg.intermediate_size, cfg.vocab_size, cfg.max_position_embeddings
        w.ln_eps = float(cfg.layer_norm_eps)
        w.word_emb = f32("embeddings.word_embeddings.weight")
        w.pos_emb = f32("embeddings.position_embeddings.weight")
        w.type_emb = f32("embeddings.token_type_embeddings.weight")
        w.emb_ln_g, w.emb_ln_b = f32("embeddings.LayerNorm.weight"), f32("embeddings.LayerNorm.bias")
        w.layer = ctypes.cast(self._layers, ctypes.POINTER(AspBertLayer))
        self._w = w
        self.hidden = cfg.hidden_size
        self._ws = None

    def forward(self, ids, seq_lens, type_ids=None, precision="bf16x3", out=None):
        """ids int [B,L] (right padded), seq_lens int [B] -> last_hidden_state fp32 CUDA [B,L,hidden]."""
        if precision not in ("bf16x3", "bf16"):
            raise ValueError(f"unknown precision {precision!r}")
        L_ = _abi.lib()
        _bind(L_)
        dev = self.device
        ids = ids.to(dev, torch.int32, non_blocking=True).contiguous()
        lens = torch.as_tensor(seq_lens).to(dev, torch.int32, non_blocking=True).contiguous()
        tt = None if type_ids is None else type_ids.to(dev, torch.int32, non_blocking=True).contiguous()
        B, L = ids.shape
        precise = 1 if precision == "bf16x3" else 0
        need = int(L_.asp_bert_workspace_bytes(ctypes.byref(self._w), B, L, precise))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        hidden = out if out is not None else torch.empty((B, L, self.hidden), dtype=torch.float32, device=dev)
        _abi.check(L_.asp_bert_forward(ctypes.byref(self._w), _abi.ptr(ids), _abi.ptr(tt), _abi.ptr(lens), B, L, precise,
                                       _abi.ptr(hidden), _abi.ptr(self._ws), self._ws.numel(), _abi.stream_of(dev)),
                   "asp_bert_forward")
        return hidden
