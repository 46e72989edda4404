"""Contextual sentence encoder front-end: batch preparation (host) + AspireConSent (device).

Mirrors the release API of the reference, keyword for keyword:
  * ``prepare_bert_sentences(batch_doc_sents, tokenizer)``  -- examples/ex_aspire_consent.py:107-181
    (= src/learning/batchers.py:555-630)
  * ``prepare_abstracts(batch_abs, pt_lm_tokenizer)``       -- examples/ex_aspire_consent.py:185-212
    (= src/learning/batchers.py:525-553)
  * ``AspireConSent(hf_model_name).forward(bert_batch, abs_lens, sent_tok_idxs)`` -- :25-101

The pooling step (K1) runs in the CUDA library (``asp_span_mean_pool``): sentence spans are contiguous token
ranges, so the kernel takes ``(start, end)`` pairs instead of the reference's dense [B,L,768] float64 masks.
"""
import threading

import numpy as np
import torch
from torch import nn

from . import _abi

MAX_WORDPIECES = 500  # examples/ex_aspire_consent.py:120


def _with_special_tokens(tokenizer, ids):
    """``[CLS] ids [SEP]`` -- ``tokenizer.build_inputs_with_special_tokens(token_ids_0=ids)`` as the reference calls it
    (:161); transformers >= 5 dropped that method from its tokenizers, so fall back to the two ids it would have used."""
    build = getattr(tokenizer, "build_inputs_with_special_tokens", None)
    if build is not None:
        return build(token_ids_0=ids)
    return [tokenizer.cls_token_id] + list(ids) + [tokenizer.sep_token_id]


def prepare_bert_sentences(batch_doc_sents, tokenizer):
    """Tokenise documents given as lists of "sentences" (element 0 is the title + ' [SEP] ').

    Per document the word-pieces of all sentences are concatenated into ONE sequence
    ``[CLS] w... [SEP]`` (no separator between abstract sentences); the token positions of every sentence
    are recorded (+1 for the leading [CLS]); at most 500 word-pieces are kept -- the sentence that crosses
    the budget is cut to fit and later ones are dropped; the title's positions are not returned.

    :return: (bert_batch{'tokid_tt','seg_tt','attnmask_tt': int64 [B,Lmax], 'seq_lens': list},
              batch_tokenized_text: list(list(str)), batch_sent_token_idxs: list(list(list(int))))
    """
    docs_ids, docs_text, docs_spans = [], [], []
    for doc_sents in batch_doc_sents:
        ids, text, spans = [], [], []
        used = 0
        for sent in doc_sents:
            pieces = tokenizer.tokenize(sent)
            piece_ids = tokenizer.convert_tokens_to_ids(pieces)
            room = MAX_WORDPIECES - used
            take = min(len(pieces), room)
            if take > 0 or len(pieces) == 0:
                # positions are shifted by one for the [CLS] prepended below
                spans.append(list(range(used + 1, used + 1 + take)))
                text.extend(pieces[:take])
                ids.extend(piece_ids[:take])
            if len(pieces) > room:
                break  # budget exhausted: this sentence was cut (or skipped when nothing fitted)
            used += take
        docs_text.append(text)
        docs_spans.append(spans[1:])  # the title is encoded but never pooled
        docs_ids.append(_with_special_tokens(tokenizer, ids))
    seq_lens = [len(x) for x in docs_ids]
    width = max(seq_lens) if seq_lens else 0
    pad = tokenizer.pad_token_id
    tok = [x + [pad] * (width - len(x)) for x in docs_ids]
    seg = [[0] * n + [pad] * (width - n) for n in seq_lens]
    att = [[1] * n + [pad] * (width - n) for n in seq_lens]
    bert_batch = {
        'tokid_tt': torch.tensor(tok),
        'seg_tt': torch.tensor(seg),
        'attnmask_tt': torch.tensor(att),
        'seq_lens': seq_lens,
    }
    return bert_batch, docs_text, docs_spans


def prepare_abstracts_per_sentence(batch_abs, pt_lm_tokenizer):
    """``prepare_abstracts`` exactly as the reference runs it: ``tokenize`` + ``convert_tokens_to_ids`` once per sentence
    (examples/ex_aspire_consent.py:185-212 over :107-181).  Works with any object exposing the four tokenizer members
    the reference touches; 0.9 k documents/s.

    :return: (bert_batch, abs_lens: list(int), sent_token_idxs: list(list(list(int))))
    """
    seqs = [[ex['TITLE'] + ' [SEP] '] + list(ex['ABSTRACT']) for ex in batch_abs]
    bert_batch, _tokenized, sent_token_idxs = prepare_bert_sentences(batch_doc_sents=seqs,
                                                                     tokenizer=pt_lm_tokenizer)
    abs_lens = [len(s) for s in sent_token_idxs]
    for n in abs_lens:
        assert (n > 0)  # an abstract whose title alone fills the budget (reference :210)
    return bert_batch, abs_lens, sent_token_idxs


def prepare_abstracts(batch_abs, pt_lm_tokenizer):
    """
    Drop-in for the reference's ``prepare_abstracts`` (examples/ex_aspire_consent.py:185-212): same arguments, same
    return value.  When the tokenizer is the plain BERT word-piece pipeline the batch is built by the library's host
    code (``prepare_abstracts_native``, 50-100 k documents/s) and the sentence positions are expanded back into the
    index lists the reference returns; any other tokenizer goes through the per-sentence protocol.

    :param batch_abs: list(dict) with 'TITLE' (str) and 'ABSTRACT' (list of sentence strings).
    :return: (bert_batch, abs_lens: list(int), sent_token_idxs: list(list(list(int))))
    """
    if native_wordpiece(pt_lm_tokenizer) is None:
        return prepare_abstracts_per_sentence(batch_abs, pt_lm_tokenizer)
    bert_batch, abs_lens, spans = prepare_abstracts_native(batch_abs, pt_lm_tokenizer)
    sent_token_idxs = [[list(range(st, en)) for st, en in doc[:n]] for doc, n in zip(spans.tolist(), abs_lens)]
    return bert_batch, abs_lens, sent_token_idxs


def prepare_abstracts_fast(batch_abs, pt_lm_tokenizer):
    """Batch-first variant of ``prepare_abstracts`` for the encoder kernels ("next" row 3 of SURVEY 8f).

    Same sequence layout, truncation rule and outputs as ``prepare_abstracts`` (examples/ex_aspire_consent.py:107-212),
    but (i) every sentence of the batch is word-pieced in ONE tokenizer call when the tokenizer is a fast (Rust-backed)
    one -- the reference calls ``tokenize`` + ``convert_tokens_to_ids`` once per sentence -- and (ii) the sentence
    spans come back as the int32 ``[B, Smax, 2]`` half-open ``(start, end)`` table that ``asp_span_mean_pool`` consumes,
    instead of per-token index lists.  Tokenizers without batch encoding fall back to the per-sentence protocol.

    :return: (bert_batch, abs_lens: list(int), spans: int32 tensor [B, max(abs_lens), 2], (-1, -1) = no sentence)
    """
    docs = [[ex['TITLE'] + ' [SEP] '] + list(ex['ABSTRACT']) for ex in batch_abs]
    flat = [s for d in docs for s in d]
    if getattr(pt_lm_tokenizer, "is_fast", False) and flat:
        piece_ids = pt_lm_tokenizer(flat, add_special_tokens=False, return_attention_mask=False,
                                    return_token_type_ids=False)["input_ids"]
    else:
        piece_ids = [pt_lm_tokenizer.convert_tokens_to_ids(pt_lm_tokenizer.tokenize(s)) for s in flat]
    docs_ids, docs_spans, cursor = [], [], 0
    for d in docs:
        ids, spans, used = [], [], 0
        for k in range(len(d)):
            p = piece_ids[cursor + k]
            room = MAX_WORDPIECES - used
            take = min(len(p), room)
            if take > 0 or len(p) == 0:
                spans.append((used + 1, used + 1 + take))  # +1: the [CLS] prepended below
                ids.extend(p[:take])
            if len(p) > room:
                break
            used += take
        cursor += len(d)
        docs_spans.append(spans[1:])  # the title is encoded but never pooled
        docs_ids.append(_with_special_tokens(pt_lm_tokenizer, ids))
    abs_lens = [len(s) for s in docs_spans]
    for n in abs_lens:
        assert (n > 0)
    seq_lens = [len(x) for x in docs_ids]
    width, pad = max(seq_lens), pt_lm_tokenizer.pad_token_id
    tok = np.full((len(docs), width), pad, dtype=np.int64)
    seg = np.full((len(docs), width), pad, dtype=np.int64)
    att = np.full((len(docs), width), pad, dtype=np.int64)
    span_arr = -np.ones((len(docs), max(abs_lens), 2), dtype=np.int32)
    for b, (ids, spans) in enumerate(zip(docs_ids, docs_spans)):
        tok[b, :len(ids)] = ids
        seg[b, :len(ids)] = 0
        att[b, :len(ids)] = 1
        for s_, (st, en) in enumerate(spans):
            if en > st:
                span_arr[b, s_] = (st, en)
    bert_batch = {'tokid_tt': torch.from_numpy(tok), 'seg_tt': torch.from_numpy(seg),
                  'attnmask_tt': torch.from_numpy(att), 'seq_lens': seq_lens}
    return bert_batch, abs_lens, torch.from_numpy(span_arr)


class NativeWordPiece:
    """The library's multi-threaded BERT word-piece tokenizer (``asp_wordpiece_*``) configured from a Hugging Face fast
    tokenizer.  Construction raises ``ValueError`` when the tokenizer is anything but the plain BERT pipeline the
    reference loads (BertNormalizer with text cleaning -> BertPreTokenizer -> WordPiece with the ``##`` prefix, special
    tokens matched verbatim), so callers can fall back to ``prepare_abstracts_fast``."""

    def __init__(self, tokenizer, threads=None):
        import json
        import os
        backend = getattr(tokenizer, "backend_tokenizer", None)
        if backend is None:
            raise ValueError("not a fast (tokenizers-backed) tokenizer")
        cfg = json.loads(backend.to_str())
        norm, pre, model = cfg.get("normalizer") or {}, cfg.get("pre_tokenizer") or {}, cfg.get("model") or {}
        if norm.get("type") != "BertNormalizer" or not norm.get("clean_text", False):
            raise ValueError("normalizer is not BertNormalizer(clean_text=True)")
        if pre.get("type") != "BertPreTokenizer":
            raise ValueError("pre-tokenizer is not BertPreTokenizer")
        if model.get("type") != "WordPiece" or model.get("continuing_subword_prefix") != "##":
            raise ValueError("model is not WordPiece with the '##' prefix")
        if cfg.get("truncation") or cfg.get("padding"):
            raise ValueError("tokenizer has truncation / padding enabled")
        vocab = dict(model["vocab"])
        specials = []
        for tok in cfg.get("added_tokens", []):
            if tok.get("single_word") or tok.get("lstrip") or tok.get("rstrip") or tok.get("normalized"):
                raise ValueError(f"added token {tok.get('content')!r} uses matching options the native path lacks")
            vocab.setdefault(tok["content"], tok["id"])
            if vocab[tok["content"]] != tok["id"]:
                raise ValueError(f"added token {tok['content']!r} has two ids")
            specials.append(tok["id"])
        by_id = [None] * (max(vocab.values()) + 1)
        for t, i in vocab.items():
            by_id[i] = t
        encoded = [(t or "").encode("utf-8") for t in by_id]
        offsets = np.zeros(len(encoded) + 1, dtype=np.int64)
        np.cumsum([len(b) for b in encoded], out=offsets[1:])
        blob = b"".join(encoded)
        special_ids = np.asarray(specials, dtype=np.int32)
        self.max_chars = int(model.get("max_input_chars_per_word", 100))
        self.threads = int(threads or min(16, os.cpu_count() or 1))
        self.tokenizer = tokenizer
        self._unicode = False
        self._install_lock = threading.Lock()
        self._lib = _abi.lib()
        self._handle = self._lib.asp_wordpiece_create(blob, offsets.ctypes.data, len(encoded), int(bool(norm.get("lowercase", True))),
                                                      int(vocab[model["unk_token"]]), special_ids.ctypes.data, len(special_ids))
        if not self._handle:
            raise ValueError("asp_wordpiece_create: " + self._lib.asp_last_error().decode("utf8", "replace"))

    def _install_unicode(self):
        """Per-code-point tables of the Basic Multilingual Plane, taken from the tokenizer's own normaliser and
        pre-tokenizer: what each character is normalised to, and whether it separates words (space) or stands alone
        (punctuation).  Context-dependent characters (capital sigma: final-sigma rule of lower-casing), surrogates and
        anything the normaliser rejects stay with the Hugging Face tokenizer, as do characters beyond the BMP."""
        self._unicode = None  # tried
        backend = self.tokenizer.backend_tokenizer
        norm, pre = backend.normalizer, backend.pre_tokenizer
        pieces, offsets = [], np.zeros(0x10001, dtype=np.uint32)
        out_class = np.zeros(0x10000, dtype=np.uint8)
        fallback = np.zeros(0x10000, dtype=np.uint8)
        total = 0
        for cp in range(0x10000):
            out = b""
            if 0xD800 <= cp <= 0xDFFF or cp == 0x3A3:
                fallback[cp] = 1
            else:
                ch = chr(cp)
                try:
                    out = norm.normalize_str(ch).encode("utf-8")
                    parts = [t for t, _ in pre.pre_tokenize_str("a" + ch + "a")]
                    if parts == ["a", "a"]:
                        out_class[cp] = 1
                    elif parts == ["a", ch, "a"]:
                        out_class[cp] = 2
                    elif parts != ["a" + ch + "a"]:
                        fallback[cp] = 1
                except Exception:
                    fallback[cp], out = 1, b""
            pieces.append(out)
            total += len(out)
            offsets[cp + 1] = total
        blob = b"".join(pieces)
        rc = self._lib.asp_wordpiece_set_unicode(self._handle, offsets.ctypes.data, blob, out_class.ctypes.data,
                                                 fallback.ctypes.data)
        self._unicode = True if rc == 0 else None

    def __del__(self):
        if getattr(self, "_handle", None):
            self._lib.asp_wordpiece_destroy(self._handle)
            self._handle = None

    def encode(self, sentences):
        """list(str) -> (ids int32 [total], offsets int64 [n+1]); same ids as ``tokenizer(s, add_special_tokens=False)``.
        Non-ASCII sentences use per-character tables read off the tokenizer itself (installed on first need); what
        those cannot express (characters beyond the BMP, capital sigma, malformed text) goes through the Hugging Face
        tokenizer."""
        raw = [s.encode("utf-8") for s in sentences]
        n = len(raw)
        offsets = np.zeros(n + 1, dtype=np.int64)
        if n:
            np.cumsum(np.fromiter(map(len, raw), dtype=np.int64, count=n), out=offsets[1:])
        ids = np.empty(max(int(offsets[-1]), 1), dtype=np.int32)
        out_offsets = np.zeros(n + 1, dtype=np.int64)
        fallback = np.zeros(max(n, 1), dtype=np.uint8)
        _abi.check(self._lib.asp_wordpiece_encode(self._handle, b"".join(raw), offsets.ctypes.data, n, self.max_chars, self.threads,
                                                  ids.ctypes.data, out_offsets.ctypes.data, fallback.ctypes.data),
                   "asp_wordpiece_encode")
        todo = np.flatnonzero(fallback[:n])
        if len(todo) and self._unicode is False:  # None = tried and unavailable
            # first non-ASCII sentence: give the library the tokenizer's own per-character tables and run it again
            with self._install_lock:
                if self._unicode is False:
                    self._install_unicode()
            if self._unicode:
                return self.encode(sentences)
        ids = ids[:out_offsets[-1]]
        if len(todo):
            extra = self.tokenizer([sentences[i] for i in todo], add_special_tokens=False, return_attention_mask=False,
                                   return_token_type_ids=False)["input_ids"]
            parts, prev = [], 0
            for i, e in zip(todo, extra):  # the native pass left zero ids for these sentences: splice theirs in
                parts.append(ids[prev:out_offsets[i]])
                parts.append(np.asarray(e, dtype=np.int32))
                prev = out_offsets[i]
            parts.append(ids[prev:])
            lens = np.diff(out_offsets)
            lens[todo] = [len(e) for e in extra]
            ids = np.concatenate(parts) if parts else ids
            out_offsets = np.zeros(n + 1, dtype=np.int64)
            np.cumsum(lens, out=out_offsets[1:])
        return ids, out_offsets


def native_wordpiece(tokenizer):
    """The cached ``NativeWordPiece`` of a tokenizer, or None when the tokenizer is not the plain BERT pipeline."""
    cached = getattr(tokenizer, "_asp_native_wordpiece", None)
    if cached is None:
        try:
            cached = NativeWordPiece(tokenizer)
        except (ValueError, KeyError, TypeError):
            cached = False
        try:
            object.__setattr__(tokenizer, "_asp_native_wordpiece", cached)
        except Exception:  # a tokenizer that refuses new attributes is simply probed again next time
            pass
    return cached or None


def prepare_abstracts_native(batch_abs, pt_lm_tokenizer):
    """``prepare_abstracts_fast`` with the word pieces, the 500-piece truncation, the [CLS]..[SEP] wrapping, the padding
    and the span table all produced by the library's host code (``asp_wordpiece_encode`` / ``asp_abstracts_plan`` /
    ``asp_abstracts_fill``; examples/ex_aspire_consent.py:107-212).  Same return value, bit for bit; falls back to
    ``prepare_abstracts_fast`` for tokenizers the native path does not cover.

    :return: (bert_batch, abs_lens: list(int), spans: int32 tensor [B, max(abs_lens), 2], (-1, -1) = no sentence)
    """
    wp = native_wordpiece(pt_lm_tokenizer)
    if wp is None:
        return prepare_abstracts_fast(batch_abs, pt_lm_tokenizer)
    flat, doc_sents = [], np.empty(len(batch_abs), dtype=np.int32)
    for d, ex in enumerate(batch_abs):
        flat.append(ex['TITLE'] + ' [SEP] ')
        flat.extend(ex['ABSTRACT'])
        doc_sents[d] = 1 + len(ex['ABSTRACT'])
    ids, offsets = wp.encode(flat)
    L_, n_docs = _abi.lib(), len(batch_abs)
    seq_lens = np.empty(max(n_docs, 1), dtype=np.int32)
    abs_lens = np.empty(max(n_docs, 1), dtype=np.int32)
    _abi.check(L_.asp_abstracts_plan(offsets.ctypes.data, doc_sents.ctypes.data, n_docs, MAX_WORDPIECES, seq_lens.ctypes.data,
                                     abs_lens.ctypes.data), "asp_abstracts_plan")
    seq_lens, abs_lens = seq_lens[:n_docs].tolist(), abs_lens[:n_docs].tolist()
    for n in abs_lens:
        assert (n > 0)
    width, max_sents = max(seq_lens), max(abs_lens)
    tok = torch.empty((n_docs, width), dtype=torch.int64)
    seg = torch.empty((n_docs, width), dtype=torch.int64)
    att = torch.empty((n_docs, width), dtype=torch.int64)
    spans = torch.empty((n_docs, max_sents, 2), dtype=torch.int32)
    ids = np.ascontiguousarray(ids)
    _abi.check(L_.asp_abstracts_fill(ids.ctypes.data if len(ids) else None, offsets.ctypes.data, doc_sents.ctypes.data, n_docs,
                                     MAX_WORDPIECES, int(pt_lm_tokenizer.cls_token_id), int(pt_lm_tokenizer.sep_token_id),
                                     int(pt_lm_tokenizer.pad_token_id), width, max_sents, tok.data_ptr(), seg.data_ptr(),
                                     att.data_ptr(), spans.data_ptr()), "asp_abstracts_fill")
    bert_batch = {'tokid_tt': tok, 'seg_tt': seg, 'attnmask_tt': att, 'seq_lens': seq_lens}
    return bert_batch, abs_lens, spans


def spans_from_token_idxs(sent_tok_idxs, max_sents):
    """list[B][S][tokens] -> int32 [B, max_sents, 2] half-open (start, end); (-1,-1) for missing sentences."""
    B = len(sent_tok_idxs)
    out = -np.ones((B, max_sents, 2), dtype=np.int32)
    for b, sents in enumerate(sent_tok_idxs):
        for s, toks in enumerate(sents[:max_sents]):
            if len(toks) == 0:
                continue
            start, end = int(toks[0]), int(toks[-1]) + 1
            if end - start != len(toks):
                raise NotImplementedError("span_mean_pool needs contiguous token ranges per sentence")
            out[b, s] = (start, end)
    return torch.from_numpy(out)


def span_mean_pool(hidden, spans):
    """K1 on the GPU: hidden fp32 CUDA [B,L,D], spans int32 CUDA [B,Smax,2] -> (cls [B,D], sent_reps [B,Smax,D])."""
    _abi.require_cuda(hidden, spans)
    B, L, D = hidden.shape
    smax = spans.shape[1]
    hidden = hidden.contiguous()
    spans = spans.contiguous()
    reps = torch.empty((B, smax, D), dtype=torch.float32, device=hidden.device)
    cls = torch.empty((B, D), dtype=torch.float32, device=hidden.device)
    _abi.check(_abi.lib().asp_span_mean_pool(_abi.ptr(hidden), _abi.ptr(spans), B, L, D, smax, _abi.ptr(reps),
                                             _abi.ptr(cls), _abi.stream_of(hidden.device)), "asp_span_mean_pool")
    return cls, reps


class AspireConSent(nn.Module):
    """Drop-in for examples/ex_aspire_consent.py:25-101 (and the multimatch copy :30-106).

    ``bert_encoder`` keeps the reference's attribute name so released state dicts load unchanged
    (SURVEY appendix A.11).  The encoder and the pooling kernel run on the current CUDA device; outputs are
    returned on the device of ``bert_batch['tokid_tt']`` (CPU in the reference's scripts).
    """

    def __init__(self, hf_model_name):
        torch.nn.Module.__init__(self)
        from transformers import AutoModel
        self.bert_encoding_dim = 768
        self.bert_layer_count = 12 + 1
        self.bert_encoder = AutoModel.from_pretrained(hf_model_name)
        self.bert_encoder.config.output_hidden_states = True

    def forward(self, bert_batch, abs_lens, sent_tok_idxs):
        """
        :return: doc_cls_reps [B, 768], sent_reps [B, max(abs_lens), 768] (zero rows past abs_lens[b]).
        """
        doc_cls_reps, sent_reps = self.consent_reps_bert(bert_batch=bert_batch, num_sents=abs_lens,
                                                         batch_senttok_idxs=sent_tok_idxs)
        if len(sent_reps.size()) == 2:
            sent_reps = sent_reps.unsqueeze(0)
        if len(doc_cls_reps.size()) == 1:
            doc_cls_reps = doc_cls_reps.unsqueeze(0)
        return doc_cls_reps, sent_reps

    # "bf16x3": fp32-equivalent split-bf16 tensor-core math (default, parity with the reference's fp32 forward);
    # "bf16": plain bf16 operands (fastest).  Both run the sm_100a kernels; there is no library backend in this class
    # (the Hugging Face forward used as a comparison lives in tools/encoder_bench.py and the tests).
    encoder_precision = "bf16x3"

    def _device(self):
        if not torch.cuda.is_available():
            raise _abi.AspireB200Error("AspireConSent needs a CUDA device; there is no CPU fallback")
        return torch.device("cuda", torch.cuda.current_device())

    def native_encoder(self):
        """The repacked weights for ``asp_bert_forward`` (built on first use; call ``reset_native_encoder`` after
        loading a new state dict)."""
        dev = self._device()
        enc = getattr(self, "_native", None)
        if enc is None or enc.device != dev:
            from .encoder import B200BertEncoder
            enc = B200BertEncoder(self.bert_encoder, device=dev)
            object.__setattr__(self, "_native", enc)
        return enc

    def reset_native_encoder(self):
        object.__setattr__(self, "_native", None)

    def load_state_dict(self, *args, **kwargs):
        self.reset_native_encoder()
        return super().load_state_dict(*args, **kwargs)

    def encode_hidden(self, tokid_tt, seg_tt, attnmask_tt, seq_lens=None):
        """BERT forward -> last_hidden_state fp32 [B,L,768] on the GPU (K0)."""
        if self.encoder_precision not in ("bf16x3", "bf16"):
            raise ValueError(f"encoder_precision must be 'bf16x3' or 'bf16' (got {self.encoder_precision!r})")
        if seq_lens is None:
            seq_lens = attnmask_tt.sum(dim=1)
        return self.native_encoder().forward(tokid_tt, seq_lens, type_ids=seg_tt, precision=self.encoder_precision)

    def consent_reps_bert(self, bert_batch, batch_senttok_idxs, num_sents):
        out_dev = bert_batch['tokid_tt'].device
        dev = self._device()
        max_sents = max(num_sents)
        tokid_tt, seg_tt, attnmask_tt = (bert_batch[k].to(dev, non_blocking=True)
                                         for k in ('tokid_tt', 'seg_tt', 'attnmask_tt'))
        if isinstance(batch_senttok_idxs, torch.Tensor):  # a ready (start, end) table from prepare_abstracts_fast
            spans = batch_senttok_idxs[:, :max_sents].to(device=dev, dtype=torch.int32, non_blocking=True)
        else:
            spans = spans_from_token_idxs(batch_senttok_idxs, max_sents).to(dev, non_blocking=True)
        hidden = self.encode_hidden(tokid_tt, seg_tt, attnmask_tt, seq_lens=bert_batch.get('seq_lens'))
        doc_cls_reps, sent_reps = span_mean_pool(hidden, spans)
        doc_cls_reps = doc_cls_reps.squeeze()  # reference :76 ([768] when B == 1; forward() re-expands)
        return doc_cls_reps.to(out_dev), sent_reps.to(out_dev)
