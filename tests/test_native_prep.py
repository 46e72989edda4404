"""Host front end of the encoder (`asp_wordpiece_*`, `asp_abstracts_*`; no GPU involved): the library's BERT word-piece
tokenizer and sequence assembly against the Hugging Face fast tokenizer + `prepare_abstracts_fast`, which in turn is
pinned to the reference-shaped per-sentence protocol in test_host_api.py.  Integer work: everything must match exactly.
"""
import ctypes
import random

import numpy as np
import pytest
import torch

from oracle import ref_shims

SYLL = ["al", "ign", "trans", "port", "op", "ti", "mal", "bio", "med", "ic", "graph", "neur", "net", "work", "re", "triev",
        "sent", "ence", "pa", "per"]
PUNCT = ".,;:()[]#-!?'\"/\\{}<>=+*&^%$@~`|_"


def _tokenizer(tmp_path, lower=True, extra=()):
    from transformers import BertTokenizerFast
    rnd = random.Random(11)
    words = sorted({"".join(rnd.choice(SYLL) for _ in range(rnd.randint(1, 3))) for _ in range(1500)})
    vocab = (["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"] + words + ["##" + s for s in SYLL] + list(PUNCT) +
             ["a", "b", "##a", "##b", "1", "2", "##1", "##2", "Trans", "##Port"] + list(extra))
    vf = tmp_path / ("vocab_%d.txt" % lower)
    vf.write_text("\n".join(vocab) + "\n")
    return BertTokenizerFast(vocab_file=str(vf), do_lower_case=lower), words


def _random_word(rnd, words):
    r = rnd.random()
    if r < 0.55:
        return rnd.choice(words)
    if r < 0.65:
        return rnd.choice(words).upper() if rnd.random() < 0.5 else rnd.choice(words).capitalize()
    if r < 0.72:
        return rnd.choice(words) + rnd.choice(PUNCT) + rnd.choice(words)
    if r < 0.78:
        return "".join(rnd.choice("ab12xyz") for _ in range(rnd.randint(1, 12)))
    if r < 0.81:
        return "x" * rnd.randint(95, 105)  # straddles max_input_chars_per_word = 100
    if r < 0.85:
        return rnd.choice(["[SEP]", "[MASK]", "[sep]", "[CLS]x", "a[UNK]b", "[PAD", "[[SEP]]", "[SEP][SEP]"])
    if r < 0.90:
        return rnd.choice(["\t", "\n", "\x0b", "\x0c", "\x01", "\x1f", "\x7f", "  ", "\r\n", "\x00"])
    if r < 0.94:
        return rnd.choice(["naïve", "résumé", "β-cell", "中文", "Å", "x y", "​", "ﬁn", "İ"])
    return rnd.choice(SYLL)


def _same(a, b):
    assert a[1] == b[1], "abs_lens differ"
    assert a[0]["seq_lens"] == b[0]["seq_lens"], "seq_lens differ"
    for k in ("tokid_tt", "seg_tt", "attnmask_tt"):
        assert a[0][k].dtype == b[0][k].dtype and torch.equal(a[0][k], b[0][k]), k
    assert a[2].dtype == b[2].dtype and torch.equal(a[2], b[2]), "span tables differ"


@pytest.mark.parametrize("lower", [True, False])
def test_native_prep_matches_hf_tokenizer_fuzz(tmp_path, lower):
    from aspire_b200.consent import native_wordpiece, prepare_abstracts_fast, prepare_abstracts_native
    tok, words = _tokenizer(tmp_path, lower)
    wp = native_wordpiece(tok)
    assert wp is not None, "the plain BERT pipeline must be handled natively"
    rnd = random.Random(5 + lower)
    asserted = 0
    for _ in range(120):
        batch = [{"TITLE": " ".join(_random_word(rnd, words) for _ in range(rnd.randint(0, 14))),
                  "ABSTRACT": [" ".join(_random_word(rnd, words) for _ in range(rnd.randint(0, 90)))
                               for _ in range(rnd.randint(1, 30))]} for _ in range(rnd.randint(1, 6))]
        flat = [s for ex in batch for s in [ex["TITLE"] + " [SEP] "] + ex["ABSTRACT"]]
        ids, offs = wp.encode(flat)
        want = tok(flat, add_special_tokens=False)["input_ids"]
        for i, w in enumerate(want):
            assert ids[offs[i]:offs[i + 1]].tolist() == w, repr(flat[i])[:200]
        try:
            ref = prepare_abstracts_fast(batch, tok)
        except AssertionError:  # a document whose sentences all vanish: both front ends must refuse it
            with pytest.raises(AssertionError):
                prepare_abstracts_native(batch, tok)
            asserted += 1
            continue
        _same(ref, prepare_abstracts_native(batch, tok))
    assert asserted < 60


def test_native_prep_truncation_and_edges(tmp_path):
    from aspire_b200.consent import MAX_WORDPIECES, prepare_abstracts, prepare_abstracts_native, spans_from_token_idxs
    from aspire_b200.consent import prepare_abstracts_per_sentence
    tok, words = _tokenizer(tmp_path)
    w = words[0]
    cases = [
        [{"TITLE": "", "ABSTRACT": [w]}],                                                    # empty title
        [{"TITLE": w, "ABSTRACT": ["", w, ""]}],                                             # empty sentences keep a slot
        [{"TITLE": w, "ABSTRACT": [" ".join([w] * 300), " ".join([w] * 300), w]}],           # second sentence is cut, third dropped
        [{"TITLE": w, "ABSTRACT": [" ".join([w] * (MAX_WORDPIECES - 2)), w, w]}],            # budget met exactly, later ones dropped
        [{"TITLE": " ".join([w] * 10), "ABSTRACT": [w] * 40}, {"TITLE": w, "ABSTRACT": [w]}],  # ragged batch
        [{"TITLE": w, "ABSTRACT": ["naïve " + w, w + " résumé"]}],                           # Unicode sentences take the fallback
    ]
    for batch in cases:
        bb, al, idxs = prepare_abstracts_per_sentence(batch_abs=batch, pt_lm_tokenizer=tok)  # the reference's protocol
        db, dal, didxs = prepare_abstracts(batch_abs=batch, pt_lm_tokenizer=tok)             # the drop-in name (native)
        assert dal == al and didxs == idxs and db["seq_lens"] == bb["seq_lens"]
        for k in ("tokid_tt", "seg_tt", "attnmask_tt"):
            assert db[k].dtype == bb[k].dtype and torch.equal(db[k], bb[k])
        nb, nal, spans = prepare_abstracts_native(batch, tok)
        assert nal == al and nb["seq_lens"] == bb["seq_lens"]
        for k in ("tokid_tt", "seg_tt", "attnmask_tt"):
            assert torch.equal(nb[k], bb[k])
        assert torch.equal(spans, spans_from_token_idxs(idxs, max(al)))
        assert max(nb["seq_lens"]) <= MAX_WORDPIECES + 2
    with pytest.raises(AssertionError):  # the title alone fills the budget: no abstract sentence survives
        prepare_abstracts_native([{"TITLE": " ".join([w] * 600), "ABSTRACT": [w]}], tok)


def test_native_prep_falls_back_for_other_tokenizers():
    from aspire_b200.consent import native_wordpiece, prepare_abstracts_fast, prepare_abstracts_native
    toy = ref_shims.ToyTokenizer()
    assert native_wordpiece(toy) is None
    batch = [{"TITLE": "optimal transport", "ABSTRACT": ["we study alignment .", "graphs of papers"]}]
    _same(prepare_abstracts_fast(batch, toy), prepare_abstracts_native(batch, toy))


def test_native_prep_abi_errors():
    from aspire_b200 import _abi
    L = _abi.lib()
    assert not L.asp_wordpiece_create(None, None, 0, 1, 0, None, 0)
    assert b"asp_wordpiece_create" in L.asp_last_error()
    offs = np.array([0, 3, 5], dtype=np.int64)
    sents = np.array([0], dtype=np.int32)  # a document without even a title element
    out = np.zeros(4, dtype=np.int32)
    rc = L.asp_abstracts_plan(offs.ctypes.data, sents.ctypes.data, 1, 500, out.ctypes.data, out[2:].ctypes.data)
    assert rc != 0 and b"no title" in L.asp_last_error()
    # fill with a width smaller than the plan asks for must refuse, not overrun
    sents[0] = 2
    ids = np.arange(5, dtype=np.int32)
    tokid = np.zeros((1, 4), dtype=np.int64)
    spans = np.zeros((1, 1, 2), dtype=np.int32)
    rc = L.asp_abstracts_fill(ids.ctypes.data, offs.ctypes.data, sents.ctypes.data, 1, 500, 101, 102, ctypes.c_longlong(0), 4, 1,
                              tokid.ctypes.data, tokid.copy().ctypes.data, tokid.copy().ctypes.data, spans.ctypes.data)
    assert rc != 0 and b"does not fit" in L.asp_last_error()


@pytest.mark.parametrize("lower", [True, False])
def test_native_wordpiece_unicode_tables_fuzz(tmp_path, lower):
    """Non-ASCII text: the per-code-point tables are read off the tokenizer's own normaliser / pre-tokenizer, so accent
    stripping, lower-casing, CJK spacing, Unicode spaces and punctuation, zero-width and control characters must all
    come out as the Hugging Face tokenizer produces them; capital sigma (context-dependent lower-casing), characters
    beyond the BMP and whatever else the tables decline are answered by the Hugging Face tokenizer itself."""
    from aspire_b200.consent import native_wordpiece
    extra = ["β", "##β", "naive", "resume", "α", "δ", "中", "文", "±", "°", "–", "—", "“", "”", "’", "é", "É", "##é", "µ", "μ",
             "ß", "ı", "i", "##i", "ж", "Ж", "##ж", "×", "→", "σ", "ς", "##ς", "##σ", "ο", "φ", "##ο", "##φ"]
    tok, words = _tokenizer(tmp_path, lower, extra=extra)
    wp = native_wordpiece(tok)
    assert wp is not None
    rnd = random.Random(31 + lower)
    pools = [list(range(0x20, 0x7f)), list(range(0xa0, 0x180)), list(range(0x370, 0x400)), list(range(0x400, 0x460)),
             list(range(0x2000, 0x2070)), list(range(0x4e00, 0x4e40)), list(range(0x300, 0x330)), list(range(0x1e00, 0x1f00)),
             [0x3a3, 0x3c3, 0x3c2, 0x130, 0x131, 0xfb01, 0x1f600, 0x200b, 0xad, 0xfffd, 0xa0, 0x3000, 0x85, 0x2028, 0xac00,
              0xd55c, 0x10348]]
    sentences = []
    for _ in range(1500):
        parts = []
        for _ in range(rnd.randint(0, 40)):
            r = rnd.random()
            if r < 0.5:
                parts.append(rnd.choice(words))
            elif r < 0.6:
                parts.append(" ")
            elif r < 0.65:
                parts.append(rnd.choice(["[SEP]", "[MASK]", "naïve", "résumé", "RÉSUMÉ", "β-cell", "中文", "Ж", "ΣΟΦΟΣ", "σοφος"]))
            else:
                parts.append("".join(chr(rnd.choice(rnd.choice(pools))) for _ in range(rnd.randint(1, 4))))
            if rnd.random() < 0.7:
                parts.append(" ")
        sentences.append("".join(parts))
    ids, offs = wp.encode(sentences)  # one batch: native sentences and fallback sentences interleaved
    assert wp._unicode is True
    want = tok(sentences, add_special_tokens=False)["input_ids"]
    for i, w in enumerate(want):
        assert ids[offs[i]:offs[i + 1]].tolist() == w, repr(sentences[i])[:200]
    # the tables really are in use: a sentence of BMP characters is answered natively, one with an emoji is not
    L = wp._lib
    for text, expect in (("naïve β-cell “x” 中文", 0), ("smile \U0001F600", 1), ("ΣΟΦΟΣ", 1)):
        raw = text.encode("utf-8")
        o = np.array([0, len(raw)], dtype=np.int64)
        out, oo, fb = np.zeros(len(raw), dtype=np.int32), np.zeros(2, dtype=np.int64), np.zeros(1, dtype=np.uint8)
        assert L.asp_wordpiece_encode(wp._handle, raw, o.ctypes.data, 1, 100, 1, out.ctypes.data, oo.ctypes.data,
                                      fb.ctypes.data) == 0
        assert int(fb[0]) == expect, text


def test_context_ner_entity_positions_batched_equals_per_call(tmp_path):
    """AspireContextNER._get_ner_token_idxs (utils/models.py:661-682): all sentences and entities through one call of
    the native tokenizer against one ``tokenizer.tokenize`` call each, on the same fast tokenizer."""
    from aspire_b200.consent import prepare_abstracts_per_sentence
    from aspire_b200.similarity import AspireContextNER
    tok, words = _tokenizer(tmp_path)
    tok_slow, _ = _tokenizer(tmp_path)
    object.__setattr__(tok_slow, "_asp_native_wordpiece", False)  # pin the per-call path
    rnd = random.Random(77)
    batch = []
    for _ in range(12):
        abstract, entities = [], []
        for _ in range(rnd.randint(1, 40)):
            sent = [_random_word(rnd, words) for _ in range(rnd.randint(1, 40))]
            ents = []
            for _ in range(rnd.randint(0, 4)):
                if rnd.random() < 0.7 and len(sent) > 1:
                    a = rnd.randrange(len(sent))
                    ents.append(" ".join(sent[a:a + rnd.randint(1, 3)]))
                else:
                    ents.append(" ".join(rnd.choice(words) for _ in range(rnd.randint(1, 3))))  # usually absent
            abstract.append(" ".join(sent))
            entities.append(ents)
        batch.append({"TITLE": " ".join(rnd.choice(words) for _ in range(8)), "ABSTRACT": abstract, "ENTITIES": entities})
    _, _, sent_idxs = prepare_abstracts_per_sentence(batch_abs=batch, pt_lm_tokenizer=tok_slow)
    fast, slow = object.__new__(AspireContextNER), object.__new__(AspireContextNER)
    fast.tokenizer, slow.tokenizer = tok, tok_slow
    got, want = fast._get_ner_token_idxs(batch, sent_idxs), slow._get_ner_token_idxs(batch, sent_idxs)
    assert got == want
    assert sum(len(x) > 0 for doc in want for x in doc) > 20 and sum(len(x) == 0 for doc in want for x in doc) > 5


def test_native_prep_is_reentrant(tmp_path):
    """encode_stream prepares batches on worker threads: concurrent calls on ONE tokenizer (one shared native handle,
    Unicode tables installed lazily by whichever thread meets the first non-ASCII sentence) must each return what a
    serial call returns."""
    from concurrent.futures import ThreadPoolExecutor
    from aspire_b200.consent import prepare_abstracts_fast, prepare_abstracts_native
    tok, words = _tokenizer(tmp_path)
    rnd = random.Random(123)
    batches = [[{"TITLE": " ".join(_random_word(rnd, words) for _ in range(rnd.randint(1, 10))),
                 "ABSTRACT": [" ".join([rnd.choice(words)] + [_random_word(rnd, words) for _ in range(rnd.randint(0, 60))])
                              for _ in range(rnd.randint(1, 20))]} for _ in range(rnd.randint(1, 8))] for _ in range(48)]
    with ThreadPoolExecutor(max_workers=6) as pool:
        got = list(pool.map(lambda b: prepare_abstracts_native(b, tok), batches))
    for b, g in zip(batches, got):
        _same(prepare_abstracts_fast(b, tok), g)
