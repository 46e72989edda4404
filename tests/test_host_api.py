"""CPU: host-side logic of the drop-in boundary -- batch preparation, API surface, sharding arithmetic."""
import inspect
import json
import os

import numpy as np
import pytest
import torch

from oracle import ref_shims

from conftest import GOLDEN


@pytest.fixture(scope="module")
def prep_golden():
    with open(os.path.join(GOLDEN, "prepare_abstracts.json")) as fh:
        return json.load(fh)


def _batch(prep_golden):
    from oracle.make_golden import README_ABSTRACTS
    return README_ABSTRACTS + [prep_golden["long_doc"]]


def test_prepare_abstracts_matches_reference_output(prep_golden):
    from aspire_b200.consent import prepare_abstracts
    bert_batch, abs_lens, sent_tok_idxs = prepare_abstracts(batch_abs=_batch(prep_golden),
                                                            pt_lm_tokenizer=ref_shims.ToyTokenizer())
    assert abs_lens == prep_golden["abs_lens"]
    assert bert_batch["seq_lens"] == prep_golden["seq_lens"]
    assert bert_batch["tokid_tt"].tolist() == prep_golden["tokid"]
    assert bert_batch["seg_tt"].tolist() == prep_golden["seg"]
    assert bert_batch["attnmask_tt"].tolist() == prep_golden["attn"]
    assert bert_batch["tokid_tt"].dtype == torch.int64
    spans = [[[s[0], s[-1] + 1] for s in doc] for doc in sent_tok_idxs]
    assert spans == prep_golden["spans"]
    for doc in sent_tok_idxs:
        for s in doc:
            assert s == list(range(s[0], s[-1] + 1))  # contiguous
    # truncation: 500 word-pieces + [CLS]/[SEP]; the crossing sentence is cut, later ones dropped
    assert max(bert_batch["seq_lens"]) == 502
    assert abs_lens[2] < 12 and sent_tok_idxs[2][-1][-1] == 500


def test_prepare_abstracts_title_fills_budget_asserts():
    from aspire_b200.consent import prepare_abstracts
    doc = {"TITLE": "t " * 600, "ABSTRACT": ["never reached"]}
    with pytest.raises(AssertionError):
        prepare_abstracts(batch_abs=[doc], pt_lm_tokenizer=ref_shims.ToyTokenizer())


def test_spans_from_token_idxs():
    from aspire_b200.consent import spans_from_token_idxs
    sp = spans_from_token_idxs([[[3, 4, 5], [6]], [[2, 3]]], 3)
    assert sp.dtype == torch.int32 and sp.shape == (2, 3, 2)
    assert sp.tolist() == [[[3, 6], [6, 7], [-1, -1]], [[2, 4], [-1, -1], [-1, -1]]]
    with pytest.raises(NotImplementedError):
        spans_from_token_idxs([[[3, 5]]], 1)


def test_api_surface_matches_reference_signatures():
    """Keyword names are part of the API (README.md:88-92, utils/models.py:201-207, notebook cell 14)."""
    import aspire_b200
    from aspire_b200 import consent, distances, similarity
    assert list(inspect.signature(consent.prepare_abstracts).parameters) == ["batch_abs", "pt_lm_tokenizer"]
    assert list(inspect.signature(consent.prepare_bert_sentences).parameters) == ["batch_doc_sents", "tokenizer"]
    assert list(inspect.signature(consent.AspireConSent.forward).parameters) == \
        ["self", "bert_batch", "abs_lens", "sent_tok_idxs"]
    assert list(inspect.signature(distances.AllPairMaskedWasserstein.compute_distance).parameters) == \
        ["self", "query", "cand", "return_pair_sims"]
    assert list(inspect.signature(distances.allpair_masked_dist_l2max).parameters) == \
        ["query", "cand", "return_pair_sims"]
    assert distances.rep_len_tup._fields == ("embed", "abs_lens")
    w = distances.AllPairMaskedWasserstein({})
    assert (w.geoml_blur, w.geoml_scaling, w.geoml_reach, w.sent_sm_temp) == (0.05, 0.9, None, 1.0)
    assert set(similarity.SimilarityModel.__abstractmethods__) == {"encode", "get_similarity"}
    with pytest.raises(NotImplementedError):
        similarity.get_model("no_such_model")
    with pytest.raises(AssertionError):
        type("M", (similarity.SimilarityModel,), {"encode": lambda s, b: None,
                                                  "get_similarity": lambda s, x, y: 0})(name="m", encoding_type="x")
    # both import spellings of the reference resolve to the drop-in
    from examples.ex_aspire_consent import AspireConSent as A1
    from examples.ex_aspire_consent_multimatch import AllPairMaskedWasserstein as W1
    assert A1 is consent.AspireConSent and W1 is distances.AllPairMaskedWasserstein


def test_epsilon_schedule_equals_oracle():
    from aspire_b200 import epsilon_schedule
    from oracle import geomloss_ref as gr
    for diam, blur, sc in [(43.117, 0.05, 0.9), (166.4, 0.05, 0.9), (12.8, 0.1, 0.9), (52.4, 1.0, 0.8), (0.03, 0.05, 0.9)]:
        assert epsilon_schedule(diam, blur, sc) == gr.epsilon_schedule(1, diam, blur, sc)


def test_faceted_encoding_and_cache(tmp_path):
    from aspire_b200.similarity import SimilarityModel

    class Dummy(SimilarityModel):
        def encode(self, batch_papers):
            return [torch.full((len(p["ABSTRACT"]), 4), float(len(p["TITLE"]))) for p in batch_papers]

        def get_similarity(self, x, y):
            return -float((x.mean() - y.mean()).abs())

    m = Dummy(name="d", encoding_type="sentence", batch_size=2)
    enc = torch.arange(12.).view(4, 3)
    paper = {"FACETS": ["objective_label", "method_label", "result_label", "method_label"]}
    assert m.get_faceted_encoding(enc, "method", paper).tolist() == enc[[1, 3]].tolist()
    assert m.get_faceted_encoding(enc, "background", paper).tolist() == enc[[0]].tolist()
    m2 = Dummy(name="d", encoding_type="sentence-entity")
    paper2 = dict(paper, ENTITIES=[["a"], ["b", "c"], [], ["d"]])
    enc2 = torch.arange(8.).view(8, 1)
    assert m2.get_faceted_encoding(enc2, "method", paper2).view(-1).tolist() == [1., 3., 5., 6., 7.]

    class DS:
        data = {"p%d" % i: {"TITLE": "t" * (i + 1), "ABSTRACT": ["s"] * (i + 1)} for i in range(5)}

        def get(self, pid):
            return self.data[pid]
    fn = str(tmp_path / "enc.npz")
    m.set_encodings_cache(fn)
    encs = m.get_encoding(["p0", "p3", "p4"], DS())
    assert set(encs) == {"p0", "p3", "p4"} and encs["p3"].shape == (4, 4)
    m.cache.close()
    m3 = Dummy(name="d", encoding_type="sentence")
    m3.set_encodings_cache(fn)
    assert set(m3.cache.keys()) == {"p0", "p3", "p4"}
    again = m3.get_encoding(["p3"], DS())
    assert torch.equal(again["p3"], encs["p3"])


def test_encodings_cache_h5_style_name_roundtrip_and_formats(tmp_path, monkeypatch):
    """The reference names its cache ``encodings.h5`` (utils/utils.py:53-60).  Without h5py the cache must come back
    from the SAME resolved path it was saved to, survive paper ids that are not identifiers, write only when dirty, and
    refuse (not silently re-encode) a real HDF5 file; with h5py importable it must open the reference's own format."""
    import sys
    import types
    import numpy as np
    from aspire_b200 import _abi
    from aspire_b200.similarity import EncodingsCache
    monkeypatch.setitem(sys.modules, "h5py", None)  # "import h5py" raises ImportError
    fn = str(tmp_path / "encodings.h5")
    c = EncodingsCache(fn)
    c.create_dataset(name="file", data=torch.ones(2, 3))       # np.savez's own keyword
    c.create_dataset(name="10.1/abc-def", data=np.arange(6.).reshape(3, 2))
    c.close()
    assert os.path.exists(fn + ".npz") and not os.path.exists(fn)
    c2 = EncodingsCache(fn)
    assert set(c2.keys()) == {"file", "10.1/abc-def"} and "file" in c2 and len(c2) == 2
    assert np.array_equal(np.array(c2.get("10.1/abc-def")), np.arange(6.).reshape(3, 2))
    stamp = os.path.getmtime(fn + ".npz")
    c2.close()                                                 # nothing changed: not rewritten
    assert os.path.getmtime(fn + ".npz") == stamp
    with open(str(tmp_path / "real.h5"), "wb") as fh:
        fh.write(b"\x89HDF\r\n\x1a\n" + b"\0" * 64)
    with pytest.raises(_abi.AspireB200Error):
        EncodingsCache(str(tmp_path / "real.h5"))
    # with h5py importable: h5py.File(filename, 'a'), 'w' when that fails, one dataset per paper id
    calls = []

    class FakeFile(dict):
        def __init__(self, name, mode):
            calls.append((name, mode))
            if mode == 'a' and name.endswith("corrupt.h5"):
                raise OSError("bad superblock")
            super().__init__()

        def create_dataset(self, name, data):
            self[name] = np.asarray(data)

        def close(self):
            calls.append("closed")
    monkeypatch.setitem(sys.modules, "h5py", types.SimpleNamespace(File=FakeFile))
    c3 = EncodingsCache(str(tmp_path / "enc.h5"))
    c3.create_dataset(name="p1", data=torch.zeros(1, 2))
    assert "p1" in c3 and np.array(c3.get("p1")).shape == (1, 2)
    c3.close()
    EncodingsCache(str(tmp_path / "corrupt.h5"))
    assert calls == [(str(tmp_path / "enc.h5"), 'a'), "closed", (str(tmp_path / "corrupt.h5"), 'a'),
                     (str(tmp_path / "corrupt.h5"), 'w')]


def test_shard_bounds_cover_pool():
    from aspire_b200.ranking import shard_bounds
    for n, w in [(1000000, 8), (1003, 8), (5, 8), (100000, 3)]:
        cuts = [shard_bounds(n, w, r) for r in range(w)]
        assert cuts[0][0] == 0 and cuts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(cuts[:-1], cuts[1:]))
        sizes = [e - s for s, e in cuts]
        assert max(sizes) - min(sizes) <= 1


def test_host_merge_order():
    from aspire_b200.ranking import host_merge
    s = torch.tensor([[1.0, 3.0, 3.0, -1.0, 2.0, 0.0]])
    i = torch.tensor([[10, 7, 4, -1, 5, 6]])
    ms, mi = host_merge(s, i, 3)
    assert mi.tolist() == [[4, 7, 5]] and ms.tolist() == [[3.0, 3.0, 2.0]]


def _ner_model(cls_name):
    """An entity-augmented model with the offline stand-ins (seeded 2-layer BERT, ToyTokenizer)."""
    import transformers
    from aspire_b200 import similarity
    om, ot = transformers.AutoModel.from_pretrained, transformers.AutoTokenizer.from_pretrained
    transformers.AutoModel.from_pretrained = staticmethod(lambda name, *a, **k: ref_shims.seeded_bert(0, num_hidden_layers=2))
    transformers.AutoTokenizer.from_pretrained = staticmethod(lambda name, *a, **k: ref_shims.ToyTokenizer())
    try:
        return similarity.get_model(cls_name)
    finally:
        transformers.AutoModel.from_pretrained, transformers.AutoTokenizer.from_pretrained = om, ot


def test_ner_models_host_logic_vs_reference_golden():
    """Entity token spans, facet filtering and entity appending vs the unmodified reference (oracle/make_golden_ner.py)."""
    from aspire_b200.similarity import AspireContextNER, AspireNER
    with open(os.path.join(GOLDEN, "ner.json")) as fh:
        z = json.load(fh)
    for sup, sub, want in z["sublist"]:
        assert AspireContextNER.find_sublist_range(sup, sub) == want
    assert AspireNER._append_entities(z["papers"]) == z["appended"]
    m = _ner_model("aspire_context_ner_compsci")
    assert type(m).__name__ == "AspireContextNER" and m.encoding_type == "sentence-entity"
    _, abs_lens, _, ner_idxs = m._preprocess_input(z["papers"])
    assert abs_lens == z["abs_lens"] and ner_idxs == z["ner_token_idxs"]
    for i, paper in enumerate(z["papers"]):
        n_rows = abs_lens[i] + sum(1 for x in ner_idxs[i] if x)
        enc = torch.arange(n_rows, dtype=torch.float32)[:, None].repeat(1, 2)
        for facet in ("background", "method", "result"):
            assert m.get_faceted_encoding(enc, facet, paper)[:, 0].tolist() == z["faceted_rows"][f"{i}_{facet}"]
    assert type(_ner_model("aspire_ner_biomed")).__name__ == "AspireNER"


@pytest.mark.needs_reference
def test_prepare_abstracts_fuzz_vs_live_reference():
    """Seeded random abstracts (1-40 sentences, 1-80 words, some far past the 500-piece budget) through the UNMODIFIED
    reference's prepare_abstracts (examples/ex_aspire_consent.py:185-212) and through ours: every output identical."""
    import random
    ref_shims.install(bert_seed=0, bert_layers=2)
    import ex_aspire_consent as ex_ref
    from aspire_b200.consent import prepare_abstracts
    rnd = random.Random(1234)
    vocab = ["alpha", "beta", "gamma", "transformer", "optimal", "transport", "sentence", "x", "biomedical",
             "representation", "learning", "a", "of", "the", "sinkhorn"]
    tok = ref_shims.ToyTokenizer()
    for trial in range(25):
        batch = []
        for _ in range(rnd.randint(1, 6)):
            n_sents = rnd.randint(1, 40)
            batch.append({"TITLE": " ".join(rnd.choice(vocab) for _ in range(rnd.randint(1, 30))),
                          "ABSTRACT": [" ".join(rnd.choice(vocab) for _ in range(rnd.randint(1, 80))) for _ in range(n_sents)]})
        try:
            want = ex_ref.prepare_abstracts(batch_abs=batch, pt_lm_tokenizer=tok)
        except AssertionError:
            with pytest.raises(AssertionError):
                prepare_abstracts(batch_abs=batch, pt_lm_tokenizer=tok)
            continue
        got = prepare_abstracts(batch_abs=batch, pt_lm_tokenizer=tok)
        assert got[1] == want[1] and got[2] == want[2]
        assert got[0]["seq_lens"] == want[0]["seq_lens"]
        for k in ("tokid_tt", "seg_tt", "attnmask_tt"):
            assert torch.equal(got[0][k], want[0][k]) and got[0][k].dtype == want[0][k].dtype


def test_prepare_abstracts_fast_equals_per_sentence_protocol(tmp_path):
    """The batch-tokenising front end (one Rust tokenizer call per batch, spans as an int32 table) against the
    reference-shaped prepare_abstracts on the SAME BertTokenizerFast (built offline from a small word-piece vocab),
    and against the ToyTokenizer fallback path."""
    import random
    from transformers import BertTokenizerFast
    from aspire_b200.consent import prepare_abstracts_per_sentence as prepare_abstracts  # the reference's protocol
    from aspire_b200.consent import prepare_abstracts_fast, spans_from_token_idxs
    words = ["optimal", "transport", "sentence", "paper", "graph", "neural", "network", "the", "of", "a", "we", "study",
             "align", "##ment", "##s", "##ing", "bio", "##medical", "retrieval", ".", ","]
    vocab = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"] + words
    vf = tmp_path / "vocab.txt"
    vf.write_text("\n".join(vocab) + "\n")
    fast_tok = BertTokenizerFast(vocab_file=str(vf), do_lower_case=True)
    assert fast_tok.is_fast
    rnd = random.Random(7)
    plain = ["optimal", "transport", "alignments", "aligning", "biomedical", "papers", "graphs", "the", "of", "we",
             "study", "retrieval", "unknownword", ".", ","]
    for tok in (fast_tok, ref_shims.ToyTokenizer()):
        for _ in range(10):
            batch = [{"TITLE": " ".join(rnd.choice(plain) for _ in range(rnd.randint(1, 12))),
                      "ABSTRACT": [" ".join(rnd.choice(plain) for _ in range(rnd.randint(1, 120)))
                                   for _ in range(rnd.randint(1, 25))]} for _ in range(rnd.randint(1, 5))]
            bb, al, idxs = prepare_abstracts(batch_abs=batch, pt_lm_tokenizer=tok)
            fb, fal, spans = prepare_abstracts_fast(batch_abs=batch, pt_lm_tokenizer=tok)
            assert fal == al and fb["seq_lens"] == bb["seq_lens"]
            for k in ("tokid_tt", "seg_tt", "attnmask_tt"):
                assert torch.equal(fb[k], bb[k])
            assert torch.equal(spans, spans_from_token_idxs(idxs, max(al)))


def test_pack_pool_native_gather_matches_padding_loop():
    """pack_pool (asp_pack_pool behind it) against the per-candidate padding loop it replaces
    (disent_models.py:274-290): numpy fp32 / fp64 / non-contiguous inputs and torch tensors, ragged lengths, an explicit
    max_sents, an empty encoding; bad shapes are refused."""
    from aspire_b200.similarity import pack_pool
    rng = np.random.default_rng(3)
    encs = []
    for j in range(257):
        s = int(rng.integers(0 if j == 5 else 1, 11))
        a = rng.standard_normal((s, 768))
        kind = j % 4
        if kind == 0:
            encs.append(a.astype(np.float32))
        elif kind == 1:
            encs.append(a)                                                   # float64, as np.load of a reps file gives
        elif kind == 2:
            encs.append(torch.from_numpy(a.astype(np.float32)))
        else:
            encs.append(np.asfortranarray(a.astype(np.float32)))             # not C-contiguous
    for max_sents in (None, 12):
        got, lens = pack_pool(encs, "cpu", max_sents=max_sents)
        smax = max_sents or max(int(e.shape[0]) for e in encs)
        want = torch.zeros((len(encs), smax, 768), dtype=torch.float32)
        for j, e in enumerate(encs):
            want[j, :e.shape[0]] = torch.as_tensor(np.asarray(e), dtype=torch.float32)
        assert got.shape == want.shape and torch.equal(got, want)
        assert lens.dtype == torch.int32 and lens.tolist() == [int(e.shape[0]) for e in encs]
        # the branch taken for GPU-resident encodings (torch ops only) must produce the same buffer
        from aspire_b200.similarity import _pad_tensors
        as_tensors = [torch.as_tensor(np.asarray(e), dtype=torch.float32) for e in encs]
        assert torch.equal(_pad_tensors(as_tensors, "cpu", smax, 768), want)
    with pytest.raises(Exception):
        pack_pool(encs, "cpu", max_sents=4)                                  # an encoding longer than max_sents
    with pytest.raises(ValueError):
        pack_pool([np.zeros((3, 768), np.float32), np.zeros((2, 700), np.float32)], "cpu")
