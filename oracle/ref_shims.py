"""TEST INFRASTRUCTURE ONLY -- makes the UNMODIFIED reference importable in the build container.

Used by ``oracle/make_golden.py`` (and the ``needs_reference`` tests) to run the reference's own Python
on seeded inputs.  /root/reference does not exist on the GPU box, so nothing under ``-m gpu``,
``smoke()`` or ``bench.py`` imports this module.

What is stubbed, and why:
  * ``geomloss``           -> oracle.geomloss_ref (not installable offline; PARITY UNPINNED for that step)
  * ``h5py``               -> dict-backed ``File`` (only used for the encodings cache)
  * ``sentence_transformers``, ``matplotlib`` -> empty shells (imported at module scope, unused on the path)
  * ``AutoModel/AutoTokenizer.from_pretrained`` -> seeded random BERT-base + a deterministic whitespace
    word-piece tokenizer exposing the four members the path uses (no HF weights/vocab offline).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("ASPIRE_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "learning"))


class ToyTokenizer:
    """Deterministic stand-in for a BERT word-piece tokenizer (hash of lower-cased whitespace tokens).

    Long words are split into 4-character pieces so word-piece counts differ from word counts like a real
    vocabulary.  ids: 0=[PAD] 101=[CLS] 102=[SEP] 103=[MASK]; real tokens in [1000, vocab).
    """
    pad_token_id, cls_token_id, sep_token_id = 0, 101, 102

    def __init__(self, vocab_size=31116):
        self.vocab_size = vocab_size

    def tokenize(self, text):
        out = []
        for w in text.lower().split():
            if w == "[sep]":
                out.append("[SEP]")
                continue
            pieces = [w[i:i + 4] for i in range(0, len(w), 4)]
            out.extend([pieces[0]] + ["##" + p for p in pieces[1:]])
        return out

    def convert_tokens_to_ids(self, tokens):
        ids = []
        for t in tokens:
            if t == "[SEP]":
                ids.append(self.sep_token_id)
                continue
            h = 2166136261
            for ch in t.encode("utf8"):
                h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
            ids.append(1000 + h % (self.vocab_size - 1000))
        return ids

    def build_inputs_with_special_tokens(self, token_ids_0, token_ids_1=None):
        return [self.cls_token_id] + list(token_ids_0) + [self.sep_token_id]


def seeded_bert(seed=0, vocab_size=31116, num_hidden_layers=12):
    """Random-init BERT-base in eval mode (HF ``from_pretrained`` also returns eval mode)."""
    import torch
    from transformers import BertConfig, BertModel
    torch.manual_seed(seed)
    cfg = BertConfig(vocab_size=vocab_size, num_hidden_layers=num_hidden_layers)
    model = BertModel(cfg)
    model.eval()
    return model


def install(bert_seed=0, bert_layers=12):
    """Register stubs and sys.path entries; returns nothing.  Idempotent."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for p in (REFERENCE_ROOT, os.path.join(REFERENCE_ROOT, "examples"),
              os.path.join(REFERENCE_ROOT, "src", "pre_process")):
        if p not in sys.path:
            sys.path.append(p)
    from . import geomloss_ref
    sys.modules["geomloss"] = geomloss_ref

    if "h5py" not in sys.modules:
        h5 = types.ModuleType("h5py")

        class File(dict):
            def __init__(self, *a, **k):
                super().__init__()

            def create_dataset(self, name, data):
                self[name] = data

            def close(self):
                pass
        h5.File = File
        sys.modules["h5py"] = h5
    if "sentence_transformers" not in sys.modules:
        st = types.ModuleType("sentence_transformers")
        st.SentenceTransformer = type("SentenceTransformer", (), {})
        st.models = types.ModuleType("sentence_transformers.models")
        sys.modules["sentence_transformers"] = st
        sys.modules["sentence_transformers.models"] = st.models
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        mpl.pyplot = types.ModuleType("matplotlib.pyplot")
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = mpl.pyplot

    import transformers
    transformers.AutoModel.from_pretrained = staticmethod(
        lambda name, *a, **k: seeded_bert(bert_seed, num_hidden_layers=bert_layers))
    transformers.AutoTokenizer.from_pretrained = staticmethod(lambda name, *a, **k: ToyTokenizer())
