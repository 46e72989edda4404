"""Host-side throughput of the abstract-preparation front ends (documents/s, CPU only): the reference's per-sentence
protocol, one batched call into the Hugging Face (Rust) tokenizer, the library's native word-piece + sequence assembly,
and the drop-in ``prepare_abstracts`` (native batch + the reference's index lists).  Synthetic ~250-token abstracts over a synthetic word-piece vocabulary (no vocab files offline)."""
import os
import random
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transformers import BertTokenizerFast  # noqa: E402

from aspire_b200.consent import (prepare_abstracts, prepare_abstracts_fast, prepare_abstracts_native,  # noqa: E402
                                  prepare_abstracts_per_sentence)


def main():
    rnd = random.Random(1)
    syll = ["al", "ign", "trans", "port", "op", "ti", "mal", "bio", "med", "ic", "graph", "neur", "net", "work", "re",
            "triev", "sent", "ence", "pa", "per"]
    words = ["".join(rnd.choice(syll) for _ in range(rnd.randint(1, 3))) for _ in range(3000)]
    vocab = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"] + sorted(set(words)) + ["##" + s for s in syll] + [".", ","]
    vf = os.path.join(tempfile.mkdtemp(), "vocab.txt")
    with open(vf, "w") as fh:
        fh.write("\n".join(vocab) + "\n")
    tok = BertTokenizerFast(vocab_file=vf, do_lower_case=True)
    docs = [{"TITLE": " ".join(rnd.choice(words) for _ in range(10)),
             "ABSTRACT": [" ".join(rnd.choice(words) + rnd.choice(["", "", "", "s", "ing"]) for _ in range(rnd.randint(15, 35))) + " ."
                          for _ in range(rnd.randint(5, 11))]} for _ in range(4096)]
    print(f"host cores: {os.cpu_count()}")
    for name, fn, n_docs in (("per-sentence protocol (reference)", prepare_abstracts_per_sentence, 512),
                             ("prepare_abstracts_fast (HF batch)", prepare_abstracts_fast, 2048),
                             ("prepare_abstracts_native", prepare_abstracts_native, 4096),
                             ("prepare_abstracts (drop-in name)", prepare_abstracts, 4096)):
        for bs in (32, 128, 512):
            fn(batch_abs=docs[:bs], pt_lm_tokenizer=tok)
            t0 = time.perf_counter()
            for i in range(0, n_docs, bs):
                out = fn(batch_abs=docs[i:i + bs], pt_lm_tokenizer=tok)
            dt = time.perf_counter() - t0
            print(f"{name:36s} batch {bs:4d}: {n_docs / dt:9.0f} docs/s   (width {out[0]['tokid_tt'].shape[1]})", flush=True)


def pack_bench():
    """pack_pool (asp_pack_pool) against the per-candidate padding loop it replaces, 1 000 candidates of 3-10 sentences."""
    import numpy as np
    import torch

    from aspire_b200.similarity import pack_pool
    rng = np.random.default_rng(0)
    for kind in ("numpy fp32", "torch fp32", "numpy fp64"):
        encs = [rng.standard_normal((int(rng.integers(3, 11)), 768)).astype(np.float32) for _ in range(1000)]
        if kind == "torch fp32":
            encs = [torch.from_numpy(e) for e in encs]
        if kind == "numpy fp64":
            encs = [e.astype(np.float64) for e in encs]

        def loop():
            lens = [int(e.shape[0]) for e in encs]
            host = torch.zeros((len(encs), max(lens), 768), dtype=torch.float32)
            for j, e in enumerate(encs):
                host[j, :lens[j]] = torch.as_tensor(np.asarray(e), dtype=torch.float32)
            return host

        for name, fn in (("per-candidate loop", loop), ("pack_pool (native)", lambda: pack_pool(encs, "cpu"))):
            fn()
            t0 = time.perf_counter()
            for _ in range(10):
                fn()
            print(f"pool of 1000 x [3-10, 768] {kind:10s} {name:20s}: {(time.perf_counter() - t0) * 100:7.2f} ms", flush=True)


if __name__ == "__main__":
    main()
    pack_bench()
