"""TEST INFRASTRUCTURE ONLY -- golden data for the entity-augmented models from the UNMODIFIED reference
(src/evaluation/utils/models.py: AspireNER :211-233, AspireContextNER :607-734, AspireConSenContextual :413-507), run in
the build container through oracle/ref_shims (seeded 2-layer BERT, ToyTokenizer).  Writes tests/golden/ner.npz/.json."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shims  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

PAPERS = [
    {"TITLE": "Optimal transport for document similarity",
     "ABSTRACT": ["We study optimal transport between sentence sets of scientific papers .",
                  "A sinkhorn solver aligns the sentences of two abstracts .",
                  "Experiments on faceted retrieval benchmarks show consistent gains ."],
     "ENTITIES": [["optimal transport", "sentence sets"], ["sinkhorn solver", "not in this sentence"],
                  ["faceted retrieval benchmarks"]],
     "FACETS": ["background_label", "method_label", "result_label"]},
    {"TITLE": "A short note",
     "ABSTRACT": ["Entity free sentence .", "Graph neural networks help molecule property prediction ."],
     "ENTITIES": [[], ["graph neural networks", "molecule property prediction"]],
     "FACETS": ["objective_label", "method_label"]},
]


def main():
    ref_shims.install(bert_seed=0, bert_layers=2)
    import transformers
    transformers.AutoTokenizer.from_pretrained = staticmethod(lambda name, *a, **k: ref_shims.ToyTokenizer())
    from src.evaluation.utils import models as ev_models
    ctx = ev_models.AspireContextNER(name="aspire_context_ner_compsci", encoding_type="sentence-entity")
    _, abs_lens, sent_idxs, ner_idxs = ctx._preprocess_input(PAPERS)
    with torch.no_grad():
        reps = ctx.encode(PAPERS)
    faceted = {}
    for i, paper in enumerate(PAPERS):
        for facet in ("background", "method", "result"):
            enc = torch.arange(reps[i].shape[0], dtype=torch.float32)[:, None].repeat(1, 2)  # row ids as "encoding"
            faceted[f"{i}_{facet}"] = ctx.get_faceted_encoding(enc, facet, paper)[:, 0].tolist()
    ner = ev_models.AspireNER(name="aspire_ner_compsci", encoding_type="sentence-entity")
    appended = ner._append_entities(PAPERS)
    with torch.no_grad():
        ner_reps = ner.encode(PAPERS)
    with open(os.path.join(GOLDEN, "ner.json"), "w") as fh:
        json.dump({"papers": PAPERS, "abs_lens": abs_lens, "ner_token_idxs": ner_idxs, "faceted_rows": faceted,
                   "appended": appended,
                   "sublist": [[list("abcabd"), list("abd"), ev_models.AspireContextNER.find_sublist_range(list("abcabd"), list("abd"))],
                               [list("abc"), list("cd"), ev_models.AspireContextNER.find_sublist_range(list("abc"), list("cd"))],
                               [list("abc"), [], ev_models.AspireContextNER.find_sublist_range(list("abc"), [])]]}, fh)
    np.savez(os.path.join(GOLDEN, "ner.npz"), **{f"ctx_{i}": r.numpy() for i, r in enumerate(reps)},
             **{f"ner_{i}": r.numpy() for i, r in enumerate(ner_reps)})
    print("wrote ner.json / ner.npz", [tuple(r.shape) for r in reps], [tuple(r.shape) for r in ner_reps], ner_idxs)


if __name__ == "__main__":
    main()
