"""TEST INFRASTRUCTURE ONLY -- golden vectors for the l2top2 and attention heads from the UNMODIFIED reference
(src/learning/facetid_models/pair_distances.py:95-135,295-345), run in the build container through oracle/ref_shims.
Writes tests/golden/heads.npz."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shims  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "heads.npz")


def main():
    ref_shims.install(bert_seed=0, bert_layers=2)
    from collections import namedtuple
    from src.learning.facetid_models import pair_distances as pd_ref
    RL = namedtuple("RepLen", ["embed", "abs_lens"])
    g = torch.Generator().manual_seed(31337)
    B, Sq, Sc, D = 24, 9, 12, 192
    q = 0.3 * torch.randn(B, Sq, D, generator=g)
    c = 0.3 * torch.randn(B, Sc, D, generator=g) + 0.05
    ql = torch.randint(1, Sq + 1, (B,), generator=g).tolist()
    cl = torch.randint(1, Sc + 1, (B,), generator=g).tolist()
    ql[0], cl[0] = 1, 1          # a single valid sentence pair: the runner-up of l2top2 is a masked entry
    ql[1], cl[1] = Sq, Sc
    for b in range(B):
        q[b, ql[b]:] = 0
        c[b, cl[b]:] = 0
    qt = RL(embed=q.permute(0, 2, 1), abs_lens=ql)
    ct = RL(embed=c.permute(0, 2, 1), abs_lens=cl)
    top2_sims, top2_pair = pd_ref.allpair_masked_dist_l2topk(query=qt, cand=ct, return_pair_sims=True)
    top2_dist = pd_ref.allpair_masked_dist_l2topk(query=qt, cand=ct, return_pair_sims=False)
    out = {"q": q.numpy(), "c": c.numpy(), "q_lens": np.array(ql), "c_lens": np.array(cl),
           "top2_sims": top2_sims.numpy(), "top2_pair": top2_pair.numpy(), "top2_dist": top2_dist.numpy()}
    for temp in (1.0, 0.25):
        att = pd_ref.AllPairMaskedAttention({"cdatt_sm_temp": temp})
        doc_sims, (pair_sims, pair_softmax, masked_sims) = att.compute_distance(query=qt, cand=ct, return_pair_sims=True)
        doc_dists = att.compute_distance(query=qt, cand=ct, return_pair_sims=False)
        k = f"t{temp}"
        out.update({f"att_sims_{k}": doc_sims.numpy(), f"att_pair_{k}": pair_sims.numpy(),
                    f"att_softmax_{k}": pair_softmax.numpy(), f"att_masked_{k}": masked_sims.numpy(),
                    f"att_dists_{k}": doc_dists.numpy()})
    np.savez(OUT, **out)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
