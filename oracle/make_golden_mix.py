"""TEST INFRASTRUCTURE ONLY -- golden vectors for caching_score's score mixing (sent_loss_prop / sentsup_loss_prop
scaling and the abs_loss_prop * (-||cls_q - cls_c||) term, src/learning/facetid_models/disent_models.py:294-307) from
the UNMODIFIED reference method, run in the build container through oracle/ref_shims.  Covers the four aggregation
heads the method is used with (l2wasserstein, l2max, l2top2, l2attention).  Writes tests/golden/caching_score_mix.npz."""
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shims  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "caching_score_mix.npz")


def main():
    ref_shims.install(bert_seed=0, bert_layers=2)
    from src.learning.facetid_models import pair_distances as pd_ref
    from src.learning.facetid_models import disent_models as dm_ref
    g = torch.Generator().manual_seed(2024)
    D = 128
    qrep = (0.3 * torch.randn(5, D, generator=g)).numpy().astype(np.float64)
    qcls = (0.3 * torch.randn(D, generator=g)).numpy().astype(np.float64)
    clens = [3, 8, 2, 5, 8, 2, 7, 4, 6, 11]
    creps = [(0.3 * torch.randn(n, D, generator=g)).numpy().astype(np.float64) for n in clens]
    ccls = [(0.3 * torch.randn(D, generator=g)).numpy().astype(np.float64) for _ in clens]
    qd = {"sent_reps": qrep, "doc_cls_reps": qcls}
    cds = [{"sent_reps": r, "doc_cls_reps": v} for r, v in zip(creps, ccls)]
    saved = {"q": qrep, "q_cls": qcls, "c_lens": np.array(clens), "c_cat": np.concatenate(creps, 0), "c_cls": np.stack(ccls)}
    fns = {"l2wasserstein": pd_ref.AllPairMaskedWasserstein({}).compute_distance,
           "l2max": pd_ref.allpair_masked_dist_l2max,
           "l2top2": pd_ref.allpair_masked_dist_l2topk,
           "l2attention": pd_ref.AllPairMaskedAttention({}).compute_distance}
    # (sent_loss_prop, sentsup_loss_prop or None, abs_loss_prop)
    mixes = {"a": (0.6, None, 0.4), "b": (0.0, 0.8, 0.25), "c": (1.0, None, 0.0)}
    for agg, fn in fns.items():
        for tag, (sp, ssp, ap) in mixes.items():
            fake = types.SimpleNamespace(score_agg_type=agg, dist_function=fn, sent_loss_prop=sp, abs_loss_prop=ap)
            if ssp is not None:
                fake.sentsup_loss_prop = ssp
            ret = dm_ref.WordSentAlignBiEnc.caching_score(fake, qd, cds)
            saved[f"{agg}_{tag}"] = np.asarray(ret["batch_scores"])
    saved["mix_a"], saved["mix_b"], saved["mix_c"] = (np.array([m[0], -1 if m[1] is None else m[1], m[2]]) for m in mixes.values())
    np.savez(OUT, **saved)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
