"""aspire_b200 -- B200-native (sm_100a) implementation of Aspire's fine-grained document-similarity scoring path.

Public surface mirrors the reference (allenai/aspire): see distances.py, consent.py, similarity.py.
"""
from .distances import (AllPairMaskedWasserstein, AllPairMaskedAttention, allpair_masked_dist_l2max,
                        allpair_masked_dist_l2topk, allpair_masked_argmax_l2max, pair_heads,
                        rep_len_tup, RepLen, epsilon_schedule, ot_scores, ot_scores_allpairs, l2max_scores, l2max_allpairs, bbox_diameter)

__version__ = "0.1.0"
