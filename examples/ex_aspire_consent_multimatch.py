"""Alias module: ``from ex_aspire_consent_multimatch import AllPairMaskedWasserstein`` (src/evaluation/utils/models.py:3,
demo notebook cell 5) resolves to the B200-native implementation."""
from transformers import AutoModel, AutoTokenizer

from aspire_b200.consent import AspireConSent, prepare_abstracts, prepare_bert_sentences
from aspire_b200.distances import AllPairMaskedWasserstein

__all__ = ["AspireConSent", "AllPairMaskedWasserstein", "prepare_abstracts", "prepare_bert_sentences", "AutoModel",
           "AutoTokenizer"]
