"""Alias module: ``from ex_aspire_consent import AspireConSent, prepare_abstracts`` (src/evaluation/utils/models.py:2,
README.md:60) resolves to the B200-native implementation."""
from transformers import AutoModel, AutoTokenizer  # re-exported like the reference module does

from aspire_b200.consent import AspireConSent, prepare_abstracts, prepare_bert_sentences

__all__ = ["AspireConSent", "prepare_abstracts", "prepare_bert_sentences", "AutoModel", "AutoTokenizer"]
