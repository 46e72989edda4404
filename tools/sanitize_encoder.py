"""Encoder-only invocation for compute-sanitizer (the pipelined attention kernel over 3-4 tiles per CTA, LayerNorm on read)."""
import sys
import torch
sys.path.insert(0, ".")
from transformers import BertConfig, BertModel
from aspire_b200.encoder import B200BertEncoder

g = torch.Generator().manual_seed(0)
torch.manual_seed(0)
enc = B200BertEncoder(BertModel(BertConfig(vocab_size=2000, num_hidden_layers=1)).eval())
ids = torch.randint(5, 1999, (20, 150), generator=g)
h = enc.forward(ids, [150, 31] * 10, precision="bf16")
torch.cuda.synchronize()
assert torch.isfinite(h).all()
print("sanitize_encoder ok")
