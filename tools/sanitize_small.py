"""Small fused-kernel invocations for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os
import sys
import torch
sys.path.insert(0, ".")
from aspire_b200 import ot_scores, epsilon_schedule, _abi
from aspire_b200.distances import l2max_scores, pair_heads

g = torch.Generator().manual_seed(0)
eps = epsilon_schedule(30.0, 0.05, 0.9)[:12]
for _once in (0,):
    for (B, Sq, Sc, D, grp) in ((700, 10, 10, 768, 100), (333, 7, 9, 256, 1), (40, 10, 10, 768, 40)):
        nq = -(-B // grp)
        q = (0.3 * torch.randn(nq, Sq, D, generator=g)).cuda()
        c = (0.3 * torch.randn(B, Sc, D, generator=g)).cuda()
        ql = torch.randint(1, Sq + 1, (nq,), generator=g).int().cuda()
        cl = torch.randint(1, Sc + 1, (B,), generator=g).int().cuda()
        if B == 40:
            ql[:] = Sq
            cl[:] = Sc
        r = ot_scores(q, ql, c, cl, eps, q_group=grp, want=("dual", "primal", "plan"))
        torch.cuda.synchronize()
        assert torch.isfinite(r["dual"]).all()
q = (0.3 * torch.randn(50, 20, 128, generator=g)).cuda()
c = (0.3 * torch.randn(50, 27, 128, generator=g)).cuda()
ql = torch.randint(1, 21, (50,), generator=g).int().cuda()
cl = torch.randint(1, 28, (50,), generator=g).int().cuda()
# ASP_SAN_SKIP_VARLEN=1: synccheck reports "Missing init" for two of the var-len kernel's mbarriers (it does not model
# cp.async.mbarrier.arrive.noinc) and then aborts the process at the next launch; with the switch the rest of the
# script (every other kernel) still runs under that tool
if os.environ.get("ASP_SAN_SKIP_VARLEN") != "1":
    ot_scores(q, ql, c, cl, eps, want=("dual",))
    l2max_scores(q, ql, c, cl)
pair_heads(q, ql, c, cl, want=("top2", "att"))
torch.cuda.synchronize()
print("sanitize_small ok")

# ---- round 2 kernels: Q x C otAspire on tcgen05, single-pass top-k (+ exact fallback, packed merge), var-len kernel,
#      tensor-core pool prototype, tcgen05 attention ----
from aspire_b200 import ot_scores_allpairs
from aspire_b200.ranking import topk, topk_merge_packed
q = (0.3 * torch.randn(13, 10, 128, generator=g)).cuda()
c = (0.3 * torch.randn(37, 10, 128, generator=g)).cuda()
ql = torch.randint(1, 11, (13,), generator=g).int().cuda()
cl = torch.randint(1, 11, (37,), generator=g).int().cuda()
for w in (12, 8):
    _abi.set_option("oa_warps", w)
    sc = ot_scores_allpairs(q, ql, c, cl, eps)
    torch.cuda.synchronize()
    assert torch.isfinite(sc).all()
_abi.set_option("oa_warps", 12)
s = torch.randn(3, 20000, generator=g).cuda()
s[1] = 0.5                                   # a fully tied row: exact fallback kernel
ts, ti, tp = topk(s, 100, base_id=7, negate=True, want_packed=True)
topk_merge_packed(torch.stack([tp, tp]).contiguous(), 100)
torch.cuda.synchronize()
_abi.set_option("ot_fused_tc", 1)
B, grp = 9600, 4800
q = (0.3 * torch.randn(2, 10, 128, generator=g)).cuda()
c = (0.3 * torch.randn(B, 10, 128, generator=g)).cuda()
r = ot_scores(q, torch.tensor([10, 7]).int().cuda(), c, torch.randint(1, 11, (B,), generator=g).int().cuda(), eps, q_group=grp)
torch.cuda.synchronize()
_abi.set_option("ot_fused_tc", 0)
assert torch.isfinite(r["dual"]).all()
from transformers import BertConfig, BertModel
from aspire_b200.encoder import B200BertEncoder
torch.manual_seed(0)
enc = B200BertEncoder(BertModel(BertConfig(vocab_size=2000, num_hidden_layers=1)).eval())
# 20 documents x 12 heads x 2 query blocks = 480 tiles: every CTA of the pipelined attention kernel walks 3-4 of them
ids = torch.randint(5, 1999, (20, 150), generator=g)
h = enc.forward(ids, [150, 31] * 10, precision="bf16")
torch.cuda.synchronize()
assert torch.isfinite(h).all()
print("sanitize_small round-2 kernels ok")
