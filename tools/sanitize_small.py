"""Small fused-kernel invocations for compute-sanitizer (memcheck / racecheck / synccheck)."""
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
ot_scores(q, ql, c, cl, eps, want=("dual",))
l2max_scores(q, ql, c, cl)
pair_heads(q, ql, c, cl, want=("top2", "att"))
torch.cuda.synchronize()
print("sanitize_small ok")
