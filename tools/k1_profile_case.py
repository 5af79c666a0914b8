import sys, os, torch, numpy as np
sys.path.insert(0, os.getcwd())
from complexhyperbolickge_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
rank, nq, n_ent, n_rel2 = 33, 1 << 20, 1 << 20, 474
n = 2 * (rank - 1)
ent = torch.randn(n_ent, 2 * rank, generator=g, device="cuda") * float(np.sqrt(0.4 / (2 * rank)))
rel = torch.randn(n_rel2, 2 * n, generator=g, device="cuda") * 0.05
rd = torch.rand(n_rel2, n, generator=g, device="cuda") * 2 - 1
c = torch.rand(n_rel2, 1, generator=g, device="cuda") + 0.5
h = torch.randint(0, n_ent, (nq,), generator=g, device="cuda")
r = torch.sort(torch.randint(0, n_rel2, (nq,), generator=g, device="cuda"))[0]
for _ in range(2):
    q, _c = ops.query_fwd(ops.CHK_ROT, rank, True, ent, rel, rd, None, c, h, r)
gq = torch.randn(nq, 2 * rank, generator=g, device="cuda")
ops.query_bwd(ops.CHK_ROT, rank, True, ent, rel, rd, None, c, h, r, gq)
torch.cuda.synchronize()
print("ok")
