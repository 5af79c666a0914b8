"""Bring-up diagnostics for the tcgen05 rank tier (run on the GPU box): raw contraction vs fp64 torch,
approximate scores vs the exact tier, band margins, counts.  Not a test; prints what it finds."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from complexhyperbolickge_b200 import ops  # noqa: E402


def run(rank, n_ent, b, seed=0, regime="trained"):
    g = torch.Generator().manual_seed(seed)
    std = float(np.sqrt(0.4 / (2 * rank))) if regime == "trained" else 1e-3
    ent = (torch.randn(n_ent, 2 * rank, generator=g) * std).cuda()
    q = (torch.randn(b, 2 * rank, generator=g) * std).cuda()
    bh = (torch.randn(b, generator=g) * 0.1).cuda()
    bt = (torch.randn(n_ent, generator=g) * 0.1).cuda()
    tails = torch.randint(0, n_ent, (b,), generator=g).cuda()
    qn, hn = ops.row_hnorm(rank, q), ops.row_hnorm(rank, ent)
    rows = ent[tails].contiguous()
    tgt = ops.target_scores(rank, q, qn, bh, rows, ops.row_hnorm(rank, rows), bt[tails].contiguous())
    S = ops.score_all(rank, q, qn, bh, ent, hn, bt)
    shadow = ops.entity_shadow(rank, ent, hn, bt)
    ws = ops.rank_mma_workspace(rank, b, ent.device)
    if True:
        os.environ["CHK_MMA_DUMP_RAW"] = "1"
        re, im, _ = ops.score_all_mma(rank, q, qn, bh, tgt, ent, hn, bt, shadow, ws)
        torch.cuda.synchronize()
        zr, zi = q[:, :rank].double(), q[:, rank:].double()
        wr, wi = ent[:, :rank].double(), ent[:, rank:].double()
        re_ref = zr @ wr.T + zi @ wi.T
        im_ref = zi @ wr.T - zr @ wi.T
        nz = q.double().norm(dim=1)[:, None] * ent.double().norm(dim=1)[None, :]
        e_re = ((re.double() - re_ref).abs() / nz).max().item()
        e_im = ((im.double() - im_ref).abs() / nz).max().item()
        print(f"[r={rank} N={n_ent} b={b} {regime}] raw contraction: max|re-ref|/(|z||w|) = {e_re:.3e}  im: {e_im:.3e}  nan={torch.isnan(re).sum().item()}")
    os.environ["CHK_MMA_DUMP_RAW"] = "0"
    St, band, counts = ops.score_all_mma(rank, q, qn, bh, tgt, ent, hn, bt, shadow, ws)
    n_list, ov = ops.rank_mma_status(ws)
    diff = (St.double() - S.double()).abs()
    pos = band > 0
    ratio = (diff[pos] / band[pos].double()).max().item() if pos.any() else 0.0
    exact0 = (diff[~pos] == 0).all().item() if (~pos).any() else True
    ref_counts = (S >= tgt[:, None]).sum(1)
    print(f"   scores: max|s~-s|={diff.max().item():.3e}  max band={band.max().item():.3e} median band={band.median().item():.3e} "
          f"max diff/band={ratio:.3f}  clamped-exact={exact0} ({(~pos).float().mean().item():.3f} of pairs)  "
          f"recheck list={n_list} ({n_list / (b * n_ent):.2e} of pairs) overflow={ov}  counts equal={torch.equal(counts, ref_counts)}")
    if not torch.equal(counts, ref_counts):
        bad = (counts != ref_counts).nonzero().flatten()[:8]
        print("   mismatching queries", bad.tolist(), counts[bad].tolist(), ref_counts[bad].tolist())


if __name__ == "__main__":
    torch.manual_seed(0)
    os.environ["CHK_MMA_CTA_PAIR"] = sys.argv[1] if len(sys.argv) > 1 else "1"
    print("CHK_MMA_CTA_PAIR =", os.environ["CHK_MMA_CTA_PAIR"])
    run(33, 1000, 150)
    run(33, 1000, 150, regime="init")
    run(65, 5000, 300)
    run(257, 20000, 500)
    run(257, 300, 1100)
