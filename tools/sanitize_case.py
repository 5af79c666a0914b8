"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / synccheck), run on the GPU box:
    compute-sanitizer --tool memcheck python tools/sanitize_case.py
Exercises every kernel family once at sizes the sanitizer finishes in a minute: rank_mma_kernel (single-CTA and, with
CHK_MMA_CTA_PAIR=1, the cta_group::2 variant), recheck_kernel, filter pass, the exact FMA tier, query_tpq_kernel + counting
sort, the lane-group K1 forward / adjoint, K3 (all modes), the fused training step (sampler, grouping, segment-reduce with
short / long segments, dense Adam apply) and chk_claim_gather_rows."""
import os
import sys
from argparse import Namespace

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import complexhyperbolickge_b200 as chk  # noqa: E402
from complexhyperbolickge_b200 import ops, synthetic  # noqa: E402
from complexhyperbolickge_b200.optim import N3  # noqa: E402
from complexhyperbolickge_b200.train import FusedKGOptimizer  # noqa: E402


def main():
    torch.manual_seed(0)
    N, R2 = 1100, 6
    for name, rank, dtype in (("FFTRotH", 33, "float"), ("FFTAttH", 17, "double")):
        args = Namespace(sizes=(N, R2, N), rank=rank, dropout=0, gamma=0, dtype=dtype, bias="learn", init_size=1e-3, multi_c=True)
        m = getattr(chk, name)(args).cuda()
        synthetic.trained_like_(m, 0)
        g = torch.Generator().manual_seed(1)
        ex = torch.stack([torch.randint(0, N, (300,), generator=g) % 30, torch.randint(0, R2, (300,), generator=g),
                          torch.randint(0, N, (300,), generator=g)], 1)
        filters = {}
        for h, r, t in ex.tolist():
            filters.setdefault((h, r), []).append(t)
        out = {}
        for algo in ("fma", "mma"):
            m.rank_algo = algo
            out[algo] = m.get_ranking(ex[:150], filters, batch_size=70)
        assert torch.equal(out["fma"], out["mma"])
        # K1 variants
        h, r = ex[:, 0].cuda(), ex[:, 1].cuda()
        a = (m.KIND, rank, True, m.entity.weight.detach(), m.rel.weight.detach(), m.rel_diag.weight.detach(),
             None if m._ctx_weight() is None else m._ctx_weight().detach(), m.c.weight.detach(), h, r)
        q0, _ = ops.query_fwd(*a, grouped=False)
        if dtype == "float":
            q1, _ = ops.query_fwd(*a, grouped=True)
            assert (q0 - q1).abs().max().item() < 1e-4
        # fused training steps: Adagrad (in-place segment-reduce; a hot head gives a long segment), Adam (dense apply), double_neg
        for opt_name, dn in (("Adagrad", False), ("Adam", True)):
            mk = (lambda ps: torch.optim.Adagrad(ps, lr=0.05)) if opt_name == "Adagrad" else (lambda ps: torch.optim.Adam(ps, lr=1e-3))
            opt = FusedKGOptimizer(m, N3(0.01), mk(m.parameters()), 100, 1, 12, dn, verbose=False, use_cuda_graph=False)
            hot = ex.clone()
            hot[:60, 2] = 5                                      # one tail named by 60 slots: the CTA (long-segment) path
            for i in range(3):
                opt.fused_step(hot[i * 100:(i + 1) * 100].cuda())
            assert torch.isfinite(opt._loss_sum).item()
        # claim-gather (C ABI, round-1 exchange send side)
        grad = torch.randn(N, 2 * rank, device="cuda", dtype=m.entity.weight.dtype)
        ops.claim_gather_rows(grad, ex[:, 0].cuda().contiguous(), torch.zeros(N, dtype=torch.int32, device="cuda"),
                              torch.ones((), dtype=torch.int32, device="cuda"))
    torch.cuda.synchronize()
    print("sanitize_case ok")


if __name__ == "__main__":
    main()
