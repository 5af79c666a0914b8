"""GPU: the fused training step (train.FusedKGOptimizer: K1 + K3 + chk_nsloss + adjoints + row-sparse Adagrad, CUDA
graph) against the unfused KGOptimizer contract path (two model() calls + autograd + dense torch.optim) on identical
batches and injected negatives: same losses, same parameters after several steps."""
from argparse import Namespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(name, rank, dtype, n_ent=700, n_rel2=10, seed=0, bias="learn", multi_c=True):
    import complexhyperbolickge_b200 as chk
    from complexhyperbolickge_b200 import synthetic
    args = Namespace(sizes=(n_ent, n_rel2, n_ent), rank=rank, dropout=0, gamma=0, dtype=dtype, bias=bias, init_size=1e-3,
                     multi_c=multi_c)
    m = getattr(chk, name)(args).cuda()
    synthetic.trained_like_(m, seed)
    return m


def _feed(cls):
    class Fed(cls):
        def get_neg_samples(self, input_batch):
            return self._negs_static
    return Fed


@pytest.mark.parametrize("graph", [False, True])
@pytest.mark.parametrize("opt_name", ["Adagrad", "Adam"])
@pytest.mark.parametrize("name,rank,dtype,reg", [("FFTRotH", 33, "double", None), ("FFTRefH", 33, "float", None),
                                                ("FFTAttH", 17, "double", None), ("FFTRotH", 257, "float", None),
                                                ("FFTRefH", 33, "double", ("N3", 0.05)), ("FFTAttH", 33, "double", ("F2", 0.01)),
                                                ("FFTRotH", 33, "float", ("N3", 0.05))])
def test_fused_step_matches_contract_path(name, rank, dtype, reg, opt_name, graph):
    from complexhyperbolickge_b200 import optim
    from complexhyperbolickge_b200.optim import KGOptimizer
    N3 = (lambda _w: getattr(optim, reg[0])(reg[1])) if reg else optim.N3      # regulariser under test (weight 0 by default)
    from complexhyperbolickge_b200.train import FusedKGOptimizer
    B, neg, steps, n_ent, n_rel2 = 48, 21, 5, 700, 10
    a, b = _mk(name, rank, dtype), _mk(name, rank, dtype)
    b.load_state_dict(a.state_dict())
    mk_opt = (lambda ps: torch.optim.Adagrad(ps, lr=0.05)) if opt_name == "Adagrad" else (lambda ps: torch.optim.Adam(ps, lr=1e-3))
    ref = _feed(KGOptimizer)(a, N3(0.0), mk_opt(a.parameters()), B, 1, neg, False, verbose=False)
    fus = _feed(FusedKGOptimizer)(b, N3(0.0), mk_opt(b.parameters()), B, 1, neg, False, verbose=False, use_cuda_graph=graph)
    assert fus.fused and fus.sparse_adagrad == (opt_name == "Adagrad")
    g = torch.Generator().manual_seed(5)
    ref._negs_static = torch.zeros(B, neg, dtype=torch.int64, device="cuda")
    fus._negs_static = torch.zeros(B, neg, dtype=torch.int64, device="cuda")
    losses_ref = []
    fus._loss_sum.zero_()
    for i in range(steps):
        batch = torch.stack([torch.randint(0, n_ent, (B,), generator=g), torch.randint(0, n_rel2, (B,), generator=g),
                             torch.randint(0, n_ent, (B,), generator=g)], 1).cuda()
        if i == 2:
            batch[:7] = batch[7:14]                    # duplicate triples / rows inside a batch
        negs = torch.randint(0, n_ent, (B, neg), generator=g).cuda()
        ref._negs_static.copy_(negs)
        fus._negs_static.copy_(negs)
        l = ref.calculate_loss(batch)
        l.backward()
        ref.optimizer.step()
        ref.optimizer.zero_grad()
        losses_ref.append(l.item())
        fus.fused_step(batch)
    dbl = dtype == "double"
    got = fus._loss_sum.item() / steps
    assert abs(got - float(np.mean(losses_ref))) <= (1e-10 if dbl else 2e-5)
    for (k, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        scale = max(pa.abs().max().item(), 1e-30)
        err = (pa - pb).abs().max().item() / scale
        assert err <= (1e-8 if dbl else 5e-3), (k, err)
        assert pb.grad is None or pb.grad.abs().max().item() == 0 or not fus.sparse_adagrad, k   # grads cleared row-wise
    if opt_name == "Adagrad":                            # optimizer state is the torch optimizer's own, kept in sync
        for pa, pb in zip(a.parameters(), b.parameters()):
            sa, sb = ref.optimizer.state[pa]["sum"], fus.optimizer.state[pb]["sum"]
            assert (sa - sb).abs().max().item() <= (1e-12 if dbl else 1e-4) * max(sa.abs().max().item(), 1e-30)


def test_fused_epoch_runs_and_learns():
    from complexhyperbolickge_b200 import synthetic
    from complexhyperbolickge_b200.optim import N3
    from complexhyperbolickge_b200.train import FusedKGOptimizer
    m = _mk("FFTRotH", 33, "float", n_ent=2000, n_rel2=8)
    opt = FusedKGOptimizer(m, N3(0.0), torch.optim.Adagrad(m.parameters(), lr=0.1), 100, 1, 50, False, verbose=False)
    g = torch.Generator().manual_seed(0)
    ex = torch.stack([torch.randint(0, 2000, (1050,), generator=g), torch.randint(0, 8, (1050,), generator=g),
                      torch.randint(0, 2000, (1050,), generator=g)], 1)                   # ragged last batch (50)
    l0 = opt.epoch(ex)
    l1 = opt.epoch(ex)
    l2 = opt.epoch(ex)
    assert np.isfinite([l0, l1, l2]).all() and l2 < l0, (l0, l1, l2)
    # gradient accumulation falls back to the contract path; N3 with a non-zero weight stays fused
    opt2 = FusedKGOptimizer(m, N3(0.01), torch.optim.Adagrad(m.parameters(), lr=0.1), 100, 2, 50, False, verbose=False)
    assert not opt2.fused and np.isfinite(opt2.epoch(ex[:300]))
    opt3 = FusedKGOptimizer(m, N3(0.01), torch.optim.Adagrad(m.parameters(), lr=0.1), 100, 1, 50, False, verbose=False)
    assert opt3.fused and np.isfinite(opt3.epoch(ex[:300]))


@pytest.mark.parametrize("dtype,width", [(torch.float32, 66), (torch.float64, 130), (torch.float32, 1)])
def test_claim_gather_rows_sends_each_row_once(dtype, width):
    """Send side of the data-parallel sparse exchange (chk_claim_gather_rows): duplicates of a row carry zeros, the
    claimed rows are cleared in the dense gradient, untouched rows are left alone; claims last for one step."""
    from complexhyperbolickge_b200 import ops
    g = torch.Generator().manual_seed(0)
    N = 50
    ref = torch.randn(N, width, generator=g).to(dtype).cuda()
    grad = ref.clone()
    rows = torch.tensor([3, 7, 3, 49, 7, 7, 0], dtype=torch.int64).cuda()
    touched = torch.unique(rows)
    untouched = torch.ones(N, dtype=torch.bool, device="cuda")
    untouched[touched] = False
    stamp = torch.zeros(N, dtype=torch.int32, device="cuda")
    step = torch.ones((), dtype=torch.int32, device="cuda")

    def summed(out):
        total = torch.zeros_like(ref)
        total.index_add_(0, rows, out)
        return total

    out = ops.claim_gather_rows(grad, rows, stamp, step)
    assert out.shape == (rows.numel(), width)
    assert torch.equal(summed(out)[touched], ref[touched])            # every distinct row travels exactly once
    assert torch.equal(grad[untouched], ref[untouched])
    assert grad[touched].abs().max().item() == 0                      # claimed rows cleared locally
    grad.copy_(ref)
    out2 = ops.claim_gather_rows(grad, rows, stamp, step)             # same step: everything is already claimed
    assert out2.abs().max().item() == 0 and torch.equal(grad, ref)
    ops.step_counter_bump(step)
    out3 = ops.claim_gather_rows(grad, rows, stamp, step)             # next step: claims are fresh
    assert torch.equal(summed(out3)[touched], ref[touched])
