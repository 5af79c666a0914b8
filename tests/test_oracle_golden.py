"""CPU: the oracle restatement against golden vectors produced by the reference itself
(oracle/make_golden.py).  Tolerances: fp64 1e-10 relative on values/grads (different but equivalent
summation orders, DFT-by-matrix vs pocketfft), fp32 2e-4 relative to the tensor's max magnitude with a
conditioning allowance near the clamps (SURVEY §7 hard part A)."""
import numpy as np
import pytest
import torch

from conftest import golden_files, load_case, oracle_params, filters_from_arrays
from oracle import chk_oracle as O


def _close(a, b, rtol, what):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-300)
    err = np.abs(a - b).max() / scale
    assert err <= rtol, f"{what}: max err / max|ref| = {err:.3e} > {rtol}"


@pytest.mark.parametrize("path", golden_files("step_"), ids=lambda p: p.split("/")[-1][5:-4])
def test_step_forward_and_grads(path):
    case = load_case(path)
    p = oracle_params(case)
    dbl = case["dtype"] == "double"
    batch, neg = torch.from_numpy(case["batch"]), torch.from_numpy(case["neg"])
    q, c = O.query_fwd(p, batch[:, 0:1], batch[:, 1:2])
    vt = 1e-10 if dbl else 3e-5
    _close(q.numpy(), case["q"], vt, "get_queries")
    _close(c.numpy().reshape(-1), np.broadcast_to(case["c"].reshape(-1, 1), (c.numel() // max(case["c"].size, 1), case["c"].size)).T.reshape(-1)
           if case["c"].size != c.numel() else case["c"].reshape(-1), vt, "curvature")
    bh = p.bh[batch[:, 0:1]]
    s_pos = O.score_pairs(p, q, bh, batch[:, 2:3])
    s_neg = O.score_pairs(p, q, bh, neg)
    st = 1e-9 if dbl else 2e-3          # fp32: x-1 cancellation, eps_mach/(x-1) conditioning
    _close(s_pos.numpy(), case["score_pos"], st, "positive scores")
    _close(s_neg.numpy(), case["score_neg"], st, "negative scores")
    sa = O.score_all(p, q, bh, chunk=16)
    _close(sa.numpy(), case["score_all"], st, "score_all")
    loss, grads = O.neg_sampling_loss(p, batch, neg)
    assert abs(loss.item() - float(case["loss"])) <= (1e-11 if dbl else 2e-5) * max(1.0, abs(float(case["loss"])))
    gt = 2e-8 if dbl else 2e-2
    if case["regime"] == "boundary" and not dbl:
        gt = 0.2                         # fp32 at |z|->1: gradients are O(1/eps) ill-conditioned
    for k, g in grads.items():
        ref = case["g_" + k]
        if np.abs(ref).max() == 0:
            assert np.abs(g.numpy()).max() <= 1e-30, k
        else:
            _close(g.numpy(), ref, gt, "grad " + k)


@pytest.mark.parametrize("path", golden_files("rank_"), ids=lambda p: p.split("/")[-1][5:-4])
def test_ranking(path):
    case = load_case(path)
    p = oracle_params(case)
    filters = filters_from_arrays(case)
    ex = torch.from_numpy(case["test"])
    ranks = O.get_ranking(p, ex, filters["rhs"], batch_size=37)
    q = torch.stack([ex[:, 2], ex[:, 1] + case["n_rel2"] // 2, ex[:, 0]], -1)
    ranks_l = O.get_ranking(p, q, filters["lhs"], batch_size=37)
    if case["dtype"] == "double":
        assert np.array_equal(ranks.numpy(), case["ranks_rhs"])
        assert np.array_equal(ranks_l.numpy(), case["ranks_lhs"])
    else:                                # fp32 near-ties: allow |d rank| <= 2 on < 3 % of queries
        for got, ref in ((ranks.numpy(), case["ranks_rhs"]), (ranks_l.numpy(), case["ranks_lhs"])):
            d = np.abs(got - ref)
            assert d.max() <= 2 and (d > 0).mean() < 0.03, (d.max(), (d > 0).mean())
    mr, mrr, hits = O.compute_metrics(p, ex, filters, case["n_rel2"], batch_size=64)
    tol = 1e-6 if case["dtype"] == "double" else 2e-2
    assert abs(mr["rhs"] - case["mr"][0]) <= tol * case["mr"][0]
    assert abs(mrr["lhs"] - case["mrr"][1]) <= tol


def _adam_step(p, g, m, v, t, lr, b1=0.9, b2=0.999, eps=1e-8):
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    p.addcdiv_(m / (1 - b1 ** t), (v / (1 - b2 ** t)).sqrt() + eps, value=-lr)


def _adagrad_step(p, g, acc, lr, eps=1e-10):
    acc.addcmul_(g, g)
    p.addcdiv_(g, acc.sqrt() + eps, value=-lr)


@pytest.mark.parametrize("path", golden_files("curve_"), ids=lambda p: p.split("/")[-1][6:-4])
def test_loss_curve(path):
    """N-step loss-curve parity: replay the reference's batches/negatives through oracle grads + a plain
    Adam/Adagrad (torch.optim semantics: dense updates of every row)."""
    case = load_case(path)
    p = oracle_params(case, "p0_")
    optim = case["regime"]                                   # meta[3] carries the optimiser name here
    lr = float(case["lr"])
    names = [k for k in ("entity", "rel", "rel_diag", "c", "bh", "bt", "context_vec") if getattr(p, k) is not None]
    st1 = {k: torch.zeros_like(getattr(p, k)) for k in names}
    st2 = {k: torch.zeros_like(getattr(p, k)) for k in names}
    off = 0
    losses = []
    nsteps = min(len(case["batch_lens"]), 40)
    for t in range(nsteps):
        L = int(case["batch_lens"][t])
        b = torch.from_numpy(case["batch_cat"][off:off + L])
        ng = torch.from_numpy(case["neg_cat"][off:off + L])
        off += L
        loss, grads = O.neg_sampling_loss(p, b, ng)
        losses.append(loss.item())
        for k in names:
            if optim == "Adam":
                _adam_step(getattr(p, k), grads[k], st1[k], st2[k], t + 1, lr)
            else:
                _adagrad_step(getattr(p, k), grads[k], st1[k], lr)
    ref = case["step_losses"][:nsteps]
    assert np.abs(np.array(losses) - ref).max() <= 1e-9, np.abs(np.array(losses) - ref).max()


def test_dft_definitions_match_numpy_fft():
    """O1/O2 pinned against an independent implementation (numpy pocketfft)."""
    rng = np.random.default_rng(0)
    for r in (9, 33, 65):
        x = rng.standard_normal((5, 2 * r))
        u = O.irfft_ortho(torch.from_numpy(x)).numpy()
        ref = np.fft.irfft(x[:, :r] + 1j * x[:, r:], norm="ortho")
        assert np.abs(u - ref).max() < 1e-13
        X = np.fft.rfft(u, norm="ortho")
        got = O.rfft_ortho(torch.from_numpy(u)).numpy()
        assert np.abs(got - np.concatenate([X.real, X.imag], -1)).max() < 1e-13
