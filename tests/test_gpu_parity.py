"""GPU parity (pytest -m gpu): the CUDA path, called through the C ABI (ops -> libchk_b200.so), against
(1) golden vectors produced by the reference itself and (2) the CPU oracle on seeded inputs.

Stated tolerances (north_star: 1e-5 relative in fp32, 1e-12 in --dtype double, plus the conditioning term of SURVEY §7
hard part A):
  scores  : PER ELEMENT  |got - truth| <= K_SCORE * eps_machine * unit,  unit = 2(x+1)(1 + 1/|zn| + 1/|wn|) + |s| + |bh| + |bt|
            (tests/parity_units.py; truth = the oracle in fp64 on the same weights with the model's clamp constant), and the flat
            north-star check |got - ref| <= rtol * max|ref| against the reference's own fixture values;
  queries : |got-ref| <= rtol * max|ref|,  rtol = 1e-11 (fp64), 3e-5 (fp32)
  grads   : 2e-8 (fp64), 2e-2 (fp32), 0.2 (fp32 boundary regime: O(1/eps) conditioning)
  ranks   : identical in fp64; fp32: |rank - fp64 rank| <= #{entities whose fp64 score is within the two error bands of the
            target's} per query, on BOTH ranking tiers (exact FMA tier and tcgen05 tier).
The worst observed value of every check is recorded (conftest.record_parity -> profiles/parity_r2.json).
"""
from argparse import Namespace

import numpy as np
import pytest
import torch

from conftest import filters_from_arrays, golden_files, load_case, oracle_params, record_parity

pytestmark = pytest.mark.gpu

K_SCORE = 16           # multiples of eps_machine * unit allowed per score element; the reference's own fp32 / fp64 outputs sit at 0.1-1.1 (profiles/parity_r2.json)
K_BAND = 16            # half-width of the near-tie band of the fp32 rank bound, same units


def _observed(a, b):
    a = a.detach().cpu().double().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b.reshape(a.shape)).max() / max(np.abs(b).max(), 1e-300))


def _model_from_case(case, prefix="p_", device="cuda"):
    import complexhyperbolickge_b200 as chk
    args = Namespace(sizes=(case["n_ent"], case["n_rel2"], case["n_ent"]), rank=case["rank"], dropout=0, gamma=0,
                     dtype=case["dtype"], bias="learn", init_size=1e-3, multi_c=case["multi_c"])
    model = getattr(chk, case["name"])(args)
    sd = {k[len(prefix):] + ".weight": torch.from_numpy(v.copy()) for k, v in case.items()
          if k.startswith(prefix) and isinstance(v, np.ndarray)}
    model.load_state_dict(sd)
    return model.to(device)


def _close(a, b, rtol, what):
    a = a.detach().cpu().double().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = max(np.abs(b).max(), 1e-300)
    err = np.abs(a - b).max() / scale
    assert np.isfinite(a).all(), what + ": non-finite"
    assert err <= rtol, f"{what}: max err / max|ref| = {err:.3e} > {rtol}"


@pytest.mark.parametrize("path", golden_files("step_"), ids=lambda p: p.split("/")[-1][5:-4])
def test_step_vs_reference_golden(path):
    case = load_case(path)
    model = _model_from_case(case)
    dbl = case["dtype"] == "double"
    batch = torch.from_numpy(case["batch"]).cuda()
    neg = torch.from_numpy(case["neg"]).cuda()
    (q, c), bh = model.get_queries(batch[:, :2].unsqueeze(1))
    _close(q, case["q"], 1e-11 if dbl else 3e-5, "get_queries")
    assert q.shape == (batch.shape[0], 1, 2 * case["rank"]) and c.dim() == 3 and bh.shape == (batch.shape[0], 1, 1)
    st = 1e-9 if dbl else 2e-3
    with torch.no_grad():
        sa = model.score(model.get_queries(batch[:, :2]), model.get_rhs(None))
    _close(sa, case["score_all"], st, "score(q, candidates)")
    model.zero_grad()
    pos, _ = model(batch[:, :2].unsqueeze(1), batch[:, 2].unsqueeze(1))
    ngs, _ = model(batch[:, :2].unsqueeze(1), neg)
    _close(pos, case["score_pos"], st, "positive scores")
    _close(ngs, case["score_neg"], st, "negative scores")
    # conditioning-aware, per element, against the fp64 truth on the same weights
    import parity_units as PU
    p = oracle_params(case)
    mdt = torch.float64 if dbl else torch.float32
    bt_, nt_ = torch.from_numpy(case["batch"]), torch.from_numpy(case["neg"])
    truth, unit = PU.pair_truth(p, bt_[:, 0], bt_[:, 1], torch.cat((bt_[:, 2:3], nt_), 1))
    got_all = torch.cat((pos.detach().reshape(-1, 1), ngs.detach().reshape(nt_.shape)), 1)
    ru = PU.ratio(got_all, truth, unit, mdt)
    tag = f"{'fp64' if dbl else 'fp32'}/{case['regime']}"
    record_parity("step_scores_units/" + tag, max_err_in_eps_units=ru)
    record_parity("step_scores_rel/" + tag, max_rel_to_max=max(_observed(pos, case["score_pos"]), _observed(ngs, case["score_neg"])))
    record_parity("step_queries_rel/" + tag, max_rel_to_max=_observed(q, case["q"]))
    assert ru <= K_SCORE, f"score error {ru:.1f} eps-units > {K_SCORE}"
    lsig = torch.nn.functional.logsigmoid
    loss = -torch.cat([lsig(pos).view(-1), lsig(-ngs).view(-1)]).mean()
    loss.backward()
    assert abs(loss.item() - float(case["loss"])) <= (1e-11 if dbl else 2e-5) * max(1.0, abs(float(case["loss"])))
    gt = 2e-8 if dbl else 2e-2
    if case["regime"] == "boundary" and not dbl:
        gt = 0.2
    for k, p in model.named_parameters():
        ref = case["g_" + k.replace(".weight", "")]
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        if np.abs(ref).max() == 0:
            assert got.abs().max().item() <= 1e-30, k
        else:
            _close(got, ref, gt, "grad " + k)
            record_parity("step_grads_rel/" + tag, max_rel_to_max=_observed(got, ref))
    # the unfused API path (get_queries / get_rhs / score with autograd) must agree with the fused forward
    model.zero_grad()
    model.fused_forward = False
    pos2, _ = model(batch[:, :2].unsqueeze(1), batch[:, 2].unsqueeze(1))
    ngs2, _ = model(batch[:, :2].unsqueeze(1), neg)
    loss2 = -torch.cat([lsig(pos2).view(-1), lsig(-ngs2).view(-1)]).mean()
    loss2.backward()
    _close(ngs2, ngs.detach().cpu().numpy(), 1e-12 if dbl else 1e-5, "unfused vs fused scores")
    for k, p in model.named_parameters():
        ref = case["g_" + k.replace(".weight", "")]
        if np.abs(ref).max() > 0:
            _close(p.grad, ref, gt, "unfused grad " + k)


@pytest.mark.parametrize("algo", ["fma", "mma"])
@pytest.mark.parametrize("path", golden_files("rank_"), ids=lambda p: p.split("/")[-1][5:-4])
def test_ranking_vs_reference_golden(path, algo):
    """Filtered ranks of both tiers against the ranks the REFERENCE computed (fixtures).  fp64: identical.  fp32: the
    reference's fp32 ranks and ours may each differ from the fp64 truth by the near-tie count of the query, so they may differ
    from each other by at most twice that count."""
    import parity_units as PU
    case = load_case(path)
    model = _model_from_case(case)
    model.eval()
    model.rank_algo = algo
    filters = filters_from_arrays(case)
    ex = torch.from_numpy(case["test"])
    ranks = model.get_ranking(ex, filters["rhs"], batch_size=37)
    q = torch.stack([ex[:, 2], ex[:, 1] + case["n_rel2"] // 2, ex[:, 0]], -1)
    ranks_l = model.get_ranking(q, filters["lhs"], batch_size=500)
    assert ranks.dtype == torch.float32 and ranks.shape == (ex.shape[0],)
    p = oracle_params(case)
    for side, qs, got, ref in (("rhs", ex, ranks.numpy(), case["ranks_rhs"]), ("lhs", q, ranks_l.numpy(), case["ranks_lhs"])):
        if case["dtype"] == "double":
            assert np.array_equal(got, ref), np.abs(got - ref).max()
        else:
            truth, near = PU.rank_band_counts(p, qs, filters[side], K_BAND)
            d_truth = np.abs(got - truth.numpy())
            assert (d_truth <= near.numpy()).all(), (d_truth.max(), near.max().item())
            d = np.abs(got - ref)
            assert (d <= 2 * near.numpy()).all(), (d.max(), near.max().item())
            record_parity(f"rank_golden_fp32/{algo}/{case['regime']}", max_abs_rank_diff_vs_reference=d.max(), frac_queries_differing=(d > 0).mean(),
                          max_abs_rank_diff_vs_fp64_truth=d_truth.max())
    mr, mrr, hits = model.compute_metrics(ex, filters, batch_size=64)
    tol = 1e-6 if case["dtype"] == "double" else 2e-2
    assert abs(mr["rhs"] - case["mr"][0]) <= tol * case["mr"][0]
    assert abs(mrr["lhs"] - case["mrr"][1]) <= tol
    assert np.abs(hits["rhs"].numpy() - case["hits"][0]).max() <= (1e-6 if case["dtype"] == "double" else 2e-2)


@pytest.mark.parametrize("path", golden_files("curve_"), ids=lambda p: p.split("/")[-1][6:-4])
def test_loss_curve_vs_reference_golden(path):
    """N-step loss curve: our model + torch.optim + the KGOptimizer contract, replaying the reference's
    batches and negatives (CPU and CUDA generators differ, so they are injected)."""
    from complexhyperbolickge_b200.optim import KGOptimizer, N3
    case = load_case(path)
    model = _model_from_case(case, "p0_")
    optim_name, lr = case["regime"], float(case["lr"])
    opt = KGOptimizer(model, N3(0.0), getattr(torch.optim, optim_name)(model.parameters(), lr=lr),
                      batch_size=64, update_steps=1, neg_sample_size=case["neg_cat"].shape[1], double_neg=False,
                      verbose=False)
    negs = iter([])
    off = 0
    losses = []
    nsteps = len(case["batch_lens"])
    for t in range(nsteps):
        L = int(case["batch_lens"][t])
        b = torch.from_numpy(case["batch_cat"][off:off + L]).cuda()
        ng = torch.from_numpy(case["neg_cat"][off:off + L]).cuda()
        off += L
        opt.get_neg_samples = lambda _b, ng=ng: ng
        l = opt.calculate_loss(b)
        l.backward()
        opt.optimizer.step()
        opt.optimizer.zero_grad()
        losses.append(l.item())
    err = np.abs(np.array(losses) - case["step_losses"]).max()
    assert err <= 1e-9, err
    for k, p in model.named_parameters():
        _close(p, case["pT_" + k.replace(".weight", "")], 1e-8, "final " + k)


# ------------------------------------------------------------------------------- oracle on seeded inputs
def _random_params(kind_name, rank, n_ent, n_rel2, dtype, multi_c, seed, regime="trained"):
    from oracle import chk_oracle as O
    g = torch.Generator().manual_seed(seed)
    n = 2 * (rank - 1)
    std = float(np.sqrt(0.4 / (2 * rank)))
    rn = lambda *s, sd=1.0: (torch.randn(*s, generator=g, dtype=torch.float64) * sd).to(dtype)
    ru = lambda *s, lo=-1.0, hi=1.0: (torch.rand(*s, generator=g, dtype=torch.float64) * (hi - lo) + lo).to(dtype)
    att = kind_name == "FFTAttH"
    return O.Params(O.KIND[kind_name], rank, multi_c, rn(n_ent, 2 * rank, sd=std), rn(n_rel2, 2 * n, sd=0.05),
                    ru(n_rel2, 2 * n if att else n), ru(n_rel2 if multi_c else 1, 1, lo=0.5, hi=2.0),
                    rn(n_ent, 1, sd=0.1), rn(n_ent, 1, sd=0.1), rn(n_rel2, n) if att else None)


def _model_from_params(p, name):
    import complexhyperbolickge_b200 as chk
    dt = "double" if p.dtype == torch.float64 else "float"
    args = Namespace(sizes=(p.entity.shape[0], p.rel.shape[0], p.entity.shape[0]), rank=p.rank, dropout=0, gamma=0,
                     dtype=dt, bias="learn", init_size=1e-3, multi_c=p.multi_c)
    model = getattr(chk, name)(args)
    sd = {"entity.weight": p.entity, "rel.weight": p.rel, "rel_diag.weight": p.rel_diag, "c.weight": p.c,
          "bh.weight": p.bh, "bt.weight": p.bt}
    if p.context_vec is not None:
        sd["context_vec.weight"] = p.context_vec
    model.load_state_dict(sd)
    return model.cuda()


@pytest.mark.parametrize("name", ["FFTRotH", "FFTRefH", "FFTAttH"])
@pytest.mark.parametrize("rank", [9, 17, 33, 65, 129, 257])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_step_vs_oracle_all_ranks(name, rank, dtype):
    from oracle import chk_oracle as O
    n_ent, n_rel2, B, neg = 301, 10, 37, 19
    p = _random_params(name, rank, n_ent, n_rel2, dtype, True, seed=rank)
    model = _model_from_params(p, name)
    g = torch.Generator().manual_seed(rank + 1)
    batch = torch.stack([torch.randint(0, n_ent, (B,), generator=g), torch.randint(0, n_rel2, (B,), generator=g),
                         torch.randint(0, n_ent, (B,), generator=g)], 1)
    negs = torch.randint(0, n_ent, (B, neg), generator=g)
    loss_ref, grads_ref = O.neg_sampling_loss(p, batch, negs)
    bc, nc = batch.cuda(), negs.cuda()
    tails = torch.cat((bc[:, 2:3], nc), 1)
    s, _ = model(bc[:, :2].unsqueeze(1), tails)
    sign = torch.ones_like(s)
    sign[:, 1:] = -1
    loss = -torch.nn.functional.logsigmoid(sign * s).mean()
    loss.backward()
    dbl = dtype == torch.float64
    assert abs(loss.item() - loss_ref.item()) <= (1e-12 if dbl else 2e-5)
    import parity_units as PU
    truth, unit = PU.pair_truth(p, batch[:, 0], batch[:, 1], torch.cat((batch[:, 2:3], negs), 1))
    ru = PU.ratio(s.detach().reshape(truth.shape), truth, unit, dtype)
    record_parity(f"oracle_scores_units/{'fp64' if dbl else 'fp32'}/r{rank}", max_err_in_eps_units=ru)
    assert ru <= K_SCORE, f"score error {ru:.1f} eps-units > {K_SCORE}"
    for k, gr in grads_ref.items():
        got = getattr(model, k).weight.grad
        _close(got, gr.numpy(), 2e-9 if dbl else 2e-2, f"grad {k}")
        record_parity(f"oracle_grads_rel/{'fp64' if dbl else 'fp32'}/r{rank}", max_rel_to_max=_observed(got, gr.numpy()))
    q_ref, _ = O.query_fwd(p, batch[:, 0], batch[:, 1])
    (q, _c), _ = model.get_queries(bc[:, :2])
    _close(q.squeeze(1), q_ref.numpy(), 1e-12 if dbl else 3e-5, "get_queries")


@pytest.mark.parametrize("algo", ["fma", "mma"])
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("name,rank", [("FFTRotH", 33), ("FFTRefH", 65), ("FFTAttH", 33), ("FFTRotH", 257)])
def test_ranking_vs_oracle(name, rank, dtype, algo):
    """Filtered ranks of both tiers on a seeded graph with heavy (Zipf) filters vs the oracle.  fp64: identical.  fp32: within
    the per-query near-tie count of the fp64 truth, and within twice that of the oracle's own fp32 ranks."""
    import parity_units as PU
    from oracle import chk_oracle as O
    n_ent, n_rel2, nq = 1500, 12, 210
    p = _random_params(name, rank, n_ent, n_rel2, dtype, True, seed=7 * rank)
    model = _model_from_params(p, name)
    model.rank_algo = algo
    rng = np.random.default_rng(rank)
    pop = 1.0 / np.arange(1, n_ent + 1)
    pop /= pop.sum()
    tri = np.unique(np.stack([rng.choice(n_ent, 6000, p=pop), rng.integers(0, n_rel2, 6000),
                              rng.choice(n_ent, 6000, p=pop)], 1), axis=0)
    filters = {}
    for h, r, t in tri:
        filters.setdefault((int(h), int(r)), []).append(int(t))
    ex = torch.from_numpy(tri[rng.permutation(len(tri))[:nq]].astype(np.int64))
    ref = O.get_ranking(p, ex, filters, batch_size=64).numpy()
    got = model.get_ranking(ex, filters, batch_size=100).numpy()
    d = np.abs(got - ref)
    if dtype == torch.float64:
        assert d.max() == 0, d.max()
    else:
        truth, near = PU.rank_band_counts(p, ex, filters, K_BAND)
        d_truth = np.abs(got - truth.numpy())
        assert (d_truth <= near.numpy()).all(), (d_truth.max(), near.max().item())
        assert (d <= 2 * near.numpy()).all(), (d.max(), near.max().item())
        record_parity(f"rank_oracle_fp32/{algo}/{name}-r{rank}", max_abs_rank_diff_vs_fp32_oracle=d.max(), frac_queries_differing=(d > 0).mean(),
                      max_abs_rank_diff_vs_fp64_truth=d_truth.max(), frac_differing_vs_fp64_truth=(d_truth > 0).mean(),
                      mean_near_tie_count=near.float().mean().item())


def test_ranking_shard_sum_equals_single(monkeypatch):
    """integer counts summed over shards are independent of the shard count — emulated on one
    GPU by slicing the table (no collective needed to test the arithmetic)."""
    from complexhyperbolickge_b200 import ops, ranking
    p = _random_params("FFTRotH", 33, 5000, 8, torch.float32, True, seed=3)
    model = _model_from_params(p, "FFTRotH")
    g = torch.Generator().manual_seed(5)
    nq = 300
    qs = torch.stack([torch.randint(0, 5000, (nq,), generator=g), torch.randint(0, 8, (nq,), generator=g),
                      torch.randint(0, 5000, (nq,), generator=g)], 1)
    filters = {(int(h), int(r)): [int(t), int((t * 7) % 5000), int((t + 13) % 5000)] for h, r, t in qs.numpy()}
    single = model.get_ranking(qs, filters, batch_size=128)

    class FakeState(ranking.EvalState):
        def __init__(self, model, world, rid):
            self.world, self.rank_id = 1, 0
            self.lo, self.hi = ranking.shard_bounds(model.sizes[0], world, rid)
            ent = model.entity.weight.detach()
            self.entity = ent[self.lo:self.hi].contiguous()
            self.bt = model.bt.weight.detach().view(-1)[self.lo:self.hi].contiguous()
            self.hn = ops.row_hnorm(model.rank, self.entity)
            self.algo, self.shadow = ops.CHK_RANK_FMA, None

    findex = model._filter_index(filters)
    for world in (2, 3, 8):
        counts = torch.zeros(nq, dtype=torch.int64, device="cuda")
        indptr, idx = findex.batch_csr(qs.numpy())
        for rid in range(world):
            st = FakeState(model, world, rid)
            ranking.rank_batch(model, st, qs.cuda(), torch.from_numpy(indptr).cuda(), torch.from_numpy(idx).cuda(),
                               int(idx.size), counts)
        assert torch.equal((counts + 1).float().cpu(), single), world


def test_rank_counts_match_score_matrix():
    """The fused count kernel vs counting on the materialised chk_score_all matrix (same canonical
    arithmetic => exact equality), at an L2-exceeding table size, fp32 and fp64."""
    from complexhyperbolickge_b200 import ops
    for dtype, n_ent, rank in ((torch.float32, 200_003, 33), (torch.float64, 50_001, 65)):
        p = _random_params("FFTRotH", rank, n_ent, 4, dtype, True, seed=11)
        model = _model_from_params(p, "FFTRotH")
        g = torch.Generator().manual_seed(1)
        b = 130
        qs = torch.stack([torch.randint(0, n_ent, (b,), generator=g), torch.randint(0, 4, (b,), generator=g),
                          torch.randint(0, n_ent, (b,), generator=g)], 1).cuda()
        with torch.no_grad():
            q, _ = ops.query_fwd(model.KIND, rank, True, model.entity.weight, model.rel.weight, model.rel_diag.weight,
                                 None, model.c.weight, qs[:, 0].contiguous(), qs[:, 1].contiguous())
            qn = ops.row_hnorm(rank, q)
            ent = model.entity.weight.detach()
            hn = ops.row_hnorm(rank, ent)
            bh = model.bh.weight.detach().view(-1)[qs[:, 0]].contiguous()
            bt = model.bt.weight.detach().view(-1).contiguous()
            rows = ent[qs[:, 2]].contiguous()
            tgt = ops.target_scores(rank, q, qn, bh, rows, ops.row_hnorm(rank, rows), bt[qs[:, 2]].contiguous())
            S = ops.score_all(rank, q, qn, bh, ent, hn, bt)
            assert torch.equal(S[torch.arange(b), qs[:, 2]], tgt)       # target == its own column, bit for bit
            counts = torch.zeros(b, dtype=torch.int64, device="cuda")
            indptr = torch.arange(b + 1, dtype=torch.int64, device="cuda")
            ops.rank_counts(ops.CHK_RANK_FMA, rank, q, qn, bh, tgt, ent, hn, bt, 0, indptr, qs[:, 2].contiguous(), b,
                            counts)
            ref = (S >= tgt[:, None]).sum(1) - 1
            assert torch.equal(counts, ref)


def test_edge_cases():
    import complexhyperbolickge_b200 as chk
    from complexhyperbolickge_b200 import ops
    p = _random_params("FFTAttH", 9, 40, 6, torch.float64, False, seed=2)
    model = _model_from_params(p, "FFTAttH")
    # empty and single-query batches
    empty = torch.zeros((0, 2), dtype=torch.int64, device="cuda")
    (q, c), bh = model.get_queries(empty)
    assert q.shape == (0, 1, 18)
    one = torch.tensor([[3, 2, 5]], device="cuda")
    s, _ = model(one[:, :2].unsqueeze(1), one[:, 2:3])
    assert s.shape == (1, 1, 1) and torch.isfinite(s).all()
    # double_neg-shaped per-negative queries (B, n, 2) — SURVEY §0.4
    qq = torch.randint(0, 6, (4, 5, 2), device="cuda")
    tails = torch.randint(0, 40, (4, 5), device="cuda")
    s1, _ = model(qq, tails)
    model.fused_forward = False
    s2, _ = model(qq, tails)
    assert s1.shape == (4, 5, 1) and torch.allclose(s1, s2, rtol=1e-12, atol=1e-14)
    # missing filter key -> KeyError like the reference (models/base.py:266)
    with pytest.raises(KeyError):
        model.get_ranking(torch.tensor([[1, 1, 1]]), {(0, 0): [1]}, batch_size=4)
    # unsupported rank / CPU tensors fail loudly
    with pytest.raises(ValueError):
        chk.FFTRotH(Namespace(sizes=(10, 2, 10), rank=10, dropout=0, gamma=0, dtype="float", bias="learn",
                              init_size=1e-3, multi_c=True))
    with pytest.raises(RuntimeError):
        ops.row_hnorm(9, torch.zeros(4, 18))
    # clamp regime ties: init weights in fp32 -> every score equals -acosh(1+4e-3)^2, rank = N - |filter|
    args = Namespace(sizes=(500, 4, 500), rank=33, dropout=0, gamma=0, dtype="float", bias="learn", init_size=1e-3,
                     multi_c=True)
    m2 = chk.FFTRotH(args).cuda()
    ex = torch.tensor([[1, 0, 2], [3, 1, 4]])
    ranks = m2.get_ranking(ex, {(1, 0): [2, 7], (3, 1): [4]}, batch_size=2)
    assert ranks.tolist() == [500 - 2 + 1, 500 - 1 + 1]


@pytest.mark.parametrize("name", ["FFTRotH", "FFTRefH", "FFTAttH"])
@pytest.mark.parametrize("rank", [9, 17, 33])
def test_grouped_query_transform_matches_lane_group_kernel(name, rank):
    """chk_query_fwd_grouped (thread-per-query K1, queries processed in relation order) vs chk_query_fwd on the same
    inputs, ragged batch, multi_c on/off, and vs the oracle."""
    from complexhyperbolickge_b200 import ops
    from oracle import chk_oracle as O
    for multi_c in (True, False):
        p = _random_params(name, rank, 900, 14, torch.float32, multi_c, seed=rank + 3)
        model = _model_from_params(p, name)
        g = torch.Generator().manual_seed(rank)
        nq = 1000 + 7
        h, r = torch.randint(0, 900, (nq,), generator=g), torch.randint(0, 14, (nq,), generator=g)
        args = (model.KIND, rank, multi_c, model.entity.weight.detach(), model.rel.weight.detach(),
                model.rel_diag.weight.detach(), None if model._ctx_weight() is None else model._ctx_weight().detach(),
                model.c.weight.detach(), h.cuda(), r.cuda())
        q0, c0 = ops.query_fwd(*args, grouped=False)
        q1, c1 = ops.query_fwd(*args, grouped=True)
        assert torch.equal(c0, c1)
        _close(q1, q0.cpu().numpy(), 2e-6, "grouped vs lane-group get_queries")
        q_ref, _ = O.query_fwd(p, h, r)
        _close(q1, q_ref.numpy(), 3e-5, "grouped get_queries vs oracle")
