"""GPU parity of the tensor-core rank tier (CHK_RANK_MMA, csrc/chk_rank_mma.cu), through the C ABI.

The tier's contract is EXACT equality of the integer counts with the exact tier (CHK_RANK_FMA), which is
itself pinned to the reference by tests/test_gpu_parity.py.  Checked here:
  * approximate scores stay inside the proven band around the exact fp32 scores (|s~ - s| <= band), with
    margin (max ratio < 0.5), and pairs decided in the clamp regime (band == 0) are bit-identical;
  * unfiltered counts == counting on the materialised exact score matrix;
  * model.get_ranking with rank_algo="mma" == rank_algo="fma" (filters, ragged sizes, b > 1024, init regime);
  * a too-small re-check list raises the sticky overflow flag and get_ranking falls back to the exact tier.
"""
from argparse import Namespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(rank, n_ent, b, seed, regime="trained", bias=True, dtype=torch.float32):
    from complexhyperbolickge_b200 import ops
    g = torch.Generator().manual_seed(seed)
    std = float(np.sqrt(0.4 / (2 * rank))) if regime == "trained" else 1e-3
    ent = (torch.randn(n_ent, 2 * rank, generator=g, dtype=torch.float64) * std).to(dtype).cuda()
    q = (torch.randn(b, 2 * rank, generator=g, dtype=torch.float64) * std).to(dtype).cuda()
    bh = (torch.randn(b, generator=g, dtype=torch.float64) * 0.1).to(dtype).cuda() if bias else None
    bt = (torch.randn(n_ent, generator=g, dtype=torch.float64) * 0.1).to(dtype).cuda() if bias else None
    tails = torch.randint(0, n_ent, (b,), generator=g).cuda()
    qn, hn = ops.row_hnorm(rank, q), ops.row_hnorm(rank, ent)
    rows = ent[tails].contiguous()
    tgt = ops.target_scores(rank, q, qn, bh, rows, ops.row_hnorm(rank, rows), bt[tails].contiguous() if bias else None)
    return ent, q, bh, bt, qn, hn, tgt


@pytest.mark.parametrize("rank,n_ent,b,regime,bias,dtype", [
    (33, 1000, 150, "trained", True, torch.float32), (33, 1000, 150, "init", True, torch.float32),
    (9, 130, 7, "trained", False, torch.float32), (65, 5000, 300, "trained", True, torch.float32),
    (257, 20000, 500, "trained", True, torch.float32), (257, 300, 1100, "trained", True, torch.float32),
    (129, 4097, 129, "trained", False, torch.float32),
    (65, 5000, 300, "trained", True, torch.float64), (33, 1000, 150, "trained", False, torch.float64),
    (257, 3000, 200, "trained", True, torch.float64), (33, 700, 40, "init", True, torch.float64)])
def test_mma_scores_within_band_and_counts_exact(rank, n_ent, b, regime, bias, dtype):
    """fp64 rows: the model / exact tier are fp64, the tensor-core prefilter and its band stay fp32."""
    from complexhyperbolickge_b200 import ops
    ent, q, bh, bt, qn, hn, tgt = _setup(rank, n_ent, b, seed=rank + n_ent, regime=regime, bias=bias, dtype=dtype)
    S = ops.score_all(rank, q, qn, bh, ent, hn, bt)
    shadow = ops.entity_shadow(rank, ent, hn, bt)
    ws = ops.rank_mma_workspace(rank, b, ent.device)
    St, band, counts = ops.score_all_mma(rank, q, qn, bh, tgt, ent, hn, bt, shadow, ws)
    n_list, overflow = ops.rank_mma_status(ws)
    assert not overflow
    assert torch.isfinite(St).all() and torch.isfinite(band).all()
    diff = (St.double() - S.double()).abs()
    pos = band > 0
    if pos.any():
        ratio = (diff[pos] / band[pos].double()).max().item()
        assert ratio < 0.5, f"approximate score leaves half of its proven band: {ratio}"
    if (~pos).any():
        assert torch.equal(St[~pos], S[~pos]), "clamp-regime pairs must be bit-identical to the exact tier"
    if regime == "init" and dtype == torch.float32:
        assert (~pos).float().mean().item() > 0.99          # init_size=1e-3 in fp32: everything is clamped
    assert torch.equal(counts, (S >= tgt[:, None]).sum(1))
    assert n_list <= (0.2 if regime == "trained" else 1.0) * b * n_ent


def _model(name, rank, n_ent, n_rel2, seed, regime="trained", dtype="float"):
    import complexhyperbolickge_b200 as chk
    from complexhyperbolickge_b200 import synthetic
    args = Namespace(sizes=(n_ent, n_rel2, n_ent), rank=rank, dropout=0, gamma=0, dtype=dtype, bias="learn",
                     init_size=1e-3, multi_c=True)
    torch.manual_seed(seed)
    m = getattr(chk, name)(args).cuda()
    if regime == "trained":
        synthetic.trained_like_(m, seed)
    return m


@pytest.mark.parametrize("name,rank,n_ent,nq,batch,regime,dtype", [
    ("FFTRotH", 33, 40943, 700, 500, "trained", "float"), ("FFTRefH", 33, 14541, 300, 128, "trained", "float"),
    ("FFTAttH", 33, 5000, 260, 1300, "trained", "float"), ("FFTRotH", 257, 30001, 500, 500, "trained", "float"),
    ("FFTRotH", 33, 3000, 64, 64, "init", "float"), ("FFTRotH", 65, 20000, 333, 200, "trained", "float"),
    ("FFTRotH", 65, 40943, 600, 500, "trained", "double"), ("FFTAttH", 33, 9000, 150, 64, "trained", "double"),
    ("FFTRefH", 17, 2000, 50, 50, "init", "double")])
def test_get_ranking_mma_equals_fma(name, rank, n_ent, nq, batch, regime, dtype):
    n_rel2 = 8
    m = _model(name, rank, n_ent, n_rel2, seed=rank, regime=regime, dtype=dtype)
    rng = np.random.default_rng(n_ent)
    pop = 1.0 / np.arange(1, n_ent + 1)
    pop /= pop.sum()
    qs = np.stack([rng.choice(n_ent, nq, p=pop), rng.integers(0, n_rel2, nq), rng.choice(n_ent, nq, p=pop)], 1).astype(np.int64)
    filters = {}
    for h, r, t in qs:
        filters.setdefault((int(h), int(r)), []).extend([int(t), int((t * 7 + 1) % n_ent), int(rng.integers(0, n_ent))])
    m.rank_algo = "fma"
    want = m.get_ranking(torch.from_numpy(qs), filters, batch_size=batch)
    m.rank_algo = "mma"
    got = m.get_ranking(torch.from_numpy(qs), filters, batch_size=batch)
    assert torch.equal(got, want), (got - want).abs().max()


def test_mma_overflow_flag_and_fallback(monkeypatch):
    from complexhyperbolickge_b200 import ops
    m = _model("FFTRotH", 33, 6000, 4, seed=1)
    with torch.no_grad():                       # duplicate one entity row many times: exact ties with the target -> band
        m.entity.weight[100:3100] = m.entity.weight[7]
        m.bt.weight[100:3100] = m.bt.weight[7]
    qs = torch.tensor([[5, 1, 7], [9, 2, 7], [11, 0, 7]])
    filters = {(5, 1): [7], (9, 2): [7], (11, 0): [7]}
    m.rank_algo = "fma"
    want = m.get_ranking(qs, filters, batch_size=8)
    real_ws = ops.rank_mma_workspace

    def tiny_ws(rank, b, device):               # fixed part + 1024 list entries only
        full = real_ws(rank, b, device)
        import complexhyperbolickge_b200._lib as L
        cap = L.lib().chk_rank_mma_workspace_bytes(rank, b) - max(b * 8192, 1 << 20) * 8
        ws = full[: cap + 1024 * 8].clone()
        ops.rank_mma_reset(ws)
        return ws

    monkeypatch.setattr(ops, "rank_mma_workspace", tiny_ws)
    m.rank_algo = "mma"
    got = m.get_ranking(qs, filters, batch_size=8)
    assert torch.equal(got, want)
    # and the flag really was raised on that path
    ent = m.entity.weight.detach()
    q, _ = ops.query_fwd(m.KIND, 33, True, ent, m.rel.weight.detach(), m.rel_diag.weight.detach(), None,
                         m.c.weight.detach(), qs[:, 0].cuda().contiguous(), qs[:, 1].cuda().contiguous())
    qn, hn = ops.row_hnorm(33, q), ops.row_hnorm(33, ent)
    bh = m.bh.weight.detach().view(-1)[qs[:, 0].cuda()].contiguous()
    bt = m.bt.weight.detach().view(-1).contiguous()
    rows = ent[qs[:, 2].cuda()].contiguous()
    tgt = ops.target_scores(33, q, qn, bh, rows, ops.row_hnorm(33, rows), bt[qs[:, 2].cuda()].contiguous())
    ws = tiny_ws(33, 3, ent.device)
    counts = torch.zeros(3, dtype=torch.int64, device="cuda")
    ops.rank_counts(ops.CHK_RANK_MMA, 33, q, qn, bh, tgt, ent, hn, bt, 0, torch.zeros(4, dtype=torch.int64, device="cuda"),
                    torch.zeros(1, dtype=torch.int64, device="cuda"), 0, counts, ops.entity_shadow(33, ent, hn, bt), ws)
    n_list, overflow = ops.rank_mma_status(ws)
    assert overflow and n_list > 1024


def test_full_size_big4m_properties():
    """BASELINE.json configs[4] at its FULL size (4,000,000 entities, rank 257): size-independent properties —
    tensor-core counts == exact-tier counts, shard sums == single pass, counts bounded by the table size and
    >= 1 without a filter (the target outranks itself), filtered count = unfiltered - #filtered hits."""
    from complexhyperbolickge_b200 import ops
    rank, n_ent, b = 257, 4_000_000, 40
    g = torch.Generator(device="cuda").manual_seed(3)
    std = float(np.sqrt(0.4 / (2 * rank)))
    ent = torch.randn(n_ent, 2 * rank, generator=g, device="cuda") * std
    q = torch.randn(b, 2 * rank, generator=g, device="cuda") * std
    bh = torch.randn(b, generator=g, device="cuda") * 0.1
    bt = torch.randn(n_ent, generator=g, device="cuda") * 0.1
    tails = torch.randint(0, n_ent, (b,), generator=g, device="cuda")
    qn, hn = ops.row_hnorm(rank, q), ops.row_hnorm(rank, ent)
    rows = ent[tails].contiguous()
    tgt = ops.target_scores(rank, q, qn, bh, rows, ops.row_hnorm(rank, rows), bt[tails].contiguous())
    empty_ip = torch.zeros(b + 1, dtype=torch.int64, device="cuda")
    dummy = torch.zeros(1, dtype=torch.int64, device="cuda")
    c_fma = torch.zeros(b, dtype=torch.int64, device="cuda")
    ops.rank_counts(ops.CHK_RANK_FMA, rank, q, qn, bh, tgt, ent, hn, bt, 0, empty_ip, dummy, 0, c_fma)
    shadow = ops.entity_shadow(rank, ent, hn, bt)
    ws = ops.rank_mma_workspace(rank, b, ent.device)
    c_mma = torch.zeros(b, dtype=torch.int64, device="cuda")
    ops.rank_counts(ops.CHK_RANK_MMA, rank, q, qn, bh, tgt, ent, hn, bt, 0, empty_ip, dummy, 0, c_mma, shadow, ws)
    assert not ops.rank_mma_status(ws)[1]
    assert torch.equal(c_mma, c_fma)
    assert (c_fma >= 1).all() and (c_fma <= n_ent).all()
    del shadow
    # three contiguous shards (tile aligned) sum to the single pass
    c_sh = torch.zeros(b, dtype=torch.int64, device="cuda")
    bounds = [0, 1_333_376, 2_666_752, n_ent]
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        e_s, h_s, b_s = ent[lo:hi], hn[lo:hi], bt[lo:hi]
        sh = ops.entity_shadow(rank, e_s.contiguous(), h_s.contiguous(), b_s.contiguous())
        ops.rank_counts(ops.CHK_RANK_MMA, rank, q, qn, bh, tgt, e_s, h_s, b_s, lo, empty_ip, dummy, 0, c_sh, sh, ws)
        del sh
    assert torch.equal(c_sh, c_fma)
    # filter = {target} U 5 random ids per query: filtered count = unfiltered - #{listed ids scoring >= target}
    extra = torch.randint(0, n_ent, (b, 5), generator=g, device="cuda")
    lists = [torch.unique(torch.cat([tails[i:i + 1], extra[i]])) for i in range(b)]
    ip = torch.tensor([0] + list(np.cumsum([len(x) for x in lists])), dtype=torch.int64, device="cuda")
    ix = torch.cat(lists).contiguous()
    c_f = torch.zeros(b, dtype=torch.int64, device="cuda")
    ops.rank_counts(ops.CHK_RANK_FMA, rank, q, qn, bh, tgt, ent, hn, bt, 0, ip, ix, ix.numel(), c_f)
    want = c_fma.clone()
    for i in range(b):
        s_i = ops.score_all(rank, q[i:i + 1].contiguous(), qn[i:i + 1].contiguous(), bh[i:i + 1].contiguous(),
                            ent[lists[i]].contiguous(), hn[lists[i]].contiguous(), bt[lists[i]].contiguous())
        want[i] -= int((s_i[0] >= tgt[i]).sum())
    assert torch.equal(c_f, want)
