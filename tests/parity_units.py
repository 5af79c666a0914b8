"""Conditioning-aware error units for the parity tests.

The score of a pair is s = (bh + bt) - acosh(x)^2 with x + 1 = 2 |<z,w> - 1|^2 / (zn wn), zn = sum|z|^2 - 1 (clamped), and
|ds/dx| = 2 acosh(x) / sqrt(x^2 - 1) <= 2.  A relative perturbation e of the inputs of x (dot products, the two norms with their
cancellation against 1) moves s by at most about

    unit * e,   unit = 2 (x + 1) (1 + 1/|zn| + 1/|wn|) + |s| + |bh| + |bt|

so errors are reported and bounded in multiples of  eps_machine * unit  PER ELEMENT instead of a flat fraction of max|s|.
The units are evaluated by the oracle in fp64 on the model's parameters with the MODEL's clamp constant (4e-3 for fp32
models), which is also the fp64 "truth" the fp32 ranks are compared with: a fp32 rank may differ from the fp64 rank by at most
the number of entities whose fp64 score lies within the two error bands of the target's."""
import contextlib

import numpy as np
import torch

from oracle import chk_oracle as O

EPS = {torch.float32: 2.0 ** -24, torch.float64: 2.0 ** -53}


@contextlib.contextmanager
def model_clamp(dtype):
    old = O.BALL_EPS[torch.float64]
    O.BALL_EPS[torch.float64] = O.BALL_EPS[dtype]
    try:
        yield
    finally:
        O.BALL_EPS[torch.float64] = old


def params64(p):
    d = lambda t: None if t is None else t.double()
    return O.Params(p.kind, p.rank, p.multi_c, d(p.entity), d(p.rel), d(p.rel_diag), d(p.c), d(p.bh), d(p.bt), d(p.context_vec), p.bias)


def _unit(st, s, bh, bt):
    x, zn, wn = st["x"], st["zn"], st["wn"]
    return 2 * (x + 1) * (1 + 1 / zn.abs() + 1 / wn.abs()) + s.abs() + bh.abs() + bt.abs()


def pair_truth(p, heads, rels, tails):
    """fp64 scores [B, nt] of the gathered tails and their error units, for a model whose dtype is p.dtype."""
    p64 = params64(p)
    with model_clamp(p.dtype):
        q, _ = O.query_fwd(p64, heads, rels)
        st = {}
        bh = p64.bh[heads].unsqueeze(1)
        s = O.score_pairs(p64, q.unsqueeze(1), bh, tails, st)
    return s.squeeze(-1), _unit(st, s, bh, p64.bt[tails]).squeeze(-1)


def table_truth(p, heads, rels):
    """fp64 scores [b, N] against the whole table and their error units."""
    p64 = params64(p)
    with model_clamp(p.dtype):
        q, _ = O.query_fwd(p64, heads, rels)
        st = {}
        d = O.distance_fwd(q.unsqueeze(1), p64.entity.unsqueeze(0), st)
        bh = p64.bh[heads].unsqueeze(1)
        s = (bh + p64.bt.unsqueeze(0)) + (-(d * d))
    return s.squeeze(-1), _unit(st, s, bh, p64.bt.unsqueeze(0)).squeeze(-1)


def ratio(got, truth, unit, dtype):
    """max over elements of |got - truth| / (eps(dtype) * unit)."""
    got = torch.as_tensor(np.asarray(got.detach().cpu() if isinstance(got, torch.Tensor) else got), dtype=torch.float64).reshape(truth.shape)
    return ((got - truth).abs() / (EPS[dtype] * unit)).max().item()


def rank_band_counts(p, queries, filters, k_units):
    """Per query: fp64 filtered rank and the number of unfiltered entities whose fp64 score lies within
    k_units * eps * (unit_e + unit_t) of the target's — the largest rank difference a correct fp-dtype implementation may show."""
    heads, rels, tails = queries[:, 0], queries[:, 1], queries[:, 2]
    s, u = table_truth(p, heads, rels)
    nq = queries.shape[0]
    idx = torch.arange(nq)
    st, ut = s[idx, tails].unsqueeze(1), u[idx, tails].unsqueeze(1)
    mask = torch.ones_like(s, dtype=torch.bool)
    for i, (h, r, t) in enumerate(queries.tolist()):
        mask[i, list(filters[(h, r)]) + [t]] = False
    ranks = 1 + ((s >= st) & mask).sum(1)
    near = ((s - st).abs() <= k_units * EPS[p.dtype] * (u + ut)) & mask
    return ranks.to(torch.float32), near.sum(1)
