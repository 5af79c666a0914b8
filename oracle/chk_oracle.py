"""CPU oracle for the FFTRotH / FFTRefH / FFTAttH hot path.  TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement (torch-CPU used purely as a multi-threaded array
library: no autograd, no torch.fft) of the algorithm that
htmai-880/ComplexHyperbolicKGE runs for the scoring hot path.  Nothing under
``complexhyperbolickge_b200/`` may import it: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs do, and there only as the checker or the CPU arm.

Parity pinning: the reference ships no tests and no golden vectors,
so this oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF run in the
build container (``oracle/make_golden.py`` imports /root/reference through a
shim, ``model.lift = True``) — forward values, autograd gradients, ranks and a
loss curve — committed under ``tests/golden/`` and checked by
``tests/test_oracle_golden.py``.

Every function cites the reference file:line it restates (paths relative to the
reference root).  All backward formulas are hand-derived; they are checked
against the reference's autograd through the golden fixtures.

Conventions:  r = rank, n = dim = 2(r-1), entity row = [Re X_0..X_{r-1} | Im X_0..X_{r-1}].
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

MIN_NORM = 1e-15                                   # utils/complexhyperbolic.py:12
BALL_EPS = {torch.float32: 4e-3, torch.float64: 1e-5}  # utils/complexhyperbolic.py:13
PROJ_EPS = 1e-5                                    # utils/complexhyperbolic.py:83 (dtype independent)
ROT, REF, ATT = 0, 1, 2
KIND = {"FFTRotH": ROT, "FFTRefH": REF, "FFTAttH": ATT}


# --------------------------------------------------------------------------- DFT (O1, O2)
_DFT_CACHE: Dict[Tuple[int, torch.dtype], Tuple[torch.Tensor, torch.Tensor]] = {}


def _dft_mats(n: int, dtype) -> Tuple[torch.Tensor, torch.Tensor]:
    """C[j,k] = cos(2*pi*j*k/n)/sqrt(n), S[j,k] = sin(2*pi*j*k/n)/sqrt(n), j<n, k<=n/2.

    Built in float64 with exact integer phase reduction, then cast."""
    key = (n, dtype)
    if key not in _DFT_CACHE:
        j = np.arange(n)[:, None]
        k = np.arange(n // 2 + 1)[None, :]
        ph = (j * k) % n
        ang = 2.0 * np.pi * ph / n
        C = np.cos(ang) / math.sqrt(n)
        S = np.sin(ang) / math.sqrt(n)
        _DFT_CACHE[key] = (torch.from_numpy(C).to(dtype), torch.from_numpy(S).to(dtype))
    return _DFT_CACHE[key]


def _wgt(r: int, dtype) -> torch.Tensor:
    w = torch.full((r,), 2.0, dtype=dtype)
    w[0] = 1.0
    w[-1] = 1.0
    return w


def irfft_ortho(x2r: torch.Tensor) -> torch.Tensor:
    """torch.fft.irfft(complex(x), norm='ortho')  (models/complexhyperbolic.py:83-84,113-114,148-149).

    u_j = n^-1/2 [a_0 + (-1)^j a_m + 2 sum_{0<k<m} (a_k cos(2 pi jk/n) - b_k sin(2 pi jk/n))];
    Im X_0 and Im X_m are ignored (C2R semantics)."""
    r = x2r.shape[-1] // 2
    n = 2 * (r - 1)
    C, S = _dft_mats(n, x2r.dtype)
    w = _wgt(r, x2r.dtype)
    a, b = x2r[..., :r], x2r[..., r:]
    return (a * w) @ C.T - (b * w) @ S.T          # sin column 0 and m are exactly 0


def irfft_ortho_bwd(g: torch.Tensor) -> torch.Tensor:
    """Adjoint of irfft_ortho: (dRe, dIm)_k = wgt_k * rfft_ortho(g)_k."""
    n = g.shape[-1]
    r = n // 2 + 1
    C, S = _dft_mats(n, g.dtype)
    w = _wgt(r, g.dtype)
    return torch.cat(((g @ C) * w, -(g @ S) * w), -1)


def rfft_ortho(v: torch.Tensor) -> torch.Tensor:
    """cat(rfft(v, norm='ortho').real, .imag)  (models/complexhyperbolic.py:92-93,118-119,162-163)."""
    n = v.shape[-1]
    C, S = _dft_mats(n, v.dtype)
    return torch.cat((v @ C, -(v @ S)), -1)


def rfft_ortho_bwd(g2r: torch.Tensor) -> torch.Tensor:
    r = g2r.shape[-1] // 2
    n = 2 * (r - 1)
    C, S = _dft_mats(n, g2r.dtype)
    return g2r[..., :r] @ C.T - g2r[..., r:] @ S.T


# --------------------------------------------------------------------------- helpers
def _dot(a, b):
    return (a * b).sum(-1, keepdim=True)


def softplus(x):
    """F.softplus default beta=1, threshold=20 (models/complexhyperbolic.py:81)."""
    return torch.where(x > 20, x, torch.log1p(torch.exp(torch.clamp(x, max=20.0))))


def softplus_bwd(x, g):
    return torch.where(x > 20, g, g * torch.sigmoid(x))


# --------------------------------------------------------------------------- O4 project
def project(x, c):
    """utils/complexhyperbolic.py:72-87."""
    nrm = torch.sqrt(_dot(x, x)).clamp_min(MIN_NORM)
    maxnorm = (1 - PROJ_EPS) / (c ** 0.5)
    cond = nrm > maxnorm
    return torch.where(cond, x / nrm * maxnorm, x)


def project_bwd(x, c, g):
    """Returns (g_x, g_c[...,1]).  where() routes the gradient to the selected branch only."""
    raw = torch.sqrt(_dot(x, x))
    nrm = raw.clamp_min(MIN_NORM)
    maxnorm = (1 - PROJ_EPS) / (c ** 0.5)
    cond = nrm > maxnorm
    gx_dot = _dot(g, x)
    live = (raw >= MIN_NORM).to(x.dtype)
    g_proj = g * (maxnorm / nrm) - live * gx_dot * maxnorm / nrm ** 3 * x
    g_c_proj = -(gx_dot / nrm) * maxnorm / (2 * c)
    g_x = torch.where(cond, g_proj, g)
    g_c = torch.where(cond, g_c_proj, torch.zeros_like(g_c_proj))
    return g_x, g_c


# --------------------------------------------------------------------------- O3 expmap0
def expmap0(u, c):
    """utils/complexhyperbolic.py:41-54 with tanh clamp :36-37."""
    sc = c ** 0.5
    nu = torch.sqrt(_dot(u, u)).clamp_min(MIN_NORM)
    gamma = torch.tanh((sc * nu).clamp(-15, 15)) * u / (sc * nu)
    return project(gamma, c)


def expmap0_bwd(u, c, g):
    sc = c ** 0.5
    raw = torch.sqrt(_dot(u, u))
    nu = raw.clamp_min(MIN_NORM)
    a = sc * nu
    th = torch.tanh(a.clamp(-15, 15))
    f = th / a
    gamma = f * u
    g_gamma, g_c = project_bwd(gamma, c, g)
    inside = ((a >= -15) & (a <= 15)).to(u.dtype)
    fprime = inside * (1 - th * th) / a - f / a        # d f / d a
    gu_dot = _dot(g_gamma, u)
    live = (raw >= MIN_NORM).to(u.dtype)
    g_u = f * g_gamma + live * gu_dot * fprime * sc / nu * u
    g_c = g_c + gu_dot * fprime * nu / (2 * sc)
    return g_u, g_c


# --------------------------------------------------------------------------- O5 mobius
def mobius_add(x, y, c):
    """real_mobius_add, utils/complexhyperbolic.py:90-106."""
    x2, y2, xy = _dot(x, x), _dot(y, y), _dot(x, y)
    num = (1 + 2 * c * xy + c * y2) * x + (1 - c * x2) * y
    den = 1 + 2 * c * xy + c ** 2 * x2 * y2
    return num / den.clamp_min(MIN_NORM)


def mobius_add_bwd(x, y, c, g):
    x2, y2, xy = _dot(x, x), _dot(y, y), _dot(x, y)
    A = 1 + 2 * c * xy + c * y2
    Bc = 1 - c * x2
    den_raw = 1 + 2 * c * xy + c ** 2 * x2 * y2
    den = den_raw.clamp_min(MIN_NORM)
    num = A * x + Bc * y
    g_num = g / den
    g_den = -_dot(g, num) / den ** 2 * (den_raw >= MIN_NORM).to(x.dtype)
    g_A = _dot(g_num, x)
    g_B = _dot(g_num, y)
    g_xy = 2 * c * (g_A + g_den)
    g_x2 = -c * g_B + g_den * c ** 2 * y2
    g_y2 = c * g_A + g_den * c ** 2 * x2
    g_x = A * g_num + g_xy * y + 2 * g_x2 * x
    g_y = Bc * g_num + g_xy * x + 2 * g_y2 * y
    g_c = g_A * (2 * xy + y2) - g_B * x2 + g_den * (2 * xy + 2 * c * x2 * y2)
    return g_x, g_y, g_c


# --------------------------------------------------------------------------- O6 / O7 givens
def _pairs(t):
    return t.reshape(*t.shape[:-1], -1, 2)


def givens_rotations(rd, x):
    """utils/euclidean.py:39-42,55-57 (scale=None): out0=g0x0-g1x1, out1=g0x1+g1x0; no eps in the norm."""
    g = _pairs(rd)
    g = g / torch.sqrt((g * g).sum(-1, keepdim=True))
    xp = _pairs(x)
    o0 = g[..., 0] * xp[..., 0] - g[..., 1] * xp[..., 1]
    o1 = g[..., 0] * xp[..., 1] + g[..., 1] * xp[..., 0]
    return torch.stack((o0, o1), -1).reshape(x.shape)


def _norm_bwd(graw, g_ghat):
    rho = torch.sqrt((graw * graw).sum(-1, keepdim=True))
    ghat = graw / rho
    return (g_ghat - (g_ghat * ghat).sum(-1, keepdim=True) * ghat) / rho


def givens_rotations_bwd(rd, x, go):
    graw = _pairs(rd)
    g = graw / torch.sqrt((graw * graw).sum(-1, keepdim=True))
    xp, gp = _pairs(x), _pairs(go)
    gx0 = g[..., 0] * gp[..., 0] + g[..., 1] * gp[..., 1]
    gx1 = -g[..., 1] * gp[..., 0] + g[..., 0] * gp[..., 1]
    gg0 = gp[..., 0] * xp[..., 0] + gp[..., 1] * xp[..., 1]
    gg1 = -gp[..., 0] * xp[..., 1] + gp[..., 1] * xp[..., 0]
    g_rd = _norm_bwd(graw, torch.stack((gg0, gg1), -1))
    return g_rd.reshape(rd.shape), torch.stack((gx0, gx1), -1).reshape(x.shape)


def givens_reflection(rd, x):
    """utils/euclidean.py:60-75 AS CODED: out0=g0x0+g1x1, out1=(g1-g0)x0 — not a reflection."""
    g = _pairs(rd)
    g = g / torch.sqrt((g * g).sum(-1, keepdim=True))
    xp = _pairs(x)
    o0 = g[..., 0] * xp[..., 0] + g[..., 1] * xp[..., 1]
    o1 = g[..., 0] * (-xp[..., 0]) + g[..., 1] * xp[..., 0]
    return torch.stack((o0, o1), -1).reshape(x.shape)


def givens_reflection_bwd(rd, x, go):
    graw = _pairs(rd)
    g = graw / torch.sqrt((graw * graw).sum(-1, keepdim=True))
    xp, gp = _pairs(x), _pairs(go)
    gx0 = g[..., 0] * gp[..., 0] + (g[..., 1] - g[..., 0]) * gp[..., 1]
    gx1 = g[..., 1] * gp[..., 0]
    gg0 = gp[..., 0] * xp[..., 0] - gp[..., 1] * xp[..., 0]
    gg1 = gp[..., 0] * xp[..., 1] + gp[..., 1] * xp[..., 0]
    g_rd = _norm_bwd(graw, torch.stack((gg0, gg1), -1))
    return g_rd.reshape(rd.shape), torch.stack((gx0, gx1), -1).reshape(x.shape)


# --------------------------------------------------------------------------- Q1-Q3 query transform
class Params:
    """Plain container of the model tables."""

    def __init__(self, kind: int, rank: int, multi_c: bool, entity, rel, rel_diag, c, bh, bt,
                 context_vec=None, bias: str = "learn"):
        self.kind, self.rank, self.multi_c, self.bias = kind, rank, multi_c, bias
        self.dim = 2 * (rank - 1)
        self.entity, self.rel, self.rel_diag, self.c = entity, rel, rel_diag, c
        self.bh, self.bt, self.context_vec = bh, bt, context_vec
        self.dtype = entity.dtype

    @staticmethod
    def from_state_dict(sd: Dict[str, torch.Tensor], kind, rank, multi_c, bias="learn") -> "Params":
        g = lambda k: sd[k].detach().cpu().clone() if k in sd else None
        return Params(kind, rank, multi_c, g("entity.weight"), g("rel.weight"), g("rel_diag.weight"),
                      g("c.weight"), g("bh.weight"), g("bt.weight"), g("context_vec.weight"), bias)

    def zeros_like_grads(self) -> Dict[str, torch.Tensor]:
        out = {}
        for k in ("entity", "rel", "rel_diag", "c", "bh", "bt", "context_vec"):
            t = getattr(self, k)
            if t is not None:
                out[k] = torch.zeros_like(t)
        return out


def _curvature(p: Params, rel_idx):
    """c = softplus(c_emb[rel]) if multi_c else c.weight RAW (no softplus) — complexhyperbolic.py:81,109,146."""
    if p.multi_c:
        return softplus(p.c[rel_idx])                  # (..., 1)
    return p.c.reshape(1, 1).expand(*rel_idx.shape, 1)


def query_fwd(p: Params, head_idx: torch.Tensor, rel_idx: torch.Tensor, save: Optional[dict] = None):
    """get_queries for FFTRotH :79-101, FFTRefH :107-127, FFTAttH :144-171 (models/complexhyperbolic.py).

    head_idx, rel_idx: int64 of identical arbitrary shape S.  Returns (q [S,2r], c [S,1])."""
    n = p.dim
    c = _curvature(p, rel_idx)
    u = irfft_ortho(p.entity[head_idx])
    relrow = p.rel[rel_idx]
    rd = p.rel_diag[rel_idx]
    st = dict(c=c, u=u, relrow=relrow, rd=rd, head_idx=head_idx, rel_idx=rel_idx)
    if p.kind == ROT:
        hu = expmap0(u, c)
        t1 = expmap0(relrow[..., :n], c)
        t2 = expmap0(relrow[..., n:], c)
        m1 = mobius_add(hu, t1, c)
        lhs = project(m1, c)
        res1 = givens_rotations(rd, lhs)
        v = mobius_add(res1, t2, c)
        st.update(hu=hu, t1=t1, t2=t2, m1=m1, lhs=lhs, res1=res1)
    elif p.kind == REF:
        t = expmap0(relrow[..., :n], c)
        refl = givens_reflection(rd, u)
        lhs = expmap0(refl, c)
        m1 = mobius_add(lhs, t, c)
        v = project(m1, c)
        st.update(t=t, refl=refl, lhs=lhs, m1=m1)
    else:
        ctx = p.context_vec[rel_idx]
        rot_q = givens_rotations(rd[..., :n], u)
        ref_q = givens_reflection(rd[..., n:], u)
        scale = 1.0 / np.sqrt(p.rank)                   # complexhyperbolic.py:138 — rank, not dim
        l_ref = _dot(ctx * ref_q, torch.full_like(ref_q, scale))
        l_rot = _dot(ctx * rot_q, torch.full_like(rot_q, scale))
        mx = torch.maximum(l_ref, l_rot)
        e_ref, e_rot = torch.exp(l_ref - mx), torch.exp(l_rot - mx)
        w_ref, w_rot = e_ref / (e_ref + e_rot), e_rot / (e_ref + e_rot)
        att = w_ref * ref_q + w_rot * rot_q
        lhs = expmap0(att, c)
        t = expmap0(relrow[..., :n], c)
        m1 = mobius_add(lhs, t, c)
        v = project(m1, c)
        st.update(ctx=ctx, rot_q=rot_q, ref_q=ref_q, w_ref=w_ref, w_rot=w_rot, att=att, lhs=lhs, t=t,
                  m1=m1, scale=scale)
    q = rfft_ortho(v)
    if save is not None:
        save.update(st)
    return q, c


def query_bwd(p: Params, st: dict, g_q: torch.Tensor, grads: Dict[str, torch.Tensor]) -> None:
    """Hand-derived adjoint of query_fwd; accumulates DENSE grads (what autograd leaves in .grad)."""
    n = p.dim
    c, u, relrow, rd = st["c"], st["u"], st["relrow"], st["rd"]
    g_v = rfft_ortho_bwd(g_q)
    g_rel = torch.zeros_like(relrow)
    if p.kind == ROT:
        g_res1, g_t2, gc = mobius_add_bwd(st["res1"], st["t2"], c, g_v)
        g_rd, g_lhs = givens_rotations_bwd(rd, st["lhs"], g_res1)
        g_m1, gc2 = project_bwd(st["m1"], c, g_lhs)
        g_hu, g_t1, gc3 = mobius_add_bwd(st["hu"], st["t1"], c, g_m1)
        g_u, gc4 = expmap0_bwd(u, c, g_hu)
        g_r1, gc5 = expmap0_bwd(relrow[..., :n], c, g_t1)
        g_r2, gc6 = expmap0_bwd(relrow[..., n:], c, g_t2)
        g_rel = torch.cat((g_r1, g_r2), -1)
        g_c = gc + gc2 + gc3 + gc4 + gc5 + gc6
        g_ctx = None
    elif p.kind == REF:
        g_m1, gc = project_bwd(st["m1"], c, g_v)
        g_lhs, g_t, gc2 = mobius_add_bwd(st["lhs"], st["t"], c, g_m1)
        g_refl, gc3 = expmap0_bwd(st["refl"], c, g_lhs)
        g_rd, g_u = givens_reflection_bwd(rd, u, g_refl)
        g_r1, gc4 = expmap0_bwd(relrow[..., :n], c, g_t)
        g_rel[..., :n] = g_r1
        g_c = gc + gc2 + gc3 + gc4
        g_ctx = None
    else:
        g_m1, gc = project_bwd(st["m1"], c, g_v)
        g_lhs, g_t, gc2 = mobius_add_bwd(st["lhs"], st["t"], c, g_m1)
        g_att, gc3 = expmap0_bwd(st["att"], c, g_lhs)
        g_r1, gc4 = expmap0_bwd(relrow[..., :n], c, g_t)
        g_rel[..., :n] = g_r1
        g_c = gc + gc2 + gc3 + gc4
        w_ref, w_rot, ref_q, rot_q, ctx, scale = (st[k] for k in ("w_ref", "w_rot", "ref_q", "rot_q", "ctx", "scale"))
        gw_ref, gw_rot = _dot(g_att, ref_q), _dot(g_att, rot_q)
        avg = w_ref * gw_ref + w_rot * gw_rot
        gl_ref, gl_rot = w_ref * (gw_ref - avg), w_rot * (gw_rot - avg)
        g_refq = w_ref * g_att + gl_ref * scale * ctx
        g_rotq = w_rot * g_att + gl_rot * scale * ctx
        g_ctx = scale * (gl_ref * ref_q + gl_rot * rot_q)
        g_rd_rot, g_u1 = givens_rotations_bwd(rd[..., :n], u, g_rotq)
        g_rd_ref, g_u2 = givens_reflection_bwd(rd[..., n:], u, g_refq)
        g_rd = torch.cat((g_rd_rot, g_rd_ref), -1)
        g_u = g_u1 + g_u2
    g_ent = irfft_ortho_bwd(g_u)
    hi, ri = st["head_idx"].reshape(-1), st["rel_idx"].reshape(-1)
    grads["entity"].index_add_(0, hi, g_ent.reshape(-1, g_ent.shape[-1]))
    grads["rel"].index_add_(0, ri, g_rel.reshape(-1, g_rel.shape[-1]))
    grads["rel_diag"].index_add_(0, ri, g_rd.reshape(-1, g_rd.shape[-1]))
    if g_ctx is not None:
        grads["context_vec"].index_add_(0, ri, g_ctx.reshape(-1, g_ctx.shape[-1]))
    if p.multi_c:
        grads["c"].index_add_(0, ri, softplus_bwd(p.c[st["rel_idx"]], g_c).reshape(-1, 1))
    else:
        grads["c"] += g_c.sum().reshape(1, 1)


# --------------------------------------------------------------------------- S / S' distance and score
def distance_fwd(q: torch.Tensor, w: torch.Tensor, save: Optional[dict] = None) -> torch.Tensor:
    """Distance.forward, lift=True semantics (utils/complexhyperbolic.py:212-237, :176-178, :187-188).

    q [..., 2r], w [..., 2r] broadcastable.  Returns acosh(x) with x clamped at 1+eps."""
    eps = BALL_EPS[q.dtype]
    r = q.shape[-1] // 2
    zr, zi, wr, wi = q[..., :r], q[..., r:], w[..., :r], w[..., r:]
    re = (zr * wr + zi * wi).sum(-1, keepdim=True) - 1       # Re sum z conj(w) - 1
    im = (zi * wr - zr * wi).sum(-1, keepdim=True)
    zn = ((zr * zr + zi * zi).sum(-1, keepdim=True) - 1).clamp(-1, -eps)
    wn = ((wr * wr + wi * wi).sum(-1, keepdim=True) - 1).clamp(-1, -eps)
    x = 2 * (re * re + im * im) / zn / wn - 1
    x = x.clamp_min(1 + eps)
    if save is not None:
        save.update(re=re, im=im, zn=zn, wn=wn, x=x)
    return torch.acosh(x)


def distance_bwd(q, w, st, g):
    """Distance.backward/grad (utils/complexhyperbolic.py:192-210,239-254): straight-through at clamps.

    Returns per-pair grads (g_q_pairs, g_w_pairs) with the broadcast shape; caller reduces."""
    eps = BALL_EPS[q.dtype]
    r = q.shape[-1] // 2
    zr, zi, wr, wi = q[..., :r], q[..., r:], w[..., :r], w[..., r:]
    re, im, zn, wn, x = st["re"], st["im"], st["zn"], st["wn"], st["x"]
    sq = torch.sqrt(x * x - 1)
    mod2 = re * re + im * im
    pz = (sq * zn * zn * wn).clamp(max=-eps)
    pw = (sq * wn * wn * zn).clamp(max=-eps)
    # zw * w = (re + i im)(wr + i wi); wz * z = (re - i im)(zr + i zi)
    gz_r = 4 * (zn * (re * wr - im * wi) - mod2 * zr) / pz
    gz_i = 4 * (zn * (re * wi + im * wr) - mod2 * zi) / pz
    gw_r = 4 * (wn * (re * zr + im * zi) - mod2 * wr) / pw
    gw_i = 4 * (wn * (re * zi - im * zr) - mod2 * wi) / pw
    return g * torch.cat((gz_r, gz_i), -1), g * torch.cat((gw_r, gw_i), -1)


def score_pairs(p: Params, q, bh_vals, tails, save: Optional[dict] = None):
    """KGModel.score on gathered tails (models/base.py:148-173): (bh + bt) + (-d^2).

    q [B,1,2r] or [B,nt,2r]; bh_vals [B,1,1]; tails int64 [B,nt]."""
    w = p.entity[tails]
    st = {}
    d = distance_fwd(q, w, st)
    s = -(d * d)
    if p.bias == "learn":
        s = (bh_vals + p.bt[tails]) + s
    if save is not None:
        save.update(st, d=d, w=w, q=q, tails=tails)
    return s


def score_all(p: Params, q, bh_vals, chunk: int = 0) -> torch.Tensor:
    """score(q, get_rhs(None)) (models/base.py:243,255): q [b,1,2r] vs the whole table -> [b,N,1].

    Broadcast-multiply-sum like the reference (materialises b x N x r temporaries); ``chunk``>0 bounds
    the temporary by tiling over entities."""
    N = p.entity.shape[0]
    outs = []
    step = chunk if chunk > 0 else N
    for lo in range(0, N, step):
        w = p.entity[lo:lo + step].unsqueeze(0)
        d = distance_fwd(q, w)
        s = -(d * d)
        if p.bias == "learn":
            s = (bh_vals + p.bt[lo:lo + step].unsqueeze(0)) + s
        outs.append(s)
    return torch.cat(outs, 1)


# --------------------------------------------------------------------------- K / M ranking + metrics
def get_ranking(p: Params, queries: torch.Tensor, filters: Dict[Tuple[int, int], List[int]],
                batch_size: int = 500, chunk: int = 0) -> torch.Tensor:
    """KGModel.get_ranking (models/base.py:228-280): rank = 1 + #{e not in filter U {t}: s_e >= s_t}.

    Does NOT mutate ``filters`` (the reference appends the target to the stored list, :267)."""
    nq = queries.shape[0]
    ranks = torch.ones(nq)
    for b0 in range(0, nq, batch_size):
        qs = queries[b0:b0 + batch_size]
        q, _ = query_fwd(p, qs[:, 0], qs[:, 1])
        q = q.unsqueeze(1)
        bh_vals = p.bh[qs[:, 0]].unsqueeze(1)
        scores = score_all(p, q, bh_vals, chunk)                 # [b,N,1]
        targets = score_pairs(p, q, bh_vals, qs[:, 2:3])         # [b,1,1]
        assert not torch.isnan(scores).any() and not torch.isnan(targets).any()
        for i, row in enumerate(qs.numpy()):
            out = list(filters[(int(row[0]), int(row[1]))]) + [int(row[2])]
            scores[i, out] = -1e6
        ranks[b0:b0 + batch_size] += (scores >= targets).to(torch.float32).sum(1).squeeze(-1)
    return ranks


def metrics_from_ranks(ranks: torch.Tensor):
    """models/base.py:305-310."""
    return (ranks.mean().item(), (1.0 / ranks).mean().item(),
            torch.tensor([(ranks <= k).float().mean().item() for k in (1, 3, 10)]))


def compute_metrics(p: Params, examples: torch.Tensor, filters, n_rel2: int, batch_size: int = 500, chunk: int = 0):
    """KGModel.compute_metrics (models/base.py:282-322); n_rel2 = sizes[1] (reciprocals included)."""
    mr, mrr, hits = {}, {}, {}
    ranks = get_ranking(p, examples, filters["rhs"], batch_size, chunk)
    mr["rhs"], mrr["rhs"], hits["rhs"] = metrics_from_ranks(ranks)
    q = torch.stack([examples[:, 2], examples[:, 1] + n_rel2 // 2, examples[:, 0]], -1)
    ranks = get_ranking(p, q, filters["lhs"], batch_size, chunk)
    mr["lhs"], mrr["lhs"], hits["lhs"] = metrics_from_ranks(ranks)
    return mr, mrr, hits


# --------------------------------------------------------------------------- L loss forward + backward
def logsigmoid(x):
    return torch.minimum(x, torch.zeros_like(x)) - torch.log1p(torch.exp(-x.abs()))


def neg_sampling_loss(p: Params, batch: torch.Tensor, neg: torch.Tensor, want_grads: bool = True,
                      reg_n3: float = 0.0):
    """KGOptimizer.neg_sampling_loss + calculate_loss (optimizers/kg_optimizer.py:101-123,174-197).

    batch [B,3] int64, neg [B,neg] int64 (already sampled: contract of get_neg_samples :92-99).
    Returns (loss, dense grads dict or None).  Two forwards share the same queries, as in the reference."""
    B = batch.shape[0]
    hi, ri = batch[:, 0:1], batch[:, 1:2]
    qst: dict = {}
    q, _ = query_fwd(p, hi, ri, qst)                             # [B,1,2r]
    bh_vals = p.bh[hi]                                           # [B,1,1]
    tails = torch.cat((batch[:, 2:3], neg), 1)                   # [B,1+neg]; col 0 = positive
    sst: dict = {}
    s = score_pairs(p, q, bh_vals, tails, sst)                   # [B,1+neg,1]
    sign = torch.ones_like(s)
    sign[:, 1:] = -1
    terms = logsigmoid(sign * s)
    cnt = terms.numel()
    loss = -terms.sum() / cnt
    if reg_n3 > 0:                                               # optimizers/regularizers.py:50-58, positive call only
        f_h, f_r, f_t = p.entity[hi], p.rel[ri], p.entity[batch[:, 2:3]]
        loss = loss + reg_n3 * (f_h.abs() ** 3).sum() / B + reg_n3 * (f_r.abs() ** 3).sum() / B \
            + reg_n3 * (f_t.abs() ** 3).sum() / B
    if not want_grads:
        return loss, None
    grads = p.zeros_like_grads()
    g_s = -(sign * torch.sigmoid(-sign * s)) / cnt               # d loss / d s
    if p.bias == "learn":
        grads["bh"].index_add_(0, hi.reshape(-1), g_s.sum(1).reshape(-1, 1))
        grads["bt"].index_add_(0, tails.reshape(-1), g_s.reshape(-1, 1))
    g_d = -2 * sst["d"] * g_s
    gq_pairs, gw_pairs = distance_bwd(q, sst["w"], sst, g_d)
    grads["entity"].index_add_(0, tails.reshape(-1), gw_pairs.reshape(-1, gw_pairs.shape[-1]))
    query_bwd(p, qst, gq_pairs.sum(1, keepdim=True), grads)
    if reg_n3 > 0:
        for idx, tab, name in ((hi, p.entity, "entity"), (ri, p.rel, "rel"), (batch[:, 2:3], p.entity, "entity")):
            f = tab[idx]
            grads[name].index_add_(0, idx.reshape(-1), (3 * reg_n3 / B * f * f.abs()).reshape(-1, f.shape[-1]))
    return loss, grads
