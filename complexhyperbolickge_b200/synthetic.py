"""Seeded synthetic knowledge graphs of the shapes named in BASELINE.json (there is no network for the real
datasets).  Triples: Zipf(1.0)-popular heads and tails, uniform relation; deduplicated; split into
train / valid / test; filters built over all splits with the semantics of the reference's
datasets/process.py:55-77 (rhs key (h, r) -> tails, lhs key (t, r + R) -> heads), directly as CSR."""
from typing import Dict, Tuple

import numpy as np
import torch

from .filters import FilterIndex

SHAPES = {
    # name: (n_ent, n_rel, n_train, n_valid, n_test)           BASELINE.json configs / SURVEY §8d
    "wn18rr": (40_943, 11, 86_835, 3_034, 3_134),
    "fb237": (14_541, 237, 272_115, 17_535, 20_466),
    "yago310": (123_182, 37, 1_079_040, 5_000, 5_000),
    "big4m": (4_000_000, 1_000, 2_000_000, 10_000, 10_000),
}


def zipf_sample(rng, n_ent: int, size: int) -> np.ndarray:
    w = 1.0 / np.arange(1, n_ent + 1, dtype=np.float64)
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    ids = np.searchsorted(cdf, rng.random(size))
    perm = rng.permutation(n_ent)           # popularity is not correlated with the id
    return perm[np.minimum(ids, n_ent - 1)]


def _csr(keys_a, keys_b, vals, n_rel2):
    code = keys_a.astype(np.int64) * n_rel2 + keys_b.astype(np.int64)
    order = np.lexsort((vals, code))
    code, vals = code[order], vals[order]
    keep = np.ones(len(code), bool)
    keep[1:] = (code[1:] != code[:-1]) | (vals[1:] != vals[:-1])
    code, vals = code[keep], vals[keep]
    ucode, start = np.unique(code, return_index=True)
    indptr = np.concatenate([start, [len(code)]]).astype(np.int64)
    return FilterIndex(ucode, indptr, vals.astype(np.int64), n_rel2)


def make_graph(shape: str, seed: int = 0, n_train: int = None) -> Dict:
    n_ent, n_rel, ntr, nva, nte = SHAPES[shape]
    if n_train is not None:
        ntr = n_train
    rng = np.random.default_rng(seed)
    tot = ntr + nva + nte
    over = int(tot * 1.25) + 1000
    tri = np.stack([zipf_sample(rng, n_ent, over), rng.integers(0, n_rel, over), zipf_sample(rng, n_ent, over)], 1)
    tri = np.unique(tri, axis=0)
    rng.shuffle(tri)
    tri = tri[:tot].astype(np.int64)
    train, valid, test = tri[:ntr], tri[ntr:ntr + nva], tri[ntr + nva:]
    n_rel2 = 2 * n_rel
    filters = {"rhs": _csr(tri[:, 0], tri[:, 1], tri[:, 2], n_rel2),
               "lhs": _csr(tri[:, 2], tri[:, 1] + n_rel, tri[:, 0], n_rel2)}
    return dict(n_ent=n_ent, n_rel2=n_rel2, train=train, valid=valid, test=test, filters=filters)


def train_examples(graph) -> torch.Tensor:
    """Reciprocal-triple augmentation of datasets/kg_dataset.py:54-60."""
    tr = graph["train"]
    inv = tr[:, [2, 1, 0]].copy()
    inv[:, 1] += graph["n_rel2"] // 2
    return torch.from_numpy(np.vstack([tr, inv]))


def filter_dict(fi: FilterIndex) -> Dict[Tuple[int, int], list]:
    """CSR -> the reference's dict-of-lists format (small graphs / oracle only)."""
    out = {}
    for i, code in enumerate(fi.keys_code):
        out[(int(code // fi.n_rel2), int(code % fi.n_rel2))] = fi.vals[fi.indptr[i]:fi.indptr[i + 1]].tolist()
    return out


def trained_like_(model, seed: int = 0):
    """'Trained-like' weight regime of SURVEY §8c so that scores are not all clamped."""
    r = model.rank
    dev, dt = model.entity.weight.device, model.entity.weight.dtype
    g = torch.Generator(device=dev).manual_seed(seed)      # same device type + seed => same weights on every rank
    with torch.no_grad():
        def fill(t, std=None, lo=None, hi=None):
            if std is not None:
                t.copy_(torch.randn(t.shape, generator=g, device=dev, dtype=torch.float32) * std)
            else:
                t.copy_(torch.rand(t.shape, generator=g, device=dev, dtype=torch.float32) * (hi - lo) + lo)
        fill(model.entity.weight, std=float(np.sqrt(0.4 / (2 * r))))
        fill(model.rel.weight, std=0.05)
        fill(model.rel_diag.weight, lo=-1.0, hi=1.0)
        fill(model.c.weight, lo=0.5, hi=2.0)
        fill(model.bh.weight, std=0.1)
        fill(model.bt.weight, std=0.1)
        if hasattr(model, "context_vec"):
            fill(model.context_vec.weight, std=1.0)
    return model
