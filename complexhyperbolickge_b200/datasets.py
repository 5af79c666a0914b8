"""Reader / writer of the reference's on-disk dataset format, host side only.

The reference's ``datasets/process.py:55-104`` turns the raw ``train/valid/test`` triple files into
``{split}.pickle`` (int64 numpy array [n, 3] of (lhs, rel, rhs)) plus ``to_skip.pickle``
(``{"lhs": {(entity, rel + n_rel): [entities]}, "rhs": {(entity, rel): [entities]}}``, sorted lists over
train ∪ valid ∪ test); ``datasets/kg_dataset.py:18-73`` loads them, appends the reciprocal triples to the
training split and reports the shape.  ``KGDataset`` below keeps that interface (``get_examples``,
``get_filters``, ``get_shape``) so ``run.py``'s loading code works unchanged, and adds ``filter_indices()``:
the same filters flattened ONCE into the CSR form the ranking kernels consume (``filters.FilterIndex``) — the
dict-of-lists never has to be walked per query.  ``build_filters`` / ``write_dataset`` are the process.py side
(used to put synthetic graphs on disk in the reference's format; the real datasets are not in this image).
"""
import os
import pickle as pkl
from typing import Dict, Tuple

import numpy as np
import torch

from .filters import FilterIndex


def build_filters(examples: np.ndarray, n_relations: int) -> Tuple[Dict, Dict]:
    """datasets/process.py:55-77: (lhs_final, rhs_final) dicts of sorted, de-duplicated entity lists with python-int
    keys; rhs key (lhs, rel) -> tails, lhs key (rhs, rel + n_relations) -> heads.  Vectorised (one lexsort per side)."""
    ex = np.asarray(examples, dtype=np.int64)

    def side(key_ent, key_rel, vals):
        order = np.lexsort((vals, key_rel, key_ent))
        e, r, v = key_ent[order], key_rel[order], vals[order]
        keep = np.ones(len(v), bool)
        keep[1:] = (e[1:] != e[:-1]) | (r[1:] != r[:-1]) | (v[1:] != v[:-1])
        e, r, v = e[keep], r[keep], v[keep]
        new_key = np.ones(len(v), bool)
        new_key[1:] = (e[1:] != e[:-1]) | (r[1:] != r[:-1])
        starts = np.flatnonzero(new_key)
        ends = np.append(starts[1:], len(v))
        return {(int(e[s]), int(r[s])): v[s:t].tolist() for s, t in zip(starts, ends)}

    rhs_final = side(ex[:, 0], ex[:, 1], ex[:, 2])
    lhs_final = side(ex[:, 2], ex[:, 1] + n_relations, ex[:, 0])
    return lhs_final, rhs_final


def write_dataset(path: str, train: np.ndarray, valid: np.ndarray, test: np.ndarray, n_relations: int) -> None:
    """The files process_dataset() leaves in a dataset directory (datasets/process.py:80-104)."""
    os.makedirs(path, exist_ok=True)
    splits = {"train": train, "valid": valid, "test": test}
    for name, arr in splits.items():
        with open(os.path.join(path, name + ".pickle"), "wb") as f:
            pkl.dump(np.asarray(arr).astype("int64"), f)
    lhs, rhs = build_filters(np.concatenate([train, valid, test], 0), n_relations)
    with open(os.path.join(path, "to_skip.pickle"), "wb") as f:
        pkl.dump({"lhs": lhs, "rhs": rhs}, f)


class KGDataset(object):
    """datasets/kg_dataset.py:18-73 (same constructor, attributes and methods)."""

    def __init__(self, data_path, debug=False):
        self.data_path, self.debug = data_path, debug

        def unpickle(stem):
            with open(os.path.join(data_path, stem + ".pickle"), "rb") as fh:
                return pkl.load(fh)

        self.data = {split: unpickle(split) for split in ("train", "test", "valid")}
        self.to_skip = unpickle("to_skip")
        top = self.data["train"].max(axis=0)                 # the shape is inferred from the training split (:39-41)
        self.n_entities = int(max(top[0], top[2])) + 1
        self.n_predicates = 2 * (int(top[1]) + 1)            # reciprocal relations included
        self._findex = None

    def get_examples(self, split, rel_idx=-1):
        """Triples of a split as int64 [n, 3]; the training split gets the reciprocal triples
        (rhs, rel + R, lhs) appended (:54-60); rel_idx >= 0 keeps one relation; debug keeps the first 1000."""
        triples = np.asarray(self.data[split])
        if split == "train":
            flipped = triples[:, ::-1].copy()
            flipped[:, 1] += self.n_predicates // 2
            triples = np.concatenate([triples, flipped], axis=0)
        if rel_idx >= 0:
            triples = triples[triples[:, 1] == rel_idx]
        if self.debug:
            triples = triples[:1000]
        return torch.from_numpy(np.ascontiguousarray(triples, dtype=np.int64))

    def get_filters(self):
        """The reference's dict-of-lists (:67-69); get_ranking/compute_metrics accept it as is."""
        return self.to_skip

    def filter_indices(self) -> Dict[str, FilterIndex]:
        """{"lhs": FilterIndex, "rhs": FilterIndex}: the filters flattened once; pass to compute_metrics instead of
        get_filters() to skip the per-call dict flattening."""
        if self._findex is None:
            self._findex = {side: FilterIndex.from_dict(self.to_skip[side], self.n_predicates) for side in ("lhs", "rhs")}
        return self._findex

    def get_shape(self):
        return self.n_entities, self.n_predicates, self.n_entities
