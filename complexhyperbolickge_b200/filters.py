"""Device-ready filter structure for filtered ranking.

The reference keeps ``filters[(entity, relation)] -> python list of entity ids`` (built by
datasets/process.py:55-77, loaded from to_skip.pickle) and walks it with a per-query Python loop that
issues one ``index_put_`` per query (models/base.py:264-268).  Here the dict is flattened ONCE into a
sorted key table + CSR (numpy), and each evaluation batch becomes one vectorised lookup that yields the
(indptr, idx) arrays ``chk_rank_counts`` consumes: per query the UNIQUE ids of filter ∪ {true tail}.
"""
from typing import Dict, List, Tuple

import numpy as np


class FilterIndex:
    def __init__(self, keys_code: np.ndarray, indptr: np.ndarray, vals: np.ndarray, n_rel2: int):
        self.keys_code, self.indptr, self.vals, self.n_rel2 = keys_code, indptr, vals, n_rel2

    @staticmethod
    def from_dict(filters: Dict[Tuple[int, int], List[int]], n_rel2: int) -> "FilterIndex":
        n = len(filters)
        codes = np.empty(n, dtype=np.int64)
        lens = np.empty(n, dtype=np.int64)
        lists = []
        for i, (k, v) in enumerate(filters.items()):
            codes[i] = int(k[0]) * n_rel2 + int(k[1])
            lens[i] = len(v)
            lists.append(v)
        order = np.argsort(codes, kind="stable")
        vals = np.concatenate([np.asarray(lists[i], dtype=np.int64) for i in order]) if n else np.zeros(0, np.int64)
        indptr = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(lens[order], out=indptr[1:])
        return FilterIndex(codes[order], indptr, vals, n_rel2)

    @staticmethod
    def from_arrays(keys: np.ndarray, indptr: np.ndarray, vals: np.ndarray, n_rel2: int) -> "FilterIndex":
        codes = keys[:, 0].astype(np.int64) * n_rel2 + keys[:, 1].astype(np.int64)
        order = np.argsort(codes, kind="stable")
        lens = np.diff(indptr)[order]
        new_indptr = np.zeros(len(codes) + 1, dtype=np.int64)
        np.cumsum(lens, out=new_indptr[1:])
        starts = indptr[:-1][order]
        gather = np.repeat(starts - new_indptr[:-1], lens) + np.arange(new_indptr[-1])
        return FilterIndex(codes[order], new_indptr, vals[gather].astype(np.int64), n_rel2)

    def batch_csr(self, queries: np.ndarray, strict: bool = True):
        """queries int64 [b,3] -> (indptr [b+1], idx [total]) with idx_i = unique(filter[(h,r)] ∪ {t}), sorted.

        strict=True mirrors the reference's KeyError when a query key is absent (models/base.py:266)."""
        b = queries.shape[0]
        code = queries[:, 0].astype(np.int64) * self.n_rel2 + queries[:, 1].astype(np.int64)
        pos = np.searchsorted(self.keys_code, code)
        pos_c = np.minimum(pos, max(len(self.keys_code) - 1, 0))
        found = (len(self.keys_code) > 0) & (self.keys_code[pos_c] == code) if len(self.keys_code) else np.zeros(b, bool)
        if strict and not np.all(found):
            bad = queries[np.argmin(found)]
            raise KeyError((int(bad[0]), int(bad[1])))
        starts = np.where(found, self.indptr[pos_c], 0)
        lens = np.where(found, self.indptr[pos_c + 1] - self.indptr[pos_c], 0) if len(self.keys_code) else np.zeros(b, np.int64)
        tot = int(lens.sum())
        off = np.zeros(b + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        src = np.repeat(starts - off[:-1], lens) + np.arange(tot)
        ent = np.concatenate([self.vals[src], queries[:, 2].astype(np.int64)])
        qid = np.concatenate([np.repeat(np.arange(b, dtype=np.int64), lens), np.arange(b, dtype=np.int64)])
        n_ent = int(ent.max()) + 1 if ent.size else 1
        key = np.unique(qid * n_ent + ent)
        qid_u, ent_u = key // n_ent, key % n_ent
        indptr = np.zeros(b + 1, dtype=np.int64)
        np.cumsum(np.bincount(qid_u, minlength=b), out=indptr[1:])
        return indptr, ent_u.astype(np.int64)
