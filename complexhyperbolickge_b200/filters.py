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

    def _normalise(self):
        """Sort and de-duplicate every list once, so that batch lookups need no per-batch sort."""
        if getattr(self, "_normalised", False):
            return
        n = len(self.keys_code)
        if n and self.vals.size:
            seg = np.repeat(np.arange(n, dtype=np.int64), np.diff(self.indptr))
            order = np.lexsort((self.vals, seg))
            seg, vals = seg[order], self.vals[order]
            keep = np.ones(vals.size, bool)
            keep[1:] = (seg[1:] != seg[:-1]) | (vals[1:] != vals[:-1])
            seg, vals = seg[keep], vals[keep]
            indptr = np.zeros(n + 1, dtype=np.int64)
            np.cumsum(np.bincount(seg, minlength=n), out=indptr[1:])
            self.indptr, self.vals = indptr, vals.astype(np.int64)
        self._normalised = True

    def device_arrays(self, device):
        """(keys_code, indptr, vals) as int64 tensors resident on `device` (uploaded once, lists sorted and unique): what
        chk_filter_lookup searches, so an evaluation batch needs no host-side CSR work and no CSR upload."""
        import torch
        self._normalise()
        cache = self.__dict__.setdefault("_dev", {})
        key = str(device)
        if key not in cache:
            up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64)).to(device)
            cache[key] = (up(self.keys_code), up(self.indptr), up(self.vals if self.vals.size else np.zeros(1, np.int64)))
        return cache[key]

    def batch_csr(self, queries: np.ndarray, strict: bool = True):
        """queries int64 [b,3] -> (indptr [b+1], idx [total]) with idx_i = unique(filter[(h,r)] ∪ {t}), sorted.

        strict=True mirrors the reference's KeyError when a query key is absent (models/base.py:266).
        O(total) vectorised work per batch: the stored lists are sorted and unique (normalised once), the true
        tail is merged in by position.  The key lookup searches with SORTED needles (numpy's binary search then
        narrows its window from one needle to the next: 30 % cheaper against a 2M-key table)."""
        self._normalise()
        queries = np.asarray(queries)
        b = queries.shape[0]
        nk = len(self.keys_code)
        code = queries[:, 0].astype(np.int64) * self.n_rel2 + queries[:, 1].astype(np.int64)
        tails = queries[:, 2].astype(np.int64)
        if nk and b:
            order = np.argsort(code, kind="stable")
            pos = np.empty(b, dtype=np.int64)
            pos[order] = np.searchsorted(self.keys_code, code[order])
            np.minimum(pos, nk - 1, out=pos)
            found = self.keys_code[pos] == code
            all_found = bool(found.all())
        else:
            pos, found, all_found = np.zeros(b, np.int64), np.zeros(b, bool), b == 0
        if strict and not all_found:
            bad = queries[np.argmin(found)]
            raise KeyError((int(bad[0]), int(bad[1])))
        if nk and b:
            starts = self.indptr[pos]
            lens = self.indptr[pos + 1] - starts
            if not all_found:
                starts = np.where(found, starts, 0)
                lens = np.where(found, lens, 0)
        else:
            starts, lens = np.zeros(b, np.int64), np.zeros(b, np.int64)
        off = np.zeros(b + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        tot = int(off[-1])
        if tot:
            ar = np.arange(tot)
            qid = np.repeat(np.arange(b, dtype=np.int64), lens)
            ent = self.vals[np.repeat(starts - off[:-1], lens) + ar]
            t_rep = np.repeat(tails, lens)
            present = np.zeros(b, bool)
            present[qid[ent == t_rep]] = True
            below = np.bincount(qid[ent < t_rep], minlength=b)          # insertion position of t inside its list
        else:
            present, below = np.zeros(b, bool), np.zeros(b, np.int64)
        add = (~present).astype(np.int64)
        indptr = np.zeros(b + 1, dtype=np.int64)
        np.cumsum(lens + add, out=indptr[1:])
        idx = np.empty(int(indptr[-1]), dtype=np.int64)
        if tot:
            shift = (np.repeat(add, lens) == 1) & (ent > t_rep)        # entries after an inserted tail move one slot
            idx[np.repeat(indptr[:-1] - off[:-1], lens) + ar + shift] = ent
        ins = np.nonzero(add)[0]
        idx[indptr[:-1][ins] + below[ins]] = tails[ins]
        return indptr, idx
