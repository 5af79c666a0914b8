"""FFTRotH / FFTRefH / FFTAttH behind the reference's KGModel API, running on the chk_b200 kernels.

Mirrors (names, argument meaning, shapes, state_dict keys) the reference classes
``models/base.py:KGModel`` (:32-66 params, :108-133 get_rhs, :148-173 score, :175-198 get_factors,
:200-226 forward, :228-280 get_ranking, :282-322 compute_metrics) and
``models/complexhyperbolic.py:FFTUnitBall/FFTRotH/FFTRefH/FFTAttH`` (:17-171) with the working
``lift=True`` semantics, so that the reference's run.py / KGOptimizer can construct and drive
these classes unchanged:  ``getattr(models, args.model)(args)``, ``model(queries, tails)``,
``model.compute_metrics(examples, filters, batch_size)``, ``state_dict()`` interchange.

``get_rhs``, ``score`` and ``get_factors`` are the unsqueeze / gather plumbing of that API, which ``north_star`` says stays
unchanged: their bodies restate the reference's (models/base.py:122-133, 164-173, 184-198) statement by statement — there is
no other way to keep ``KGOptimizer.neg_sampling_loss`` and the regularisers working against these classes — and so do the
caller-contract methods in ``optim.py`` and the 40-line pickle reader in ``datasets.py``.  Everything below that surface is new.

All arithmetic of the hot path happens in hand-written sm_100a kernels reached through the C ABI
(``ops.py`` -> ``libchk_b200.so``).  There is no eager / CPU fallback: a model on a non-CUDA device raises.
"""
from abc import ABC, abstractmethod
from typing import Dict, Optional

import numpy as np
import torch
from torch import nn

from . import ops
from .filters import FilterIndex

CHYP_MODELS = ["FFTRotH", "FFTRefH", "FFTAttH"]


class KGModel(nn.Module, ABC):
    """Same constructor and attributes as the reference base class (models/base.py:32-66)."""

    def __init__(self, sizes, rank, dropout, gamma, data_type, bias, init_size):
        super().__init__()
        if data_type == "double":
            self.data_type = torch.double
        elif data_type == "float":
            self.data_type = torch.float
        else:                       # the reference leaves data_type unset and crashes later
            raise ValueError(f"dtype must be 'float' or 'double', got {data_type!r}")
        if bias not in ("learn", "none"):
            raise ValueError("bias must be 'learn' or 'none' ('constant' is broken in the reference, SURVEY §5)")
        self.sizes, self.rank, self.dropout, self.bias = sizes, rank, dropout, bias
        self.init_size, self.gamma = init_size, gamma
        self.entity = nn.Embedding(sizes[0], rank)
        self.rel = nn.Embedding(sizes[1], rank)
        self.bh = nn.Embedding(sizes[0], 1)
        self.bt = nn.Embedding(sizes[0], 1)
        with torch.no_grad():
            nn.init.normal_(self.entity.weight, 0.0, init_size)
            nn.init.normal_(self.rel.weight, 0.0, init_size)
            nn.init.zeros_(self.bh.weight)
            nn.init.zeros_(self.bt.weight)

    def __setattr__(self, name, value):
        # every embedding table is kept in data_type (models/base.py:84-94)
        if isinstance(value, nn.Embedding):
            with torch.no_grad():
                value.weight.data = value.weight.data.to(self.data_type)
        super().__setattr__(name, value)

    @abstractmethod
    def get_queries(self, queries):
        ...

    @abstractmethod
    def similarity_score(self, lhs_e, rhs_e):
        ...

    def get_rhs(self, tails=None):
        """models/base.py:108-133."""
        if tails is None:
            rhs_e, rhs_biases = self.entity.weight, self.bt.weight
            while rhs_e.dim() < 3:
                rhs_e = rhs_e.unsqueeze(0)
            while rhs_biases.dim() < 3:
                rhs_biases = rhs_biases.unsqueeze(0)
        else:
            rhs_e, rhs_biases = self.entity(tails), self.bt(tails)
            while rhs_e.dim() < 3:
                rhs_e = rhs_e.unsqueeze(1)
            while rhs_biases.dim() < 3:
                rhs_biases = rhs_biases.unsqueeze(1)
        return rhs_e, rhs_biases

    def score(self, lhs, rhs):
        """models/base.py:148-173; bias add order (bh + bt) + score."""
        lhs_e, lhs_biases = lhs
        rhs_e, rhs_biases = rhs
        score = self.similarity_score(lhs_e, rhs_e)
        if self.bias == "learn":
            return lhs_biases + rhs_biases + score
        return score

    def get_factors(self, queries, tails=None):
        """models/base.py:175-198 — raw embeddings for the regulariser (plain gathers)."""
        head_e = self.entity(queries[..., 0])
        rel_e = self.rel(queries[..., 1])
        while head_e.dim() < 3:
            head_e = head_e.unsqueeze(1)
        while rel_e.dim() < 3:
            rel_e = rel_e.unsqueeze(1)
        if tails is None:
            rhs_e = self.entity.weight
            while rhs_e.dim() < 3:
                rhs_e = rhs_e.unsqueeze(0)
        else:
            rhs_e = self.entity(tails)
            while rhs_e.dim() < 3:
                rhs_e = rhs_e.unsqueeze(1)
        return head_e, rel_e, rhs_e


class FFTUnitBall(KGModel):
    """models/complexhyperbolic.py:17-73 (lift=True semantics)."""

    KIND = None

    def __init__(self, args):
        super().__init__(args.sizes, args.rank, args.dropout, args.gamma, args.dtype, args.bias, args.init_size)
        if not ops.supported_rank(self.rank):
            raise ValueError(f"rank={self.rank}: 2(rank-1) must be a power of two in [16, 512] for the warp FFT")
        self.dim = 2 * (self.rank - 1)
        del self.entity
        self.entity = nn.Embedding(self.sizes[0], 2 * self.rank)
        del self.rel
        self.rel = nn.Embedding(self.sizes[1], 2 * self.dim)
        self.rel_diag = nn.Embedding(self.sizes[1], self.dim)
        self.multi_c = args.multi_c
        self.c = nn.Embedding(self.sizes[1] if self.multi_c else 1, 1)
        with torch.no_grad():
            nn.init.normal_(self.entity.weight, 0.0, self.init_size)
            nn.init.normal_(self.rel.weight, 0.0, self.init_size)
            nn.init.uniform_(self.rel_diag.weight, -1.0, 1.0)
            nn.init.ones_(self.c.weight)
        self.lift = True
        self.rank_algo = "auto"          # "auto" (tcgen05 tier when the table is large enough to pay for the shadow) | "fma" | "mma"
        self._param_epoch = 0            # bumped by optimizers that write parameters through raw pointers (train.py, parallel.py)
        self.process_group = None        # set to shard the entity table across ranks in get_ranking
        self._filter_cache: Dict[int, tuple] = {}
        self._eval_cache = None          # (key, ranking.EvalState): shard view, Hermitian norms, bf16 shadow
        self._eval_ws = None
        self.fused_forward = True

    # ------------------------------------------------------------------ pieces of the reference API
    def _ctx_weight(self):
        return None

    def _meta(self):
        return (self.KIND, self.rank, bool(self.multi_c))

    def get_queries(self, queries):
        """models/complexhyperbolic.py:79-101 / 107-127 / 144-171 — one fused kernel (csrc/chk_query.cu)."""
        lead = queries.shape[:-1]
        head_idx = queries[..., 0].reshape(-1).contiguous()
        rel_idx = queries[..., 1].reshape(-1).contiguous()
        q, c = ops.QueryTransformFn.apply(self._meta(), self.entity.weight, self.rel.weight, self.rel_diag.weight,
                                          self._ctx_weight(), self.c.weight, head_idx, rel_idx)
        res = q.view(*lead, 2 * self.rank)
        c = c.view(*lead, 1) if self.multi_c else self.c.weight
        lhs_biases = self.bh(queries[..., 0])
        while res.dim() < 3:
            res = res.unsqueeze(1)
        while c.dim() < 3:
            c = c.unsqueeze(1)
        while lhs_biases.dim() < 3:
            lhs_biases = lhs_biases.unsqueeze(1)
        return (res, c), lhs_biases

    def similarity_score(self, lhs_e, rhs_e):
        """-Distance(lhs, rhs)^2 (models/complexhyperbolic.py:45-59); the curvature in the tuple is unused."""
        lhs_e, _c = lhs_e
        if (not torch.is_grad_enabled() or not (lhs_e.requires_grad or rhs_e.requires_grad)) \
                and rhs_e.shape[0] == 1 and lhs_e.shape[1] == 1:
            q = lhs_e.reshape(-1, 2 * self.rank).contiguous()
            table = rhs_e[0].contiguous()
            s = ops.score_all(self.rank, q, ops.row_hnorm(self.rank, q), None, table, ops.row_hnorm(self.rank, table),
                              None)
            return s.unsqueeze(-1)
        return ops.ScoreRowsFn.apply(self.rank, lhs_e, rhs_e)

    def forward(self, queries, tails=None):
        """models/base.py:200-226.  With tails given (training) K1 + gather-scoring run fused."""
        while queries.dim() < 3:
            queries = queries.unsqueeze(1)
        if tails is not None:
            while tails.dim() < 2:
                tails = tails.unsqueeze(0)
        if tails is not None and self.fused_forward and queries.shape[1] in (1, tails.shape[1]):
            meta = (self.KIND, self.rank, bool(self.multi_c), self.bias == "learn")
            predictions = ops.FusedForwardFn.apply(meta, self.entity.weight, self.rel.weight, self.rel_diag.weight,
                                                   self._ctx_weight(), self.c.weight, self.bh.weight, self.bt.weight,
                                                   queries, tails)
        else:
            lhs_e, lhs_biases = self.get_queries(queries)
            rhs_e, rhs_biases = self.get_rhs(tails)
            predictions = self.score((lhs_e, lhs_biases), (rhs_e, rhs_biases))
        factors = self.get_factors(queries, tails)
        return predictions, factors

    # ------------------------------------------------------------------ evaluation
    def _filter_index(self, filters) -> FilterIndex:
        if isinstance(filters, FilterIndex):
            return filters
        key = id(filters)
        hit = self._filter_cache.get(key)
        # the entry keeps the dict alive, so its id cannot be recycled for another dict while it is cached
        if hit is None or hit[0] is not filters or hit[1] != len(filters):
            if len(self._filter_cache) >= 4:                     # rhs/lhs of valid + test; bounded
                self._filter_cache.pop(next(iter(self._filter_cache)))
            hit = self._filter_cache[key] = (filters, len(filters), FilterIndex.from_dict(filters, self.sizes[1]))
        return hit[2]

    def get_ranking(self, queries, filters, batch_size=500):
        """Filtered ranks of the true tails (models/base.py:228-280), float32 CPU tensor [n].

        rank = 1 + #{e not in filters[(h,r)] ∪ {t} : score(e) >= score(t)} — counted on the GPU by
        chk_rank_counts; no (b, N) score matrix and no per-query host loop.  Unlike the reference the
        filter lists are NOT mutated.  With ``self.process_group`` set, every rank counts over its
        contiguous slice of the entity table and the int64 counts are summed with one all_reduce."""
        from .ranking import rank_queries
        if getattr(self, "_replicas_stale", False):
            raise RuntimeError("the entity / bias tables are owner-sharded by FusedDataParallelKGOptimizer and this rank's copy is not "
                               "current: call optimizer.sync_replicas() (epoch() does) before evaluating or saving the model")
        if isinstance(queries, np.ndarray):
            queries = torch.from_numpy(queries)
        return rank_queries(self, queries, self._filter_index(filters), batch_size)

    def parameters_changed(self):
        """Tell the model its tables were written through raw device pointers (no torch ``_version`` bump): the cached
        evaluation state (Hermitian norms, bf16 shadow) is stale.  The fused optimizers call this after every step."""
        self._param_epoch += 1

    def resolved_rank_algo(self) -> str:
        """"auto": the tcgen05 tier when it is built and the table is big enough that the contraction, not the per-batch
        prologue, dominates (N >= 8192 rows; below that the exact FMA tier finishes a 500-query batch in < 0.1 ms)."""
        if self.rank_algo != "auto":
            return self.rank_algo
        return "mma" if (self.sizes[0] >= 8192 and ops.mma_available(self.rank)) else "fma"

    def release_eval_cache(self):
        """Drop the cached evaluation state (entity shadow, norms, workspace) to give the memory back."""
        self._eval_cache = None
        self._eval_ws = None

    def compute_metrics(self, examples, filters, batch_size=10):
        """models/base.py:282-322."""
        mean_rank, mean_reciprocal_rank, hits_at = {}, {}, {}
        if isinstance(examples, tuple):
            examples = examples[0]
        if isinstance(examples, np.ndarray):
            examples = torch.from_numpy(examples)
        sides = (("rhs", examples),
                 ("lhs", torch.stack([examples[..., 2], examples[..., 1] + self.sizes[1] // 2, examples[..., 0]], -1)))
        for side, q in sides:
            ranks = self.get_ranking(q, filters[side], batch_size=batch_size)
            mean_rank[side] = torch.mean(ranks).item()
            mean_reciprocal_rank[side] = torch.mean(1.0 / ranks).item()
            hits_at[side] = torch.FloatTensor([torch.mean((ranks <= k).float()).item() for k in (1, 3, 10)])
        return mean_rank, mean_reciprocal_rank, hits_at


class FFTRotH(FFTUnitBall):
    """Hyperbolic Givens rotations in the Fourier-dual space (models/complexhyperbolic.py:76-101)."""
    KIND = ops.CHK_ROT


class FFTRefH(FFTUnitBall):
    """Hyperbolic Givens "reflections", formula as coded in the reference (models/complexhyperbolic.py:104-127,
    utils/euclidean.py:60-75)."""
    KIND = ops.CHK_REF


class FFTAttH(FFTUnitBall):
    """Attention over rotation / reflection candidates (models/complexhyperbolic.py:130-171)."""
    KIND = ops.CHK_ATT

    def __init__(self, args):
        super().__init__(args)
        self.rel_diag = nn.Embedding(self.sizes[1], 2 * self.dim)
        self.context_vec = nn.Embedding(self.sizes[1], self.dim)
        self.scale = 1.0 / np.sqrt(self.rank)
        with torch.no_grad():
            nn.init.uniform_(self.rel_diag.weight, -1.0, 1.0)
            nn.init.normal_(self.context_vec.weight, 0.0, self.init_size)

    def _ctx_weight(self):
        return self.context_vec.weight
