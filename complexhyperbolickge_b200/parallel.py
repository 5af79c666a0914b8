"""Multi-GPU plumbing of the hot path: one process per GPU, ``torch.distributed`` (NCCL over
NVLink on the box, gloo in the CPU tests).  The reference has no distributed code at all; this is the new
functionality ``north_star`` defines:

* filtered ranking: entity-table row shards + ONE int64 ``all_reduce`` of the rank counts (``ranking.py``);
* training: data parallel with replicated parameters.  Each rank runs the kernels on its slice of the
  batch; the gradients of the big row-sparse tables (entity, bh, bt) are exchanged as (row id, gradient row)
  pairs with ``all_gather`` — a dense all_reduce of the N x 2r gradient (8.2 GB per step at the 4M-entity
  config) is never done — and the small relation tables (rel, rel_diag, context_vec, c) with a dense
  ``all_reduce``.  Every rank then holds the identical averaged dense ``.grad`` and applies the identical
  ``torch.optim`` update (dense-equivalent semantics); contributions are accumulated in rank
  order, so the result is deterministic.

Everything here is plain tensor code on whatever device the gradients live on, so the same functions run
under gloo on CPU tensors (tests/test_distributed_cpu.py) and under NCCL on the GPUs.
"""
from typing import Dict, Iterable, List, Optional

import torch
import torch.distributed as dist

from . import ops
from .optim import KGOptimizer
from .train import FusedKGOptimizer

SPARSE_TABLES = ("entity", "bh", "bt")


def _world(pg) -> int:
    return dist.get_world_size(pg) if (dist.is_available() and dist.is_initialized()) else 1


def exchange_sparse_rows(grad: torch.Tensor, rows: torch.Tensor, pg=None) -> torch.Tensor:
    """Average a row-sparse dense gradient over the ranks of ``pg`` by exchanging only the touched rows.

    grad [N, w] holds this rank's contribution (non-zero only in ``rows``); rows int64 [m] (duplicates
    allowed).  On return ``grad`` holds (sum over ranks of the contributions) / world, identical on every rank.
    """
    world = _world(pg)
    if world == 1:
        return grad
    rows = torch.unique(rows)
    vals = grad.index_select(0, rows) / world      # only the touched rows are scaled: no pass over the N x w table
    n_local = torch.tensor([rows.numel()], dtype=torch.int64, device=grad.device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=pg)
    m = int(max(c.item() for c in counts))
    pad_rows = torch.zeros(m, dtype=torch.int64, device=grad.device)
    pad_vals = torch.zeros((m,) + tuple(grad.shape[1:]), dtype=grad.dtype, device=grad.device)
    pad_rows[: rows.numel()] = rows
    pad_vals[: rows.numel()] = vals
    all_rows = [torch.empty_like(pad_rows) for _ in range(world)]
    all_vals = [torch.empty_like(pad_vals) for _ in range(world)]
    dist.all_gather(all_rows, pad_rows, group=pg)
    dist.all_gather(all_vals, pad_vals, group=pg)
    grad.index_fill_(0, rows, 0)                   # drop the local contribution, then add everyone's in rank order
    for k in range(world):
        nk = int(counts[k].item())
        if nk:
            grad.index_add_(0, all_rows[k][:nk], all_vals[k][:nk])
    return grad


def allreduce_dense(grads: Iterable[torch.Tensor], pg=None) -> None:
    world = _world(pg)
    if world == 1:
        return
    for g in grads:
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=pg)
        g.div_(world)


def reduce_gradients(model, touched: Dict[str, torch.Tensor], pg=None) -> None:
    """Average ``model``'s gradients over the data-parallel group.  ``touched[name]`` lists the rows of the
    sparse tables this rank's step read (heads / tails / negatives)."""
    dense: List[torch.Tensor] = []
    for name, p in model.named_parameters():
        if p.grad is None:
            continue
        table = name.split(".")[0]
        if table in SPARSE_TABLES:
            exchange_sparse_rows(p.grad, touched[table], pg)
        else:
            dense.append(p.grad)
    allreduce_dense(dense, pg)


class DataParallelKGOptimizer(KGOptimizer):
    """The KGOptimizer contract (optim.py; reference optimizers/kg_optimizer.py:69-123,174-197,239-277) run data
    parallel: ``step(global_batch)`` takes the same (B, 3) batch on every rank, trains on rows
    ``rank::world`` of it and exchanges sparse gradients before the (replicated) optimizer step.  The loss is
    a mean over the local B/world * (1+neg) terms, so averaging the gradients over the ranks gives the
    gradient of the global mean when B is a multiple of world."""

    def __init__(self, *args, process_group=None, **kw):
        super().__init__(*args, **kw)
        self.pg = process_group
        self.world = _world(process_group)
        self.rank_id = dist.get_rank(process_group) if self.world > 1 else 0
        self._negs: Optional[torch.Tensor] = None

    def get_neg_samples(self, input_batch):
        self._negs = super().get_neg_samples(input_batch)
        return self._negs

    def local_slice(self, global_batch: torch.Tensor) -> torch.Tensor:
        return global_batch[self.rank_id::self.world]

    def step(self, global_batch: torch.Tensor) -> torch.Tensor:
        batch = self.local_slice(global_batch).to(self.device)
        loss = self.calculate_loss(batch)
        loss.backward()
        tails = torch.cat([batch[:, 2], self._negs.reshape(-1)])
        reduce_gradients(self.model, {"entity": torch.cat([batch[:, 0], tails]), "bh": batch[:, 0], "bt": tails}, self.pg)
        self.optimizer.step()
        self.optimizer.zero_grad()
        return loss.detach()

    def epoch(self, examples):
        """Same shuffle on every rank (seeded generator broadcast from rank 0 is the caller's job: pass the
        permuted examples or seed torch identically); returns the mean of the local losses averaged over ranks."""
        actual = examples[torch.randperm(examples.shape[0]), :]
        if self.world > 1:
            actual = actual.to(self.device)
            dist.broadcast(actual, src=dist.get_global_rank(self.pg, 0) if self.pg is not None else 0, group=self.pg)
        total = torch.zeros((), dtype=torch.float64, device=self.device)
        n = 0
        for b0 in range(0, examples.shape[0], self.batch_size):
            total += self.step(actual[b0:b0 + self.batch_size]).double()
            n += 1
        if self.world > 1:
            dist.all_reduce(total, op=dist.ReduceOp.SUM, group=self.pg)
            total /= self.world
        return (total / max(n, 1)).item()


class FusedDataParallelKGOptimizer(FusedKGOptimizer):
    """Data-parallel version of the fused step (train.FusedKGOptimizer) for ``torch.optim.Adagrad``: every rank runs
    the fused forward/backward chain on rows ``rank::world`` of the global batch (captured in a CUDA graph), then

    * the touched row ids of every rank are all_gathered (fixed size, no host sync),
    * each table's gradient is summed over the ranks — a dense all_reduce when the table is smaller than the rows
      that would have to travel (FB15k-237 / WN18RR-size entity tables, all relation tables), otherwise the sparse
      (row id, row) exchange of ``exchange_sparse_rows`` (4M-entity table) —
    * and the row-sparse Adagrad step runs on the union of the touched rows, identically on every rank.

    Local gradients are pre-scaled by 1/world, so the summed gradient is the gradient of the global mean loss."""

    def __init__(self, *args, process_group=None, sparse_exchange=None, **kw):
        """sparse_exchange: None = decide per table by size (below); True = the row-sparse tables (entity, bh, bt)
        always use the sparse row exchange (tests / small-scale checks of the big-table path)."""
        super().__init__(*args, **kw)
        self.pg = process_group
        self.world = _world(process_group)
        self.rank_id = dist.get_rank(process_group) if self.world > 1 else 0
        self.grad_scale = 1.0 / self.world
        self.local_batch_size = self.batch_size // self.world      # batch_size is the GLOBAL batch
        if not (self.fused and self.sparse_adagrad):
            raise ValueError("FusedDataParallelKGOptimizer needs torch.optim.Adagrad (lr_decay=0, weight_decay=0), a zero "
                             "regulariser weight and update_steps=1; use DataParallelKGOptimizer otherwise")
        # Tables whose dense gradient is smaller than the rows that would have to travel are summed with ONE dense
        # all_reduce: their .grad tensors are views into one flat buffer.  The others use the sparse row exchange.
        rows_per_step = self.local_batch_size * (2 + self.neg_sample_size) * self.world
        params = list(self.model.parameters())
        forced = set()
        if sparse_exchange:
            forced = {id(p) for n, p in self.model.named_parameters() if n.split(".")[0] in SPARSE_TABLES}
        self._dense = [p for p in params if id(p) not in forced and
                       p.numel() * p.element_size() <= rows_per_step * ((p.shape[1] if p.dim() > 1 else 1) * p.element_size() + 8)]
        self._sparse = [p for p in params if not any(p is d for d in self._dense)]
        if self._dense:
            flat = torch.zeros(sum(p.numel() for p in self._dense), dtype=params[0].dtype, device=params[0].device)
            o = 0
            for p in self._dense:
                p.grad = flat[o:o + p.numel()].view_as(p)
                o += p.numel()
            self._flat_grad = flat

    def _step_body(self, batch):
        """Local forward/backward, the collectives and the optimizer: every shape is static and nothing synchronises
        with the host, so the CUDA graph holds the whole step including the NCCL calls."""
        self._touched = self._forward_backward(batch)
        self._exchange_and_update()

    def _post_step(self):
        for p in self.model.parameters():
            self.optimizer.state[p]["step"] += 1

    def _exchange_and_update(self):
        m, opt = self.model, self.optimizer
        heads, rels, tails = self._touched
        tails = tails.reshape(-1)
        nb = heads.numel()
        W = self.world
        if W > 1:                                    # one all_gather for every touched row id of every rank
            ids = torch.cat([heads, rels, tails])
            allids = torch.empty((W, ids.numel()), dtype=ids.dtype, device=ids.device)
            dist.all_gather_into_tensor(allids, ids, group=self.pg)
            heads_all, rels_all, tails_all = allids[:, :nb], allids[:, nb:2 * nb], allids[:, 2 * nb:]
            if self._dense:
                dist.all_reduce(self._flat_grad, op=dist.ReduceOp.SUM, group=self.pg)
        else:
            heads_all, rels_all, tails_all = heads.view(1, -1), rels.view(1, -1), tails.view(1, -1)
        ent_all = torch.cat([heads_all, tails_all], 1)                   # per rank: [heads | tails], the order of the sent rows
        zero_row = torch.zeros((1, 1), dtype=torch.int64, device=heads.device)
        # (table, rows of every rank [W, m_k], this rank's rows)
        plan = [(m.entity.weight, ent_all, torch.cat([heads, tails])), (m.rel.weight, rels_all, rels),
                (m.rel_diag.weight, rels_all, rels), (m.c.weight, rels_all if m.multi_c else zero_row, rels if m.multi_c else zero_row.view(-1))]
        if m._ctx_weight() is not None:
            plan.append((m._ctx_weight(), rels_all, rels))
        if m.bias == "learn":
            plan += [(m.bh.weight, heads_all, heads), (m.bt.weight, tails_all, tails)]
        if W > 1 and self._sparse:
            # Big row-sparse tables (the 4M-entity table and its biases): only the touched rows travel.  Send side:
            # chk_claim_gather_rows (each row once, cleared locally); one all_gather per table; receive side: the
            # contributions are added back ONE RANK AT A TIME (a launch holds at most one non-zero contribution per
            # row), so the sum has the same bits on every replica.
            recv = []
            for p, rows_all, rows_local in plan:
                if not any(p is s_ for s_ in self._sparse):
                    continue
                sent = ops.claim_gather_rows(p.grad, rows_local.contiguous(), self._stamps[p], self._step_id)
                got = torch.empty((W,) + tuple(sent.shape), dtype=sent.dtype, device=sent.device)
                dist.all_gather_into_tensor(got, sent, group=self.pg)
                recv.append((p, rows_all, got))
            for k in range(W):
                ops.multi_scatter_add([dict(grad=p.grad, rows=rows_all[k].contiguous(), src_rows=got[k]) for p, rows_all, got in recv])
        lr, eps = opt.param_groups[0]["lr"], opt.param_groups[0]["eps"]
        ops.multi_sparse_adagrad([dict(param=p.data, grad=p.grad, state_sum=opt.state[p]["sum"],
                                       rows=rows_all.reshape(-1).contiguous(), stamp=self._stamps[p]) for p, rows_all, _ in plan],
                                 lr, eps, self._step_id)
        ops.step_counter_bump(self._step_id)

    def step(self, global_batch):
        self.fused_step(global_batch[self.rank_id::self.world].to(self.device))
