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
    """Data-parallel version of the fused step (train.FusedKGOptimizer): ``batch_size`` is the GLOBAL batch; every rank runs
    the fused forward / backward chain on rows ``rank::world`` of it, padded with masked rows to ``ceil(batch_size / world)``
    so that every step — a ragged last batch and a non-divisible batch included — has the same shapes on every rank and
    replays ONE CUDA graph that contains the NCCL calls.  The loss terms are normalised by the GLOBAL number of terms, so the
    sum over ranks of the local gradients is the gradient of the global mean loss.  Gradient reduction, per key group:

    * **sparse row exchange** (the 4M-entity ``entity`` / ``bh`` / ``bt`` tables): the ranks all_gather their slot ids (small,
      right after sampling; the grouping of ALL ranks' slots runs beside the local forward / backward) and their flat
      contribution buffers (gradient rows per slot, duplicates not yet merged); then ONE kernel (chk_reduce_apply) walks
      the union of the touched rows, sums each row's contributions in (rank, slot) order — the same bits on every replica —
      and applies the row-sparse Adagrad update in place.  No dense N x 2r gradient exists, nothing is scattered back.
      With NCCL the big tables are additionally **owner-sharded** (``owner_sharded``): every rank keeps a full-layout copy in
      symmetric (peer-accessible) memory, but only the rows of its own block ``[rank * rows_per_owner, ...)`` are current.  K3 reads
      every tail row from its owner's copy over NVLink (chk_score_gather_train_peer), the head rows are fetched the same way
      (chk_peer_gather_rows), and each rank reduces and updates ONLY the rows it owns — the update work per rank no longer grows
      with the number of ranks.  ``sync_replicas()`` (one in-place all_gather per table, called at the end of ``epoch()``) makes
      every copy current again for evaluation / ``state_dict()``.  The two all_gathers of a step order the accesses: a rank reads
      a peer's rows only after that peer finished the previous step's update, and updates its rows only after every peer's K3.
    * **dense all_reduce** (tables smaller than the rows that would travel: every table at FB15k-237 / WN18RR size, and the
      relation tables always): the local contributions are segment-reduced into dense gradients that are views of one flat
      buffer, ONE all_reduce sums it, chk_dense_apply runs torch.optim.Adagrad / Adam over the tables and clears it."""

    def __init__(self, *args, process_group=None, sparse_exchange=None, owner_sharded=None, peer_dense=None, peer_exchange=None, **kw):
        """sparse_exchange: None = decide by size (below); True / False force the entity-keyed tables onto the sparse row
        exchange / the dense all_reduce (tests and small-scale checks of the big-table path).  owner_sharded: None = shard the
        update of the sparse tables by owner whenever that path exists (NCCL, world > 1, pair coefficients, no N3 / F2);
        False = every replica applies every update.  peer_dense: None = under NCCL the dense tables' gradient all_reduce, the
        optimizer and the parameter broadcast run as ONE kernel over NVLink peer memory (chk_dp_fused_apply); False = NCCL
        all_reduce + chk_dense_apply on every replica.  peer_exchange: None = under NCCL the two all_gathers of the sparse exchange
        (slot ids, contribution buffers) are peer reads of symmetric buffers behind a flag barrier (chk_dp_all_gather);
        False = NCCL all_gather."""
        super().__init__(*args, **kw)
        self.pg = process_group
        self.world = _world(process_group)
        self.rank_id = dist.get_rank(process_group) if self.world > 1 else 0
        self.stream_id = self.rank_id
        self.local_batch_size = (self.batch_size + self.world - 1) // self.world
        self._n_global = self.batch_size
        if not self.fused or self.kind == "other":
            raise ValueError("FusedDataParallelKGOptimizer needs torch.optim.Adagrad (lr_decay=0, weight_decay=0) or torch.optim.Adam "
                             "(defaults), a zero / N3 / F2 regulariser and update_steps=1; use DataParallelKGOptimizer otherwise")
        m = self.model
        ent = m.entity.weight
        rows_per_step = self.local_batch_size * (2 + self.neg_sample_size) * self.world
        travelling = rows_per_step * (ent.shape[1] * ent.element_size() + 8)
        self.sparse_entity = bool(sparse_exchange) if sparse_exchange is not None else ent.numel() * ent.element_size() > travelling
        if self.sparse_entity and self.kind != "adagrad":
            raise ValueError("the sparse row exchange applies row-sparse Adagrad; Adam has dense semantics (every row moves every step)")
        sparse_names = SPARSE_TABLES if self.sparse_entity else ()
        self._dense = [p for n, p in m.named_parameters() if n.split(".")[0] not in sparse_names]
        flat = torch.zeros(sum(p.numel() for p in self._dense), dtype=ent.dtype, device=ent.device)
        o = 0
        for p in self._dense:                       # dense gradients are views of ONE flat buffer: one all_reduce sums them all
            p.grad = flat[o:o + p.numel()].view_as(p)
            o += p.numel()
        self._flat_grad = flat
        self.owner_sharded = False
        self.own = None
        self.peer_dense = False
        self.peer_exchange = False
        nccl = self.world > 1 and dist.get_backend(self.pg) == "nccl"
        if nccl and self.world <= 8 and (peer_dense is not False or peer_exchange is not False):
            self._setup_peer_signals()
            self.peer_exchange = peer_exchange is not False and self.sparse_entity
        elif peer_exchange:
            raise ValueError("peer_exchange=True needs NCCL and 2..8 ranks")
        if nccl and peer_dense is not False and self.world <= 8 and flat.numel() > 0:
            self._setup_peer_dense()
        elif peer_dense:
            raise ValueError("peer_dense=True needs NCCL and 2..8 ranks")
        can = (self.sparse_entity and self.world > 1 and self._reg is None and self._pair_coef_mode()
               and dist.get_backend(self.pg) == "nccl")
        if owner_sharded and not can:
            raise ValueError("owner_sharded=True needs NCCL, world > 1, the sparse exchange, pair coefficients and no N3 / F2 regulariser")
        if can and owner_sharded is not False:
            self._setup_owner_shards()

    # ------------------------------------------------------------------------------------------ peer memory
    def _symm_alloc(self, numel, dtype):
        """(zeroed symmetric buffer, handle, device tensor of the `world` peer base pointers); collective."""
        import torch.distributed._symmetric_memory as symm_mem
        group = self.pg if self.pg is not None else dist.group.WORLD
        dev = self.model.entity.weight.device
        t = symm_mem.empty((numel,), dtype=dtype, device=dev)
        hdl = symm_mem.rendezvous(t, group)
        t.zero_()
        return t, hdl, torch.tensor(list(hdl.buffer_ptrs), dtype=torch.int64, device=dev)

    def _setup_peer_signals(self):
        """Flag arrays of the peer-memory kernels (chk_dp_fused_apply, chk_dp_all_gather): symmetric int32[4 * world] + local state."""
        self._peer_sig = self._symm_alloc(4 * self.world, torch.int32)
        self._peer_local = torch.zeros(8, dtype=torch.int32, device=self.model.entity.weight.device)
        torch.cuda.synchronize()
        dist.barrier(group=self.pg)

    def _exchange_buffer(self, pl, name, numel, dtype, device):
        if not self.peer_exchange:
            return super()._exchange_buffer(pl, name, numel, dtype, device)
        t, hdl, ptrs = self._symm_alloc(numel, dtype)          # collective: every rank builds the same plans in the same order
        setattr(pl, "_x_" + name, (t, hdl, ptrs))
        return t

    def _all_gather(self, pl, name, dst, src, channel):
        """dst[k] = rank k's `src`; also the ordering point of the step (see the class docstring)."""
        if self.peer_exchange:
            ops.dp_all_gather(self.world, self.rank_id, getattr(pl, "_x_" + name)[2], src.numel() * src.element_size(), dst,
                              self._peer_sig[2], channel, self._peer_local)
        else:
            dist.all_gather_into_tensor(dst.view(-1), src, group=self.pg)

    def _setup_peer_dense(self):
        """Flat gradient / parameter / optimizer-state buffers of the dense tables in symmetric memory (same layout on every
        rank), a symmetric signal array, and the device arrays of peer base pointers chk_dp_fused_apply takes."""
        q = 4 * self.world
        n = (self._flat_grad.numel() + q - 1) // q * q          # whole 4-element vectors per rank slice (the tail is padding)
        dev, dt = self._flat_grad.device, self._flat_grad.dtype
        keys = ("sum", None) if self.kind == "adagrad" else ("exp_avg", "exp_avg_sq")
        bufs = {"grad": self._symm_alloc(n, dt), "param": self._symm_alloc(n, dt), "s0": self._symm_alloc(n, dt)}
        if keys[1] is not None:
            bufs["s1"] = self._symm_alloc(n, dt)
        bufs["sig"] = self._peer_sig
        o = 0
        for p in self._dense:                         # rebind gradients, parameters and optimizer state to views of the flat buffers
            k = p.numel()
            st = self.optimizer.state[p]
            bufs["param"][0][o:o + k].view_as(p).copy_(p.data)
            p.data = bufs["param"][0][o:o + k].view_as(p)
            p.grad = bufs["grad"][0][o:o + k].view_as(p)
            bufs["s0"][0][o:o + k].view_as(p).copy_(st[keys[0]])
            st[keys[0]] = bufs["s0"][0][o:o + k].view_as(p)
            if keys[1] is not None:
                bufs["s1"][0][o:o + k].view_as(p).copy_(st[keys[1]])
                st[keys[1]] = bufs["s1"][0][o:o + k].view_as(p)
            o += k
        self._flat_grad = bufs["grad"][0]
        self._peer = bufs
        torch.cuda.synchronize()
        dist.barrier(group=self.pg)
        self.peer_dense = True
        self.model.parameters_changed()

    def check_peer_status(self):
        """Raises if a peer did not arrive at a chk_dp_fused_apply step within its timeout (host sync; tests and epoch ends)."""
        if (self.peer_dense or self.peer_exchange) and int(self._peer_local[2].item()) != 0:
            raise RuntimeError("a rank did not arrive at a peer-memory step (timeout); the replicas are out of step")

    def _dense_step(self):
        """Sum of the dense gradients over the ranks + optimizer + identical new values on every replica."""
        m, st = self.model, self.optimizer.state
        if self.peer_dense:
            b = self._peer
            ops.dp_fused_apply(m.entity.weight, ops.CHK_OPT_ADAM if self.kind == "adam" else ops.CHK_OPT_ADAGRAD, self.world, self.rank_id,
                               b["grad"][2], b["param"][2], b["s0"][2], b["s1"][2] if "s1" in b else None, b["sig"][2],
                               self._flat_grad.numel(), self._hyper, self._step_id, self._peer_local)
            return
        if self.world > 1:
            dist.all_reduce(self._flat_grad, op=dist.ReduceOp.SUM, group=self.pg)
        if self.kind == "adam":
            tabs = [(p.data, p.grad, st[p]["exp_avg"], st[p]["exp_avg_sq"]) for p in self._dense]
        else:
            tabs = [(p.data, p.grad, st[p]["sum"], None) for p in self._dense]
        ops.dense_apply(ops.CHK_OPT_ADAM if self.kind == "adam" else ops.CHK_OPT_ADAGRAD, tabs, self._hyper, self._step_id)

    # ------------------------------------------------------------------------------------------ owner-sharded tables
    def _sparse_params(self):
        m = self.model
        return [("entity", m.entity.weight), ("bh", m.bh.weight), ("bt", m.bt.weight)]

    def _setup_owner_shards(self):
        """Moves entity / bh / bt into symmetric memory (padded to world * rows_per_owner rows so that the replica sync is an
        in-place all_gather), pads their Adagrad state likewise, and builds the device arrays of peer base pointers."""
        import torch.distributed._symmetric_memory as symm_mem
        m, W, r = self.model, self.world, self.rank_id
        N = m.sizes[0]
        group = self.pg if self.pg is not None else dist.group.WORLD
        rpo = (N + W - 1) // W
        self.rows_per_owner = rpo
        self.own = (min(r * rpo, N), min((r + 1) * rpo, N))
        self._symm, self._state_pad = {}, {}
        dev = m.entity.weight.device
        for name, p in self._sparse_params():
            buf = symm_mem.empty((rpo * W, p.shape[1]), dtype=p.dtype, device=dev)
            hdl = symm_mem.rendezvous(buf, group)
            buf.zero_()
            buf[:N].copy_(p.data)
            p.data = buf[:N]
            self._symm[name] = (buf, hdl, torch.tensor(list(hdl.buffer_ptrs), dtype=torch.int64, device=dev))
            st = self.optimizer.state[p]
            pad = torch.zeros((rpo * W, p.shape[1]), dtype=st["sum"].dtype, device=dev)
            pad[:N].copy_(st["sum"])
            st["sum"] = pad[:N]
            self._state_pad[name] = pad
        torch.cuda.synchronize()
        dist.barrier(group=self.pg)
        self.owner_sharded = True
        self.model.parameters_changed()

    def sync_replicas(self):
        """Owner-sharded tables: every rank receives the current rows (and Adagrad state) of the other owners.  Needed before
        the parameters are read outside the training step (evaluation, state_dict, checkpoint); epoch() calls it."""
        if not self.owner_sharded:
            return
        r, rpo = self.rank_id, self.rows_per_owner
        for name, _ in self._sparse_params():
            for t in (self._symm[name][0], self._state_pad[name]):
                dist.all_gather_into_tensor(t.view(-1), t[r * rpo:(r + 1) * rpo].reshape(-1), group=self.pg)
        self.model._replicas_stale = False
        self.model.parameters_changed()

    def _post_step(self):
        super()._post_step()
        if self.owner_sharded:
            self.model._replicas_stale = True           # get_ranking refuses to run on a copy that is not current

    # ------------------------------------------------------------------------------------------ plan
    def _build_groups(self, pl):
        m = self.model
        dev = ent_dev = m.entity.weight.device
        N, R2, W = m.sizes[0], m.rel.weight.shape[0], self.world
        B, Bq, S_e = pl.B, pl.Bq, pl.S_e
        pl.w_rel = ops.group_workspace(R2, B, dev)
        dense_col = lambda p, *src: dict(param=p.data, state0=None, dense=p.grad, src=list(src))
        if self.sparse_entity:
            L = pl.flat.numel()
            pl.all_ids = torch.zeros((W, S_e), dtype=torch.int64, device=ent_dev)
            pl.all_flat = torch.zeros((W, L), dtype=pl.flat.dtype, device=ent_dev)
            pl.w_ent = ops.group_workspace(N, W * S_e, dev)
            pl.ent_group_ids, pl.ent_group_slots = pl.all_ids.view(-1), W * S_e
            f0, off = pl.all_flat[0], pl.offsets
            sp_col = lambda p, *src: dict(param=p.data, state0=self._state_of(p), dense=None, src=list(src))
            esrc, epair = self._entity_sources(pl, (f0[off["g_ent"]:], 0, Bq, L), rank_stride=L, base=f0)
            ecols = [dict(sp_col(m.entity.weight, *esrc), pair=epair)]
            if m.bias == "learn":
                ecols.append(sp_col(m.bh.weight, (f0[off["gs" if pl.dn else "g_bh"]:], 0, Bq, L)))
                ecols.append(sp_col(m.bt.weight, (f0[off["gs"]:], Bq, S_e, L)))
            groups = [dict(ids=pl.ent_group_ids, n_keys=N, slots_per_rank=S_e, world=W, work=pl.w_ent, cols=ecols)]
            if self.owner_sharded:                       # local staging of the head rows / head biases fetched from their owners
                pl.stage_ent = torch.zeros((Bq, m.entity.weight.shape[1]), dtype=pl.flat.dtype, device=ent_dev)
                pl.stage_bh = torch.zeros((Bq,), dtype=pl.flat.dtype, device=ent_dev)
                pl.arange_q = torch.arange(Bq, dtype=torch.int64, device=ent_dev)
        else:
            pl.w_ent = ops.group_workspace(N, S_e, dev)
            pl.ent_group_ids, pl.ent_group_slots = pl.ent_ids, S_e
            esrc, epair = self._entity_sources(pl, (pl.g_ent, 0, Bq, 0))
            ecols = [dict(dense_col(m.entity.weight, *esrc), pair=epair)]
            if m.bias == "learn":
                ecols.append(dense_col(m.bh.weight, (pl.gs if pl.dn else pl.g_bh, 0, Bq, 0)))
                ecols.append(dense_col(m.bt.weight, (pl.gs, Bq, S_e, 0)))
            groups = [dict(ids=pl.ent_ids, n_keys=N, slots_per_rank=S_e, world=1, work=pl.w_ent, cols=ecols)]
        groups += self._relation_groups(pl, dense_col)
        pl.groups = groups
        pl.red = ops._red_groups(groups)
        if self.sparse_entity:                           # two launches: the relation tables' all_reduce runs beside the entity reduce
            pl.red_ent, pl.red_rel = ops._red_groups(groups[:1]), ops._red_groups(groups[1:])
        pl.works = [pl.w_ent, pl.w_rel]

    def _graph_batch(self):
        return self.local_batch_size

    def _global_rows(self, n_valid_local):
        return self._n_global

    # ------------------------------------------------------------------------------------------ step
    def _after_prep(self, pl):
        if self.sparse_entity and self.world > 1:           # everyone's slot ids, so the union can be grouped during the local pass
            self._all_gather(pl, "ids", pl.all_ids, pl.ent_ids, 0)
        elif self.sparse_entity:
            pl.all_ids.view(-1).copy_(pl.ent_ids)
        cur = torch.cuda.current_stream()
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):                 # owner-sharded: only the slots naming this rank's rows are grouped
            ops.group_build(pl.rels_b, self.model.rel.weight.shape[0], pl.w_rel)
            ops.group_build(pl.ent_group_ids, self.model.sizes[0], pl.w_ent, own=self.own)
            self._ev_grouped = torch.cuda.Event()
            self._ev_grouped.record(self._side)

    def _head_tables(self, pl, learn):
        if not self.owner_sharded:
            return super()._head_tables(pl, learn)
        rpo = self.rows_per_owner
        ops.peer_gather_rows(self._symm["entity"][2], rpo, pl.heads, pl.stage_ent.shape[1], pl.stage_ent)
        if learn:
            ops.peer_gather_rows(self._symm["bh"][2], rpo, pl.heads, 1, pl.stage_bh)
        return pl.stage_ent, pl.arange_q, (pl.stage_bh if learn else None)

    def _score_train(self, pl, head_ix, bh_tab, learn):
        if not self.owner_sharded:
            return super()._score_train(pl, head_ix, bh_tab, learn)
        m = self.model
        qsb, qsj = (pl.nt, 1) if pl.dn else (1, 0)
        ops.score_gather_train_peer(m.rank, pl.B, pl.nt, pl.q, qsb, qsj, self._symm["entity"][2], self._symm["bt"][2] if learn else None,
                                    self.rows_per_owner, pl.tails, head_ix, qsb, qsj, bh_tab, self._hyper, pl.loss_part, pl.gs,
                                    pl.grad_q, pl.grow, pl.g_bh if learn else None, pair_coef=pl.coef)

    def _apply(self, pl):
        m = self.model
        W = self.world
        cur = torch.cuda.current_stream()
        if self.sparse_entity and W > 1:
            # side stream (behind the grouping): relation tables — local row sums into the dense gradients, then their cross-rank
            # step — WHILE the main stream gathers every rank's contribution buffer, sums the entity-keyed rows of all ranks in
            # (rank, slot) order and applies Adagrad in place
            self._side.wait_stream(cur)                                     # the K1 adjoint's relation-row gradients
            with torch.cuda.stream(self._side):
                ops.reduce_apply(m.entity.weight, ops.CHK_OPT_NONE, pl.red_rel, self._hyper)
                self._dense_step()
            self._all_gather(pl, "flat", pl.all_flat, pl.flat, 1)
            cur.wait_event(self._ev_grouped)                                # the union of the entity slots is grouped
            ops.reduce_apply(m.entity.weight, ops.CHK_OPT_ADAGRAD, pl.red_ent, self._hyper)
            cur.wait_stream(self._side)
        else:
            if self.sparse_entity:
                pl.all_flat.view(-1).copy_(pl.flat)
            cur.wait_stream(self._side)
            # one launch: the local row sums written into the dense gradients (and, single rank, the sparse tables updated in place)
            ops.reduce_apply(m.entity.weight, ops.CHK_OPT_ADAGRAD if self.sparse_entity else ops.CHK_OPT_NONE, pl.red, self._hyper)
            self._dense_step()
        ops.step_finish(m.entity.weight, pl.works, pl.loss_part, self._loss_sum, self._step_id)

    def step(self, global_batch):
        """One data-parallel step on a GLOBAL batch (same tensor on every rank, any number of rows <= batch_size)."""
        n_global = global_batch.shape[0]
        local = global_batch[self.rank_id::self.world]
        n_valid = local.shape[0]
        Bl = self.local_batch_size
        if n_valid < Bl:                                    # masked padding rows: identical shapes on every rank, every step
            pad = torch.zeros((Bl, 3), dtype=global_batch.dtype, device=local.device)
            pad[:n_valid] = local
            local = pad
        self._n_global = n_global
        self.fused_step(local.to(self.device, non_blocking=True), n_valid=n_valid)

    def epoch(self, examples):
        """KGOptimizer.epoch (reference optimizers/kg_optimizer.py:239-277) data parallel: rank 0's permutation is broadcast,
        every global batch (the ragged last one too) goes through ``step``; returns the mean over the batches of the GLOBAL
        mean loss (the same number on every rank)."""
        n = examples.shape[0]
        perm = torch.randperm(n).to(self.device)
        if self.world > 1:
            dist.broadcast(perm, src=dist.get_global_rank(self.pg, 0) if self.pg is not None else 0, group=self.pg)
        actual = examples.to(self.device)[perm]
        self._loss_sum.zero_()
        nb = 0
        for b0 in range(0, n, self.batch_size):
            self.step(actual[b0:b0 + self.batch_size])
            nb += 1
        total = self._loss_sum.double().clone()
        if self.world > 1:
            dist.all_reduce(total, op=dist.ReduceOp.SUM, group=self.pg)
        self.sync_optimizer_state()
        self.sync_replicas()
        self.check_peer_status()
        return total.item() / max(nb, 1)
