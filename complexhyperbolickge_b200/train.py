"""Fused negative-sampling training step behind the KGOptimizer contract.

``FusedKGOptimizer`` has the constructor and the methods of ``KGOptimizer`` (optim.py; reference
optimizers/kg_optimizer.py:14-316) and produces the same loss and the same parameter update, but ``epoch`` runs each
batch as ONE short chain of our kernels instead of two autograd graphs of eager ops:

    chk_train_prep (device sampler, id arrays) -> K1 (once: the positive and the negative call share their queries)
    -> chk_score_gather_train (K3 forward on the (B, 1+neg) tails + loss terms + adjoint, tail rows gathered once)
    -> K1 adjoint -> chk_reduce_apply (segment-reduce of every touched row's gradient contributions in slot order + the
    optimizer update of that row) -> chk_step_finish          [chk_group_build runs beside K1/K3 on a second stream]

No torch op, no floating-point atomic and no dense N x 2r gradient is on that chain: duplicate rows are segment-reduced in
a fixed order, so a step is bit-reproducible.  Optimizers:

* ``torch.optim.Adagrad`` (lr_decay = 0, weight_decay = 0): row-sparse and exact, applied in place by chk_reduce_apply on
  the optimizer's own ``state['sum']`` tensors (a zero-gradient row is a no-op in Adagrad, so sparse == dense);
* ``torch.optim.Adam`` (no weight decay, no amsgrad): dense semantics (every row moves every step) — the row sums
  are written into dense gradients and chk_dense_apply updates whole tables with torch's arithmetic on the optimizer's own
  ``exp_avg`` / ``exp_avg_sq`` tensors;
* any other optimizer gets the dense ``.grad`` and its own ``step()``.

The chain has static shapes and is captured in a CUDA graph after the first batch (one graph launch per step; a ragged last
batch runs the same kernels eagerly on its own buffers).  The loss is accumulated on the device: one host sync per epoch
instead of one per step (reference :273 ``l.item()``).  ``double_neg`` (reference :46, :78-91 — commented out at HEAD) is
restored with the original's semantics: every negative corrupts the head as well as the tail, so the step runs K1 on
B(1+neg) per-pair queries.  N3 / F2 (optimizers/regularizers.py:21-58, on the positive call's factors) are evaluated by
chk_reg_factors inside the chain.  Other regularisers with a non-zero weight and gradient accumulation (update_steps > 1)
fall back to the unfused contract path of the base class.
"""
import torch

from . import ops
from .optim import KGOptimizer


def _round4(n: int) -> int:
    return (n + 3) // 4 * 4


class _Plan:
    """Static buffers and reduce descriptors of the step for one local batch size."""

    def __init__(self, o: "FusedKGOptimizer", B: int):
        m = o.model
        dev, dt = m.entity.weight.device, m.entity.weight.dtype
        r, n = m.rank, m.dim
        self.B, self.nt = B, o.neg_sample_size + 1
        nt = self.nt
        self.dn = bool(o.double_neg)
        self.Bq = Bq = B * nt if self.dn else B           # queries of the step
        self.P = P = B * nt                               # (query, tail) pairs
        i64 = dict(dtype=torch.int64, device=dev)
        self.batch = torch.zeros((B, 3), **i64)
        self.ent_ids = o._exchange_buffer(self, "ids", Bq + P, torch.int64, dev)      # slot -> entity id: [heads | tails], the entity group's key array
        self.heads, self.tails = self.ent_ids[:Bq], self.ent_ids[Bq:]
        self.rels = torch.zeros((Bq,), **i64)
        self.rels_b = self.rels if not self.dn else torch.zeros((B,), **i64)
        att = m._ctx_weight() is not None
        wrd = m.rel_diag.weight.shape[1]
        # contributions of the entity group live in ONE flat buffer (it is what travels in the data-parallel exchange):
        #   stored rows:        [ g_ent Bq x 2r | grow P x 2r          | gs P | g_bh B ]
        #   pair coefficients:  [ g_ent Bq x 2r | q Bq x 2r | coef P x 4 | gs P | g_bh B ]   (tail-row gradients rebuilt by the reduce)
        self.coef_mode = o._pair_coef_mode()
        mk = lambda *s: torch.zeros(s, dtype=dt, device=dev)
        o_ent, o_row = 0, _round4(Bq * 2 * r)
        if self.coef_mode:
            o_coef = o_row + _round4(Bq * 2 * r)
            o_gs = o_coef + P * 4
        else:
            o_coef = None
            o_gs = o_row + _round4(P * 2 * r)
        o_bh = o_gs + _round4(P)
        flat_len = o_bh + (0 if self.dn else _round4(B))
        self.flat = o._exchange_buffer(self, "flat", flat_len, dt, dev)
        self.g_ent = self.flat[o_ent:o_ent + Bq * 2 * r].view(Bq, 2 * r)
        if self.coef_mode:
            self.q = self.flat[o_row:o_row + Bq * 2 * r].view(Bq, 2 * r)
            self.coef = self.flat[o_coef:o_coef + P * 4].view(P, 4)
            self.grow = None
        else:
            self.q, self.coef = mk(Bq, 2 * r), None
            self.grow = self.flat[o_row:o_row + P * 2 * r].view(P, 2 * r)
        self.gs = self.flat[o_gs:o_gs + P].view(B, nt)
        self.g_bh = None if self.dn else self.flat[o_bh:o_bh + B]
        self.c_out, self.grad_q = mk(Bq), mk(Bq, 2 * r)
        self.g_rel, self.g_rd, self.g_c = mk(Bq, 2 * n), mk(Bq, wrd), mk(Bq)
        self.g_ctx = mk(Bq, n) if att else None
        if self.dn:                                       # per-pair relation-row gradients are summed over j first
            self.s_rel, self.s_rd, self.s_c = mk(B, 2 * n), mk(B, wrd), mk(B)
            self.s_ctx = mk(B, n) if att else None
        else:
            self.s_rel, self.s_rd, self.s_c, self.s_ctx = self.g_rel, self.g_rd, self.g_c, self.g_ctx
        self.loss_part = mk(B)
        self.inj_t = self.inj_h = None
        self.S_e = Bq + P                                 # entity-group slots per rank
        self.graph = None
        self.offsets = dict(g_ent=o_ent, grow=o_row, q=o_row, coef=o_coef, gs=o_gs, g_bh=o_bh)      # element offsets inside ``flat``
        o._build_groups(self)


class FusedKGOptimizer(KGOptimizer):
    PAIR_COEF_RANKS = (9, 17, 33, 65, 129, 257)      # ranks chk_reduce_apply has a computed-source instantiation for

    def __init__(self, *args, use_cuda_graph: bool = True, seed=None, pair_coef=None, **kw):
        """pair_coef: None = rebuild the tail-row gradients in the reduce from three scalars per pair whenever that path
        exists (rank in PAIR_COEF_RANKS, no N3 / F2 term on the tail rows); False = always store the rows (the two are
        bit-identical; tests compare them)."""
        super().__init__(*args, **kw)
        self._pair_coef = pair_coef
        m = self.model
        w = getattr(self.regularizer, "weight", None)
        kind = type(self.regularizer).__name__
        self._reg = None                             # (power, weight) of a closed-form regulariser with non-zero weight
        if isinstance(w, (int, float)) and w != 0 and kind in ("N3", "F2"):
            self._reg = (3 if kind == "N3" else 2, float(w))
        self.fused = (w == 0 or w == 0.0 or self._reg is not None) and self.update_steps == 1 and m.entity.weight.is_cuda
        opt = self.optimizer
        if len(opt.param_groups) != 1:               # one set of hyper-parameters is baked into the device scalars
            self.fused = False
        g0 = opt.param_groups[0]
        if type(opt) is torch.optim.Adagrad and g0["lr_decay"] == 0 and g0["weight_decay"] == 0 and not g0.get("maximize", False) \
                and g0.get("initial_accumulator_value", 0) == 0:
            self.kind = "adagrad"
        elif type(opt) is torch.optim.Adam and g0["weight_decay"] == 0 and not g0["amsgrad"] and not g0.get("maximize", False) \
                and not g0.get("capturable", False) and not g0.get("fused", False):
            self.kind = "adam"
        else:
            self.kind = "other"
        self.sparse_adagrad = self.kind == "adagrad"
        self.use_cuda_graph = use_cuda_graph
        self.seed = int(torch.initial_seed() if seed is None else seed)
        self.stream_id = 0                           # data parallel: the rank (distinct negative streams per rank)
        self.world = 1
        self._plans = {}
        self._steps_done = 0
        if not self.fused:
            return
        dev, dt = m.entity.weight.device, m.entity.weight.dtype
        self._loss_sum = torch.zeros((), dtype=dt, device=dev)
        self._hyper = torch.zeros(ops.CHK_HYPER_LEN, dtype=torch.float64, device=dev)
        self._hyper_key = None
        self._side = torch.cuda.Stream(device=dev)
        start = 0
        if self.kind == "adam":
            for p in m.parameters():                 # torch.optim.Adam creates its state lazily: create it the way it would
                st = opt.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                start = max(start, int(st["step"].item()))
        elif self.kind == "adagrad":
            start = max(int(opt.state[p]["step"].item()) for p in m.parameters())
        self._steps_done = start
        self._step_id = torch.full((), start + 1, dtype=torch.int32, device=dev)      # 1-based number of the NEXT step
        if self.kind != "adagrad":
            for p in m.parameters():                 # dense gradients, all-zero between steps
                if p.grad is None:
                    p.grad = torch.zeros_like(p)

    # ------------------------------------------------------------------------------------------ pieces
    def _state_of(self, p):
        return self.optimizer.state[p]["sum"]

    def _pair_coef_mode(self) -> bool:
        ok = self.model.rank in self.PAIR_COEF_RANKS and self._reg is None
        if self._pair_coef and not ok:
            raise ValueError("pair_coef=True needs a rank in %s and no N3 / F2 regulariser" % (self.PAIR_COEF_RANKS,))
        return ok if self._pair_coef is None else bool(self._pair_coef)

    def _entity_sources(self, pl, head_src, rank_stride=0, base=None):
        """(src list, pair descriptor) of the entity column: head-entity gradient rows + tail contributions (stored rows or
        query rows + pair coefficients).  base: a flat tensor holding rank 0's buffer (data parallel) or None (plan views)."""
        Bq, S_e, off = pl.Bq, pl.S_e, pl.offsets
        view = (lambda name, t: t) if base is None else (lambda name, t: base[off[name]:])
        if pl.coef_mode:
            return ([head_src, (view("q", pl.q), Bq, S_e, rank_stride)], (view("coef", pl.coef), 0 if pl.dn else pl.nt, rank_stride))
        return [head_src, (view("grow", pl.grow), Bq, S_e, rank_stride)], None

    def _exchange_buffer(self, pl, name, numel, dtype, device):
        """Buffers of a plan that travel in the data-parallel exchange (slot ids, contribution buffer); plain memory here."""
        return torch.zeros((numel,), dtype=dtype, device=device)

    def _injected_sampler(self):
        return type(self).get_neg_samples is not KGOptimizer.get_neg_samples

    def _global_rows(self, B_local_valid: int) -> int:
        return B_local_valid                          # data parallel: the valid rows of ALL ranks

    def _set_hyper(self, n_valid_local: int, n_rows_global: int):
        g = self.optimizer.param_groups[0]
        b1, b2 = g.get("betas", (0.0, 0.0))
        eps = g.get("eps", 0.0)
        key = (g["lr"], eps, n_valid_local, n_rows_global, b1, b2)
        if key == self._hyper_key:
            return
        h = torch.tensor([g["lr"], eps, 1.0 / (n_rows_global * (self.neg_sample_size + 1)), float(n_valid_local), b1, b2,
                          1.0 / n_rows_global, 0.0], dtype=torch.float64)
        self._hyper.copy_(h)                          # stream-ordered after the earlier steps, which read the old values
        self._hyper_key = key

    def _build_groups(self, pl: _Plan):
        """Grouping workspaces + reduce descriptors (single process: every slot is local)."""
        m = self.model
        dev = m.entity.weight.device
        N, R2 = m.sizes[0], m.rel.weight.shape[0]
        B, Bq, S_e = pl.B, pl.Bq, pl.S_e
        pl.w_ent = ops.group_workspace(N, S_e, dev)
        pl.w_rel = ops.group_workspace(R2, B, dev)
        pl.ent_group_ids, pl.ent_group_slots = pl.ent_ids, S_e
        inplace = self.kind == "adagrad"

        def col(p, *src):
            return dict(param=p.data, state0=self._state_of(p) if inplace else None, dense=None if inplace else p.grad, src=list(src))
        esrc, epair = self._entity_sources(pl, (pl.g_ent, 0, Bq, 0))
        ecols = [dict(col(m.entity.weight, *esrc), pair=epair)]
        if m.bias == "learn":
            ecols.append(col(m.bh.weight, (pl.gs if pl.dn else pl.g_bh, 0, Bq, 0)))
            ecols.append(col(m.bt.weight, (pl.gs, Bq, S_e, 0)))
        groups = [dict(ids=pl.ent_ids, n_keys=N, slots_per_rank=S_e, world=1, work=pl.w_ent, cols=ecols)]
        groups += self._relation_groups(pl, col)
        pl.groups = groups
        pl.red = ops._red_groups(groups)
        pl.works = [pl.w_ent, pl.w_rel]

    def _relation_groups(self, pl: _Plan, col):
        """Relation-keyed tables (rel, rel_diag, context_vec, c): B slots, one per triple."""
        m = self.model
        B, R2 = pl.B, m.rel.weight.shape[0]
        rcols = [col(m.rel.weight, (pl.s_rel, 0, B, 0)), col(m.rel_diag.weight, (pl.s_rd, 0, B, 0))]
        if pl.s_ctx is not None:
            rcols.append(col(m._ctx_weight(), (pl.s_ctx, 0, B, 0)))
        if m.multi_c:
            rcols.append(col(m.c.weight, (pl.s_c, 0, B, 0)))
            return [dict(ids=pl.rels_b, n_keys=R2, slots_per_rank=B, world=1, work=pl.w_rel, cols=rcols)]
        return [dict(ids=pl.rels_b, n_keys=R2, slots_per_rank=B, world=1, work=pl.w_rel, cols=rcols),
                dict(ids=None, slots_per_rank=B, world=1, work=None, single_row=True, cols=[col(m.c.weight, (pl.s_c, 0, B, 0))])]

    def _plan(self, B):
        pl = self._plans.get(B)
        if pl is None:
            pl = self._plans[B] = _Plan(self, B)
        return pl

    # ------------------------------------------------------------------------------------------ one fused step
    def _forward_backward(self, pl: _Plan):
        """prep -> K1 -> K3 training pass -> K1 adjoint [-> row sums over j, regulariser]; grouping on the side stream."""
        m = self.model
        r = m.rank
        ent, rel, rd, cw = m.entity.weight.data, m.rel.weight.data, m.rel_diag.weight.data, m.c.weight.data
        ctx = m._ctx_weight()
        ctx = None if ctx is None else ctx.data
        B, nt, Bq = pl.B, pl.nt, pl.Bq
        learn = m.bias == "learn"
        inj_t = inj_h = None
        if self._injected_sampler():                  # an overridden get_neg_samples is honoured (tests inject negatives)
            inj_t = self.get_neg_samples(pl.batch).contiguous()
            if pl.dn:
                inj_h = self.get_neg_heads(pl.batch).contiguous()
        ops.train_prep(pl.batch, self.neg_sample_size, self.n_entities, pl.dn, self.seed, self._step_id, self.stream_id,
                       pl.heads, pl.rels, pl.tails, inj_t, inj_h)
        if pl.dn:
            pl.rels_b.copy_(pl.batch[:, 1])
        self._after_prep(pl)
        ent_h, head_ix, bh_tab = self._head_tables(pl, learn)           # where K1 / K3 find the head rows and head biases
        ops.query_fwd(m.KIND, r, bool(m.multi_c), ent_h, rel, rd, ctx, cw, head_ix, pl.rels, out=(pl.q, pl.c_out))
        self._score_train(pl, head_ix if learn else None, bh_tab, learn)
        ops.query_bwd_into(m.KIND, r, bool(m.multi_c), ent_h, rel, rd, ctx, cw, head_ix, pl.rels, pl.grad_q, pl.g_ent, pl.g_rel,
                           pl.g_rd, pl.g_ctx, pl.g_c)
        if pl.dn:
            ops.rowsum_groups(pl.g_rel, B, nt, pl.g_rel.shape[1], pl.s_rel)
            ops.rowsum_groups(pl.g_rd, B, nt, pl.g_rd.shape[1], pl.s_rd)
            ops.rowsum_groups(pl.g_c, B, nt, 1, pl.s_c)
            if pl.g_ctx is not None:
                ops.rowsum_groups(pl.g_ctx, B, nt, pl.g_ctx.shape[1], pl.s_ctx)
        if self._reg is not None:
            power, w = self._reg
            hs = nt if pl.dn else 1
            ops.reg_factors(power, w, self._hyper, B, ent, rel, pl.heads, hs, pl.rels, pl.tails, nt, pl.g_ent, hs * 2 * r,
                            pl.s_rel, pl.s_rel.shape[1], pl.grow, nt * 2 * r, pl.loss_part)

    def _head_tables(self, pl, learn):
        """(entity table, head index, bh table) through which K1, its adjoint and K3 read the head rows / head biases."""
        m = self.model
        return m.entity.weight.data, pl.heads, (m.bh.weight.data.view(-1) if learn else None)

    def _score_train(self, pl, head_ix, bh_tab, learn):
        """K3 training pass on the step's (B, 1+neg) tails."""
        m = self.model
        qsb, qsj = (pl.nt, 1) if pl.dn else (1, 0)
        ops.score_gather_train(m.rank, pl.B, pl.nt, pl.q, qsb, qsj, m.entity.weight.data, pl.tails, head_ix, qsb, qsj, bh_tab,
                               m.bt.weight.data.view(-1) if learn else None, self._hyper, pl.loss_part, pl.gs, pl.grad_q, pl.grow,
                               pl.g_bh if learn else None, pair_coef=pl.coef)

    def _after_prep(self, pl):
        """Ids are known: group the slots by row beside the forward / backward kernels (second stream)."""
        cur = torch.cuda.current_stream()
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            ops.group_build(pl.ent_group_ids, self.model.sizes[0], pl.w_ent)
            ops.group_build(pl.rels_b, self.model.rel.weight.shape[0], pl.w_rel)

    def _apply(self, pl):
        """Segment-reduce + optimizer."""
        m = self.model
        torch.cuda.current_stream().wait_stream(self._side)
        if self.kind == "adam":                       # Adam reads the step number after the reduce: the step ends in its own launch
            ops.reduce_apply(m.entity.weight, ops.CHK_OPT_NONE, pl.red, self._hyper)
            st = self.optimizer.state
            ops.dense_apply(ops.CHK_OPT_ADAM, [(p.data, p.grad, st[p]["exp_avg"], st[p]["exp_avg_sq"]) for p in m.parameters()],
                            self._hyper, self._step_id)
            ops.step_finish(m.entity.weight, pl.works, pl.loss_part, self._loss_sum, self._step_id)
        else:                                         # the reduce kernel's last block ends the step
            ops.reduce_apply(m.entity.weight, ops.CHK_OPT_ADAGRAD if self.kind == "adagrad" else ops.CHK_OPT_NONE, pl.red, self._hyper,
                             finish=(pl.loss_part, self._loss_sum, self._step_id))

    def _step_body(self, pl):
        self._forward_backward(pl)
        self._apply(pl)

    def fused_step(self, batch, n_valid=None):
        """One training step on a batch (B, 3) — a device tensor or a pinned host tensor (one asynchronous copy into the step's
        id buffer); the loss is added to the device-side epoch accumulator.  n_valid < B marks the trailing rows as padding
        (zero loss, zero gradient)."""
        B = batch.shape[0]
        if B == 0:
            return
        n_valid = B if n_valid is None else n_valid
        self._set_hyper(n_valid, self._global_rows(n_valid))
        pl = self._plan(B)
        graph_ok = self.use_cuda_graph and self.kind != "other" and B == self._graph_batch()
        pl.batch.copy_(batch, non_blocking=True)
        if graph_ok:
            if pl.graph is None:
                self._step_body(pl)                               # warm-up (lazy init, allocator) — a real step
                self._post_step()
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._step_body(pl)
                pl.graph = g
                return                                            # the capture pass does not execute: this batch was the warm-up step
            pl.graph.replay()
        else:
            self._step_body(pl)
        self._post_step()

    def _graph_batch(self):
        return self.batch_size

    def _post_step(self):
        self._steps_done += 1
        self.model.parameters_changed()                           # cached evaluation state (norms, shadow) is stale now
        if self.kind == "other":
            self.optimizer.step()
            self.optimizer.zero_grad(set_to_none=False)

    def sync_optimizer_state(self):
        """torch's per-parameter step counters (bookkeeping only) follow the fused steps; called once per epoch."""
        if self.kind in ("adagrad", "adam"):
            for p in self.model.parameters():
                self.optimizer.state[p]["step"].fill_(float(self._steps_done))

    # ------------------------------------------------------------------------------------------ contract
    def epoch(self, examples):
        if not self.fused:
            return super().epoch(examples)
        actual = examples[torch.randperm(examples.shape[0]), :].to(self.device)
        self._loss_sum.zero_()
        n = 0
        for b0 in range(0, examples.shape[0], self.batch_size):
            self.fused_step(actual[b0:b0 + self.batch_size])
            n += 1
        self.sync_optimizer_state()
        return self._loss_sum.item() / max(n, 1)
