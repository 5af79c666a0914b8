"""Fused negative-sampling training step behind the KGOptimizer contract.

``FusedKGOptimizer`` has the constructor and the methods of ``KGOptimizer`` (optim.py; reference
optimizers/kg_optimizer.py:14-316) and produces the same loss and the same parameter update, but ``epoch`` runs
each batch as ONE chain of our kernels instead of two autograd graphs of eager ops:

    negatives -> K1 (once: the positive and the negative call share their queries) -> K3 forward on the
    (B, 1+neg) tails -> chk_nsloss (loss + d/dscores) -> K3 adjoint (tail-row gradients accumulated straight into the
    dense entity gradient) -> K1 adjoint -> row scatters -> optimizer

With ``torch.optim.Adagrad`` (lr_decay = 0, weight_decay = 0) the optimizer step is row-sparse and exact
(chk_sparse_adagrad on the touched rows of every table, sharing the optimizer's own ``state['sum']`` tensors, so
``optimizer.state_dict()`` stays valid); any other optimizer gets the dense ``.grad`` and its own ``step()``.  The
whole chain has static shapes and is captured in a CUDA graph after the first batch (one graph launch per step; a
ragged last batch replays eagerly).  The loss is accumulated on the device: one host sync per epoch instead of one
per step (reference :273 ``l.item()``).  The N3 / F2 regularisers (optimizers/regularizers.py:21-58, on the positive
call's factors entity[h], rel[r], entity[t]) are evaluated in closed form inside the chain: value added to the loss,
gradient rows 3w|f|f/B (2wf/B) added to the row gradients.  Other regularisers with a non-zero weight and gradient
accumulation (update_steps > 1) fall back to the unfused contract path of the base class.
"""
import torch

from . import ops
from .optim import KGOptimizer


class FusedKGOptimizer(KGOptimizer):
    def __init__(self, *args, use_cuda_graph: bool = True, **kw):
        super().__init__(*args, **kw)
        m = self.model
        w = getattr(self.regularizer, "weight", None)
        kind = type(self.regularizer).__name__
        self._reg = None                             # (power, weight) of a closed-form regulariser with non-zero weight
        if isinstance(w, (int, float)) and w != 0 and kind in ("N3", "F2"):
            self._reg = (3 if kind == "N3" else 2, float(w))
        self.fused = (w == 0 or w == 0.0 or self._reg is not None) and self.update_steps == 1 and m.entity.weight.is_cuda
        opt = self.optimizer
        self.sparse_adagrad = (type(opt) is torch.optim.Adagrad and
                               all(g["lr_decay"] == 0 and g["weight_decay"] == 0 and not g.get("maximize", False)
                                   for g in opt.param_groups))
        self.use_cuda_graph = use_cuda_graph
        self._graph = None
        self._static_batch = None
        dev = m.entity.weight.device
        self._loss_sum = torch.zeros((), dtype=m.entity.weight.dtype, device=dev)
        self._step_id = torch.ones((), dtype=torch.int32, device=dev)
        self._stamps = {}
        self.local_batch_size = self.batch_size      # data parallel: batch_size / world
        self.grad_scale = 1.0                        # data parallel: 1/world (mean over ranks of the local mean losses)
        if self.fused:
            for p in m.parameters():                 # static dense gradient buffers, all-zero between steps
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
            if self.sparse_adagrad:
                for p in m.parameters():
                    self._stamps[p] = torch.zeros(p.shape[0], dtype=torch.int32, device=dev)

    # ------------------------------------------------------------------------------------------ one fused step
    def _forward_backward(self, batch):
        m = self.model
        r = m.rank
        ent, rel, rd, cw = m.entity.weight, m.rel.weight, m.rel_diag.weight, m.c.weight
        ctx = m._ctx_weight()
        heads, rels = batch[:, 0].contiguous(), batch[:, 1].contiguous()
        negs = self.get_neg_samples(batch)
        tails = torch.cat([batch[:, 2:3], negs], 1).contiguous()
        B, nt = tails.shape
        learn = m.bias == "learn"
        with torch.no_grad():
            q, _ = ops.query_fwd(m.KIND, r, bool(m.multi_c), ent, rel, rd, ctx, cw, heads, rels)
            bh_vals = m.bh.weight.view(-1)[heads].contiguous() if learn else None
            scores = ops.score_gather_fwd(r, B, nt, q, 1, 0, ent, tails, 0, bh_vals, 1 if learn else 0, 0,
                                          m.bt.weight.view(-1) if learn else None)
            gs = ops.nsloss(scores, self._loss_sum)
            if self.grad_scale != 1.0:
                gs.mul_(self.grad_scale)
            grad_q = ops.score_gather_bwd_scatter(r, B, nt, q, 1, 0, ent, tails, gs, ent.grad)
            g_ent, g_rel, g_rd, g_ctx, g_c = ops.query_bwd(m.KIND, r, bool(m.multi_c), ent, rel, rd, ctx, cw, heads, rels, grad_q)
            tabs = [dict(grad=ent.grad, rows=heads, src_rows=g_ent), dict(grad=rel.grad, rows=rels, src_rows=g_rel),
                    dict(grad=rd.grad, rows=rels, src_rows=g_rd)]
            if self._reg is not None:                # reg = w * sum_f sum |f|^p / B over (entity[h], rel[r], entity[t])
                power, w = self._reg
                pos = batch[:, 2].contiguous()
                fh, fr, ft = ent[heads], rel[rels], ent[pos]
                if power == 3:
                    val = (fh.abs() ** 3).sum() + (fr.abs() ** 3).sum() + (ft.abs() ** 3).sum()
                    dh, dr, dt = 3 * fh.abs() * fh, 3 * fr.abs() * fr, 3 * ft.abs() * ft
                else:
                    val = (fh ** 2).sum() + (fr ** 2).sum() + (ft ** 2).sum()
                    dh, dr, dt = 2 * fh, 2 * fr, 2 * ft
                self._loss_sum += val * (w / B)
                k = w / B * self.grad_scale
                g_ent.add_(dh, alpha=k)
                g_rel.add_(dr, alpha=k)
                tabs.append(dict(grad=ent.grad, rows=pos, src_rows=(dt * k).contiguous()))
            if ctx is not None:
                tabs.append(dict(grad=ctx.grad, rows=rels, src_rows=g_ctx))
            if m.multi_c:
                tabs.append(dict(grad=cw.grad, rows=rels, src_rows=g_c))
            else:
                cw.grad += g_c.sum()
            if learn:
                tabs.append(dict(grad=m.bh.weight.grad, rows=heads, src_rows=gs.sum(1).contiguous()))
                tabs.append(dict(grad=m.bt.weight.grad, rows=tails.view(-1), src_rows=gs))
            ops.multi_scatter_add(tabs)          # all row scatters of the step in one launch
        return heads, rels, tails

    def _sparse_step(self, heads, rels, tails):
        m, opt = self.model, self.optimizer
        lr, eps = opt.param_groups[0]["lr"], opt.param_groups[0]["eps"]
        ent_rows = torch.cat([heads, tails.view(-1)])
        zero_row = torch.zeros(1, dtype=torch.int64, device=heads.device)
        plan = [(m.entity.weight, ent_rows), (m.rel.weight, rels), (m.rel_diag.weight, rels),
                (m.c.weight, rels if m.multi_c else zero_row)]
        if m._ctx_weight() is not None:
            plan.append((m._ctx_weight(), rels))
        if m.bias == "learn":
            plan += [(m.bh.weight, heads), (m.bt.weight, tails.view(-1))]
        ops.multi_sparse_adagrad([dict(param=p.data, grad=p.grad, state_sum=opt.state[p]["sum"], rows=rows.contiguous(),
                                       stamp=self._stamps[p]) for p, rows in plan], lr, eps, self._step_id)
        ops.step_counter_bump(self._step_id)

    def _step_body(self, batch):
        heads, rels, tails = self._forward_backward(batch)
        if self.sparse_adagrad:
            self._sparse_step(heads, rels, tails)

    def fused_step(self, batch):
        """One training step on a device batch (B, 3); the loss is added to the device-side epoch accumulator."""
        full = batch.shape[0] == self.local_batch_size
        if self.use_cuda_graph and full:
            if self._graph is None:
                self._static_batch = batch.clone()
                self._step_body(self._static_batch)           # warm-up (allocator, lazy init) — a real step
                self._post_step()
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._step_body(self._static_batch)
                self._graph = g
                return                                        # the capture pass does not execute: this batch was the warm-up step
            self._static_batch.copy_(batch)
            self._graph.replay()
        else:
            self._step_body(batch)
        self._post_step()

    def _post_step(self):
        if self.sparse_adagrad:
            for p in self.model.parameters():                 # keep torch's bookkeeping in sync (unused when lr_decay = 0)
                self.optimizer.state[p]["step"] += 1
        else:
            self.optimizer.step()
            self.optimizer.zero_grad(set_to_none=False)

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
        return self._loss_sum.item() / max(n, 1)
