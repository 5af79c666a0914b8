"""Host wrappers + autograd Functions over the C ABI (include/chk_b200.h).

PyTorch is plumbing here: it owns device memory and the CUDA stream, and autograd is only used to hand
the kernels' analytic gradients to ``torch.optim`` so that the reference's training loop
(optimizers/kg_optimizer.py:239-277) runs unchanged.  Every op requires CUDA tensors and raises otherwise.
"""
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import CHK_ATT, CHK_F32, CHK_F64, CHK_RANK_FMA, CHK_RANK_MMA, CHK_REF, CHK_ROT  # noqa: F401


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return CHK_F32
    if t.dtype == torch.float64:
        return CHK_F64
    raise TypeError(f"chk_b200 supports float32/float64 tables, got {t.dtype}")


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _chk(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("chk_b200 ops need CUDA tensors (there is no CPU fallback)")
        if not t.is_contiguous():
            raise RuntimeError("chk_b200 ops need contiguous tensors")


def _stream():
    return torch.cuda.current_stream().cuda_stream


# number of kernels of libchk_b200.so launched through these wrappers (bench.py reports the delta as gpu_launches)
launch_count = 0


def _launched(n: int) -> None:
    global launch_count
    launch_count += n


def supported_rank(rank: int) -> bool:
    n = 2 * (rank - 1)
    return 16 <= n <= 512 and (n & (n - 1)) == 0


# ------------------------------------------------------------------------------------------- raw calls
GROUPED_MIN_QUERIES = 1 << 16      # from here on the thread-per-query K1 (+ argsort by relation) beats the lane-group kernel


def query_fwd(kind, rank, multi_c, entity, rel, rel_diag, ctx, c_table, head_idx, rel_idx, grouped=None, out=None):
    """get_queries.  grouped=None picks the throughput variant (chk_query_fwd_grouped, queries processed in relation
    order) for large fp32 batches at rank <= 33; True / False force it.  out = (q, c) preallocated outputs."""
    _chk(entity, rel, rel_diag, ctx, c_table, head_idx, rel_idx)
    nq = head_idx.numel()
    if out is not None:
        q, c = out
        _chk(q, c)
    else:
        q = torch.empty((nq, 2 * rank), dtype=entity.dtype, device=entity.device)
        c = torch.empty((nq,), dtype=entity.dtype, device=entity.device)
    can_group = entity.dtype == torch.float32 and rank in (9, 17, 33)
    if grouped is None:
        grouped = can_group and nq >= GROUPED_MIN_QUERIES
    if grouped:
        if not can_group:
            raise RuntimeError("the grouped query transform is fp32, rank in {9, 17, 33} only")
        n_keys = rel.shape[0]
        if n_keys <= 12288:
            perm = torch.empty((nq,), dtype=torch.int32, device=entity.device)
            scratch = torch.empty((n_keys,), dtype=torch.int32, device=entity.device)
            _lib.check(_lib.lib().chk_group_by_key(_p(rel_idx), nq, n_keys, _p(perm), _p(scratch), _stream()), "chk_group_by_key")
            _launched(3)
        else:
            perm = torch.argsort(rel_idx).to(torch.int32)
        _lib.check(_lib.lib().chk_query_fwd_grouped(kind, _dt(entity), rank, nq, int(multi_c), _p(entity), _p(rel),
                                                    _p(rel_diag), _p(ctx), _p(c_table), _p(head_idx), _p(rel_idx), _p(perm),
                                                    _p(q), _p(c), _stream()), "chk_query_fwd_grouped")
        _launched(1)
        return q, c
    _lib.check(_lib.lib().chk_query_fwd(kind, _dt(entity), rank, nq, int(multi_c), _p(entity), _p(rel), _p(rel_diag),
                                        _p(ctx), _p(c_table), _p(head_idx), _p(rel_idx), _p(q), _p(c), _stream()),
               "chk_query_fwd")
    _launched(1)
    return q, c


def query_bwd(kind, rank, multi_c, entity, rel, rel_diag, ctx, c_table, head_idx, rel_idx, grad_q):
    _chk(entity, rel, rel_diag, ctx, c_table, head_idx, rel_idx, grad_q)
    nq = head_idx.numel()
    n = 2 * (rank - 1)
    mk = lambda w: torch.empty((nq, w), dtype=entity.dtype, device=entity.device)
    g_ent, g_rel, g_rd = mk(2 * rank), mk(2 * n), mk(2 * n if kind == CHK_ATT else n)
    g_ctx = mk(n) if kind == CHK_ATT else None
    g_c = mk(1)
    _lib.check(_lib.lib().chk_query_bwd(kind, _dt(entity), rank, nq, int(multi_c), _p(entity), _p(rel), _p(rel_diag),
                                        _p(ctx), _p(c_table), _p(head_idx), _p(rel_idx), _p(grad_q), _p(g_ent),
                                        _p(g_rel), _p(g_rd), _p(g_ctx), _p(g_c), _stream()), "chk_query_bwd")
    _launched(1)
    return g_ent, g_rel, g_rd, g_ctx, g_c


def score_gather_fwd(rank, B, nt, q, q_stride_b, q_stride_j, table, tail_idx, row_stride_b, bh_vals, bh_sb, bh_sj, bt):
    _chk(q, table, tail_idx, bh_vals, bt)
    scores = torch.empty((B, nt), dtype=q.dtype, device=q.device)
    _lib.check(_lib.lib().chk_score_gather_fwd(_dt(q), rank, B, nt, _p(q), q_stride_b, q_stride_j, _p(table),
                                               _p(tail_idx), row_stride_b, _p(bh_vals), bh_sb, bh_sj, _p(bt),
                                               _p(scores), _stream()), "chk_score_gather_fwd")
    _launched(1)
    return scores


def score_gather_bwd(rank, B, nt, q, q_stride_b, q_stride_j, table, tail_idx, row_stride_b, grad_scores):
    _chk(q, table, tail_idx, grad_scores)
    grad_q = torch.empty_like(q)
    grad_rows = torch.empty((B * nt, 2 * rank), dtype=q.dtype, device=q.device)
    _lib.check(_lib.lib().chk_score_gather_bwd(_dt(q), rank, B, nt, _p(q), q_stride_b, q_stride_j, _p(table),
                                               _p(tail_idx), row_stride_b, _p(grad_scores), _p(grad_q),
                                               _p(grad_rows), _stream()), "chk_score_gather_bwd")
    _launched(1)
    return grad_q, grad_rows


def score_gather_bwd_scatter(rank, B, nt, q, q_stride_b, q_stride_j, table, tail_idx, grad_scores, grad_table_dense):
    """K3 adjoint with the tail-row gradients accumulated straight into the dense table gradient."""
    _chk(q, table, tail_idx, grad_scores, grad_table_dense)
    grad_q = torch.empty_like(q)
    _lib.check(_lib.lib().chk_score_gather_bwd_scatter(_dt(q), rank, B, nt, _p(q), q_stride_b, q_stride_j, _p(table),
                                                       _p(tail_idx), _p(grad_scores), _p(grad_q), _p(grad_table_dense),
                                                       _stream()), "chk_score_gather_bwd_scatter")
    _launched(1)
    return grad_q


def nsloss(scores, loss_accum):
    """Negative-sampling loss on scores [B, 1+neg] (column 0 positive): adds the mean loss to loss_accum, returns d/dscores."""
    _chk(scores, loss_accum)
    B, nt = scores.shape
    grad = torch.empty_like(scores)
    _lib.check(_lib.lib().chk_nsloss(_dt(scores), B, nt, _p(scores), _p(loss_accum), _p(grad), _stream()), "chk_nsloss")
    _launched(1)
    return grad


def sparse_adagrad(param, grad, state_sum, rows, lr, eps, stamp, step_id):
    _chk(param, grad, state_sum, rows, stamp, step_id)
    width = param.shape[1] if param.dim() > 1 else 1
    _lib.check(_lib.lib().chk_sparse_adagrad(_dt(param), _p(param), _p(grad), _p(state_sum), _p(rows), rows.numel(), width,
                                             float(lr), float(eps), _p(stamp), _p(step_id), _stream()), "chk_sparse_adagrad")
    _launched(1)


def _descs(tables):
    """tables: list of dicts(param=, grad=, state_sum=, rows=, src_rows=, stamp=) of CUDA tensors (or None)."""
    arr = (_lib.TableDesc * len(tables))()
    for d, t in zip(arr, tables):
        g = t["grad"]
        _chk(t.get("param"), g, t.get("state_sum"), t["rows"], t.get("src_rows"), t.get("stamp"))
        d.param, d.grad, d.state_sum = _p(t.get("param")), _p(g), _p(t.get("state_sum"))
        d.rows, d.m, d.src_rows = _p(t["rows"]), t["rows"].numel(), _p(t.get("src_rows"))
        d.width = g.shape[1] if g.dim() > 1 else 1
        d.stamp = _p(t.get("stamp"))
        if t.get("src_rows") is not None:
            assert t["src_rows"].numel() == d.m * d.width, (t["src_rows"].shape, d.m, d.width)
    return arr


def multi_scatter_add(tables):
    """grad[rows] += src_rows for several tables in one launch."""
    arr = _descs(tables)
    _lib.check(_lib.lib().chk_multi_scatter_add(_dt(tables[0]["grad"]), arr, len(tables), _stream()), "chk_multi_scatter_add")
    _launched(1)


def multi_sparse_adagrad(tables, lr, eps, step_id):
    arr = _descs(tables)
    _chk(step_id)
    _lib.check(_lib.lib().chk_multi_sparse_adagrad(_dt(tables[0]["grad"]), arr, len(tables), float(lr), float(eps),
                                                   _p(step_id), _stream()), "chk_multi_sparse_adagrad")
    _launched(1)


def claim_gather_rows(grad, rows, stamp, step_id, out=None):
    """Send side of the data-parallel sparse exchange: [m, width] gradient rows (duplicate slots carry zeros), the
    claimed rows of the dense ``grad`` are cleared."""
    _chk(grad, rows, stamp, step_id, out)
    width = grad.shape[1] if grad.dim() > 1 else 1
    m = rows.numel()
    if out is None:
        out = torch.empty((m, width), dtype=grad.dtype, device=grad.device)
    _lib.check(_lib.lib().chk_claim_gather_rows(_dt(grad), _p(grad), _p(rows), m, width, _p(stamp), _p(step_id), _p(out),
                                                _stream()), "chk_claim_gather_rows")
    _launched(1)
    return out


def step_counter_bump(counter):
    _chk(counter)
    _lib.check(_lib.lib().chk_step_counter_bump(_p(counter), _stream()), "chk_step_counter_bump")
    _launched(1)


def scatter_add_rows(dense, idx, rows):
    _chk(dense, idx, rows)
    width = dense.shape[1] if dense.dim() > 1 else 1
    n_rows = idx.numel()
    assert rows.numel() == n_rows * width, (rows.shape, n_rows, width)
    _lib.check(_lib.lib().chk_scatter_add_rows(_dt(dense), _p(dense), _p(idx), _p(rows), n_rows, width, _stream()),
               "chk_scatter_add_rows")
    _launched(1)


def row_hnorm(rank, table):
    _chk(table)
    n = table.shape[0]
    out = torch.empty((n,), dtype=table.dtype, device=table.device)
    _lib.check(_lib.lib().chk_row_hnorm(_dt(table), rank, n, _p(table), _p(out), _stream()), "chk_row_hnorm")
    _launched(1)
    return out


def score_all(rank, q, qn, bh_vals, entity, hn, bt):
    _chk(q, qn, bh_vals, entity, hn, bt)
    b, n = q.shape[0], entity.shape[0]
    out = torch.empty((b, n), dtype=q.dtype, device=q.device)
    _lib.check(_lib.lib().chk_score_all(_dt(q), rank, b, _p(q), _p(qn), _p(bh_vals), _p(entity), _p(hn), _p(bt), n,
                                        _p(out), _stream()), "chk_score_all")
    _launched(1)
    return out


def target_scores(rank, q, qn, bh_vals, tail_rows, tail_hn, tail_bt):
    _chk(q, qn, bh_vals, tail_rows, tail_hn, tail_bt)
    b = q.shape[0]
    out = torch.empty((b,), dtype=q.dtype, device=q.device)
    _lib.check(_lib.lib().chk_target_scores(_dt(q), rank, b, _p(q), _p(qn), _p(bh_vals), _p(tail_rows), _p(tail_hn),
                                            _p(tail_bt), _p(out), _stream()), "chk_target_scores")
    _launched(1)
    return out


def rank_counts(algo, rank, q, qn, bh_vals, target, entity, hn, bt, shard_offset, filter_indptr, filter_idx,
                filter_total, counts, shadow=None, workspace=None):
    _chk(q, qn, bh_vals, target, entity, hn, bt, filter_indptr, filter_idx, counts, shadow, workspace)
    assert counts.dtype == torch.int64
    ws_bytes = 0 if workspace is None else workspace.numel() * workspace.element_size()
    _lib.check(_lib.lib().chk_rank_counts(algo, _dt(q), rank, q.shape[0], _p(q), _p(qn), _p(bh_vals), _p(target),
                                          _p(entity), _p(hn), _p(bt), entity.shape[0], shard_offset,
                                          _p(filter_indptr), _p(filter_idx), filter_total, _p(shadow), _p(workspace),
                                          ws_bytes, _p(counts), _stream()), "chk_rank_counts")
    _launched((4 * ((q.shape[0] + 1023) // 1024) if algo == CHK_RANK_MMA else (entity.shape[0] + 64 * 65535 - 1) // (64 * 65535)) + (1 if filter_total > 0 else 0))
    return counts


def entity_shadow(rank, entity, hn=None, bt=None):
    """Shadow of an entity shard for the tcgen05 tier: bf16 hi/lo operand blocks + fp32 epilogue inputs (fp32 or fp64 table)."""
    _chk(entity, hn, bt)
    if hn is None:
        hn = row_hnorm(rank, entity)
    nbytes = _lib.lib().chk_entity_shadow_bytes(rank, entity.shape[0])
    if nbytes <= 0:
        raise RuntimeError("CHK_RANK_MMA shadow unavailable: " + _lib.lib().chk_last_error().decode())
    buf = torch.empty((nbytes,), dtype=torch.uint8, device=entity.device)
    _lib.check(_lib.lib().chk_entity_shadow_build(_dt(entity), rank, entity.shape[0], _p(entity), _p(hn), _p(bt), _p(buf),
                                                  _stream()), "chk_entity_shadow_build")
    _launched(1)
    return buf


def mma_available(rank: int = 257) -> bool:
    """True when the tcgen05 tier (CHK_RANK_MMA) is built into the library."""
    return _lib.lib().chk_entity_shadow_bytes(rank, 128) > 0


def rank_mma_workspace(rank, b, device):
    nbytes = _lib.lib().chk_rank_mma_workspace_bytes(rank, b)
    if nbytes <= 0:
        raise RuntimeError("CHK_RANK_MMA workspace unavailable: " + _lib.lib().chk_last_error().decode())
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=device)
    rank_mma_reset(ws)
    return ws


def rank_mma_reset(workspace):
    _chk(workspace)
    _lib.check(_lib.lib().chk_rank_mma_reset(_p(workspace), _stream()), "chk_rank_mma_reset")


def rank_mma_status(workspace):
    """(length of the last re-check list, sticky overflow flag); synchronises the current stream."""
    import ctypes
    _chk(workspace)
    n, ov = ctypes.c_int64(0), ctypes.c_int(0)
    _lib.check(_lib.lib().chk_rank_mma_status(_p(workspace), ctypes.byref(n), ctypes.byref(ov), _stream()),
               "chk_rank_mma_status")
    return n.value, bool(ov.value)


def rank_mma_profile_events(start=None, stop=None):
    """Measurement support: arm (two recorded torch.cuda.Event(enable_timing=True)) or disarm (None, None) the events the
    library records immediately around rank_mma_kernel inside chk_rank_counts."""
    h = lambda ev: None if ev is None else ev.cuda_event
    _lib.check(_lib.lib().chk_rank_mma_profile_events(h(start), h(stop)), "chk_rank_mma_profile_events")


def score_all_mma(rank, q, qn, bh_vals, target, entity, hn, bt, shadow, workspace):
    """Test support: tensor-core tier approximate scores, error bands and (unfiltered) counts."""
    _chk(q, qn, bh_vals, target, entity, hn, bt, shadow, workspace)
    b, n = q.shape[0], entity.shape[0]
    scores = torch.full((b, n), float("nan"), dtype=torch.float32, device=q.device)
    band = torch.full((b, n), float("nan"), dtype=torch.float32, device=q.device)
    counts = torch.zeros((b,), dtype=torch.int64, device=q.device)
    _lib.check(_lib.lib().chk_score_all_mma(_dt(q), rank, b, _p(q), _p(qn), _p(bh_vals), _p(target), _p(entity), _p(hn),
                                            _p(bt), n, _p(shadow), _p(workspace),
                                            workspace.numel() * workspace.element_size(), _p(counts), _p(scores),
                                            _p(band), _stream()), "chk_score_all_mma")
    return scores, band, counts


# ------------------------------------------------------------------------------------------- autograd
class QueryTransformFn(torch.autograd.Function):
    """get_queries as ONE kernel forward and ONE kernel backward (reference: ~35 eager ops each way)."""

    @staticmethod
    def forward(ctx, meta, entity, rel, rel_diag, ctx_vec, c_table, head_idx, rel_idx):
        kind, rank, multi_c = meta
        q, c = query_fwd(kind, rank, multi_c, entity, rel, rel_diag, ctx_vec, c_table, head_idx, rel_idx)
        ctx.meta = meta
        ctx.save_for_backward(entity, rel, rel_diag, ctx_vec if ctx_vec is not None else entity.new_empty(0), c_table,
                              head_idx, rel_idx)
        ctx.has_ctx = ctx_vec is not None
        ctx.mark_non_differentiable(c)       # the curvature rides along but Distance ignores it (complexhyperbolic.py:59)
        return q, c

    @staticmethod
    def backward(ctx, grad_q, _grad_c):
        kind, rank, multi_c = ctx.meta
        entity, rel, rel_diag, ctx_vec, c_table, head_idx, rel_idx = ctx.saved_tensors
        ctx_vec = ctx_vec if ctx.has_ctx else None
        grads = query_param_grads(kind, rank, multi_c, entity, rel, rel_diag, ctx_vec, c_table, head_idx, rel_idx,
                                  grad_q.contiguous(), None)
        return (None, grads["entity"], grads["rel"], grads["rel_diag"], grads.get("context_vec"), grads["c"], None,
                None)


def query_param_grads(kind, rank, multi_c, entity, rel, rel_diag, ctx_vec, c_table, head_idx, rel_idx, grad_q,
                      g_entity_dense):
    """chk_query_bwd + scatter into DENSE parameter grads (what embedding_dense_backward leaves in .grad)."""
    g_ent, g_rel, g_rd, g_ctx, g_c = query_bwd(kind, rank, multi_c, entity, rel, rel_diag, ctx_vec, c_table,
                                               head_idx, rel_idx, grad_q)
    out = {}
    out["entity"] = torch.zeros_like(entity) if g_entity_dense is None else g_entity_dense
    scatter_add_rows(out["entity"], head_idx, g_ent)
    out["rel"] = torch.zeros_like(rel)
    scatter_add_rows(out["rel"], rel_idx, g_rel)
    out["rel_diag"] = torch.zeros_like(rel_diag)
    scatter_add_rows(out["rel_diag"], rel_idx, g_rd)
    if ctx_vec is not None:
        out["context_vec"] = torch.zeros_like(ctx_vec)
        scatter_add_rows(out["context_vec"], rel_idx, g_ctx)
    out["c"] = torch.zeros_like(c_table)
    if multi_c:
        scatter_add_rows(out["c"], rel_idx, g_c)
    else:
        out["c"] += g_c.sum()
    return out


class ScoreRowsFn(torch.autograd.Function):
    """similarity_score on explicit tensors: q [B,1|nt,2r] vs rhs rows [B,nt,2r] or [1,nt,2r] (no bias)."""

    @staticmethod
    def forward(ctx, rank, q, rhs):
        B = max(q.shape[0], rhs.shape[0])
        nt = max(q.shape[1], rhs.shape[1])
        q2 = q.expand(B, q.shape[1], 2 * rank).contiguous()
        qsj = 1 if q.shape[1] > 1 else 0
        qsb = q2.shape[1]
        rsb = nt if rhs.shape[0] > 1 else 0
        if rhs.shape[1] != nt:
            rhs = rhs.expand(rhs.shape[0], nt, 2 * rank)
        table = rhs.contiguous().view(-1, 2 * rank)
        scores = score_gather_fwd(rank, B, nt, q2.view(-1, 2 * rank), qsb, qsj, table, None, rsb, None, 0, 0, None)
        ctx.cfg = (rank, B, nt, qsb, qsj, rsb, tuple(q.shape), tuple(rhs.shape))
        ctx.save_for_backward(q2, table)
        return scores.unsqueeze(-1)

    @staticmethod
    def backward(ctx, grad):
        rank, B, nt, qsb, qsj, rsb, qshape, rshape = ctx.cfg
        q2, table = ctx.saved_tensors
        gq, grows = score_gather_bwd(rank, B, nt, q2.view(-1, 2 * rank), qsb, qsj, table, None, rsb,
                                     grad.reshape(B, nt).contiguous())
        gq = gq.view(B, -1, 2 * rank)
        if qshape[0] == 1 and B > 1:
            gq = gq.sum(0, keepdim=True)
        grows = grows.view(B, nt, 2 * rank)
        if rshape[0] == 1 and B > 1:
            grows = grows.sum(0, keepdim=True)
        return None, gq, grows


class FusedForwardFn(torch.autograd.Function):
    """KGModel.forward(queries, tails) for the negative-sampling loss: K1 + gather-scoring in two kernels,
    backward = scoring adjoint -> query adjoint -> one dense scatter (reference: two eager graphs through
    get_queries / get_rhs / score, models/base.py:200-226)."""

    @staticmethod
    def forward(ctx, meta, entity, rel, rel_diag, ctx_vec, c_table, bh, bt, queries, tails):
        kind, rank, multi_c, learn_bias = meta
        B, nq1 = queries.shape[0], queries.shape[1]
        nt = tails.shape[1]
        head_idx = queries[..., 0].reshape(-1).contiguous()
        rel_idx = queries[..., 1].reshape(-1).contiguous()
        tails = tails.contiguous()
        q, _ = query_fwd(kind, rank, multi_c, entity, rel, rel_diag, ctx_vec, c_table, head_idx, rel_idx)
        qsj = 1 if nq1 > 1 else 0
        if learn_bias:
            bh_vals = bh.view(-1)[head_idx].contiguous()
            scores = score_gather_fwd(rank, B, nt, q, nq1, qsj, entity, tails, 0, bh_vals, nq1, qsj, bt.view(-1))
        else:
            scores = score_gather_fwd(rank, B, nt, q, nq1, qsj, entity, tails, 0, None, 0, 0, None)
        ctx.meta = meta
        ctx.dims = (B, nq1, nt, qsj)
        ctx.has_ctx = ctx_vec is not None
        ctx.save_for_backward(entity, rel, rel_diag, ctx_vec if ctx_vec is not None else entity.new_empty(0), c_table,
                              bh, bt, head_idx, rel_idx, tails, q)
        return scores.unsqueeze(-1)

    @staticmethod
    def backward(ctx, grad):
        kind, rank, multi_c, learn_bias = ctx.meta
        B, nq1, nt, qsj = ctx.dims
        entity, rel, rel_diag, ctx_vec, c_table, bh, bt, head_idx, rel_idx, tails, q = ctx.saved_tensors
        ctx_vec = ctx_vec if ctx.has_ctx else None
        g = grad.reshape(B, nt).contiguous()
        grad_q, grad_rows = score_gather_bwd(rank, B, nt, q, nq1, qsj, entity, tails, 0, g)
        g_entity = torch.zeros_like(entity)
        scatter_add_rows(g_entity, tails.view(-1), grad_rows)
        grads = query_param_grads(kind, rank, multi_c, entity, rel, rel_diag, ctx_vec, c_table, head_idx, rel_idx,
                                  grad_q, g_entity)
        g_bh = g_bt = None
        if learn_bias:
            g_bh, g_bt = torch.zeros_like(bh), torch.zeros_like(bt)
            gh = g if qsj else g.sum(1)
            scatter_add_rows(g_bh, head_idx, gh.contiguous())
            scatter_add_rows(g_bt, tails.view(-1), g)
        return (None, grads["entity"], grads["rel"], grads["rel_diag"], grads.get("context_vec"), grads["c"], g_bh,
                g_bt, None, None)


# ------------------------------------------------------------------------------------------- fused training step
from ._lib import CHK_HYPER_LEN, CHK_OPT_ADAGRAD, CHK_OPT_ADAM, CHK_OPT_NONE  # noqa: E402,F401


def train_prep(batch, neg, n_entities, double_neg, seed, step_id, stream_id, heads, rels, tails, injected_tails=None,
               injected_heads=None):
    """Device sampler + id arrays of one step (chk_train_prep); heads / rels / tails are preallocated int64 outputs."""
    _chk(batch, step_id, heads, rels, tails, injected_tails, injected_heads)
    B = batch.shape[0]
    _lib.check(_lib.lib().chk_train_prep(_p(batch), B, neg, n_entities, int(double_neg), _p(injected_tails), _p(injected_heads),
                                         int(seed) & 0xFFFFFFFFFFFFFFFF, _p(step_id), int(stream_id) & 0xFFFFFFFF, _p(heads), _p(rels),
                                         _p(tails), _stream()), "chk_train_prep")
    _launched(1)


def score_gather_train(rank, B, nt, q, q_stride_b, q_stride_j, table, tail_idx, head_idx, head_stride_b, head_stride_j, bh, bt,
                       hyper, loss_part, grad_scores, grad_q, grad_rows, g_bh, pair_coef=None):
    """K3 training pass: scores + negative-sampling loss terms + adjoint (all outputs preallocated).  pair_coef [B*nt, 4]:
    the tail-row gradients are left as three scalars per pair for chk_reduce_apply to rebuild (grad_rows may be None)."""
    _chk(q, table, tail_idx, head_idx, bh, bt, hyper, loss_part, grad_scores, grad_q, grad_rows, g_bh, pair_coef)
    _lib.check(_lib.lib().chk_score_gather_train(_dt(q), rank, B, nt, _p(q), q_stride_b, q_stride_j, _p(table), _p(tail_idx),
                                                 _p(head_idx), head_stride_b, head_stride_j, _p(bh), _p(bt), _p(hyper),
                                                 _p(loss_part), _p(grad_scores), _p(grad_q), _p(grad_rows), _p(pair_coef), _p(g_bh),
                                                 _stream()),
               "chk_score_gather_train")
    _launched(1)


def query_bwd_into(kind, rank, multi_c, entity, rel, rel_diag, ctx, c_table, head_idx, rel_idx, grad_q, g_ent, g_rel, g_rd, g_ctx,
                   g_c):
    """chk_query_bwd into preallocated gradient-row buffers."""
    _chk(entity, rel, rel_diag, ctx, c_table, head_idx, rel_idx, grad_q, g_ent, g_rel, g_rd, g_ctx, g_c)
    _lib.check(_lib.lib().chk_query_bwd(kind, _dt(entity), rank, head_idx.numel(), int(multi_c), _p(entity), _p(rel), _p(rel_diag),
                                        _p(ctx), _p(c_table), _p(head_idx), _p(rel_idx), _p(grad_q), _p(g_ent), _p(g_rel), _p(g_rd),
                                        _p(g_ctx), _p(g_c), _stream()), "chk_query_bwd")
    _launched(1)


def group_workspace(n_keys, total_slots, device):
    """Zero-initialised int32 workspace of chk_group_build for a key space of n_keys rows and total_slots slots."""
    nbytes = _lib.lib().chk_group_workspace_bytes(n_keys, total_slots)
    if nbytes <= 0:
        raise RuntimeError(f"chk_group_workspace_bytes({n_keys}, {total_slots}) failed")
    return torch.zeros((nbytes // 4,), dtype=torch.int32, device=device)


def group_build(ids, n_keys, work, own=None):
    """own = (lo, hi): group only the slots whose row lies in [lo, hi) (owner-sharded tables); None = every slot."""
    _chk(ids, work)
    lo, hi = (0, n_keys) if own is None else own
    _lib.check(_lib.lib().chk_group_build(_p(ids), ids.numel(), n_keys, lo, hi, _p(work), _stream()), "chk_group_build")
    _launched(3)


def score_gather_train_peer(rank, B, nt, q, q_stride_b, q_stride_j, peer_tables, peer_bt, rows_per_owner, tail_idx, head_idx,
                            head_stride_b, head_stride_j, bh, hyper, loss_part, grad_scores, grad_q, grad_rows, g_bh, pair_coef=None):
    """K3 training pass on owner-sharded tables: peer_tables / peer_bt are int64 device tensors of `world` base pointers."""
    _chk(q, peer_tables, peer_bt, tail_idx, head_idx, bh, hyper, loss_part, grad_scores, grad_q, grad_rows, g_bh, pair_coef)
    _lib.check(_lib.lib().chk_score_gather_train_peer(_dt(q), rank, B, nt, _p(q), q_stride_b, q_stride_j, _p(peer_tables), _p(peer_bt),
                                                      rows_per_owner, peer_tables.numel(), _p(tail_idx), _p(head_idx), head_stride_b, head_stride_j, _p(bh),
                                                      _p(hyper), _p(loss_part), _p(grad_scores), _p(grad_q), _p(grad_rows),
                                                      _p(pair_coef), _p(g_bh), _stream()), "chk_score_gather_train_peer")
    _launched(1)


def peer_gather_rows(peer_tables, rows_per_owner, ids, width, out):
    """out[i, :] = (table copy of the owner of row ids[i])[ids[i], :]."""
    _chk(peer_tables, ids, out)
    _lib.check(_lib.lib().chk_peer_gather_rows(_dt(out), _p(peer_tables), rows_per_owner, _p(ids), ids.numel(), width, _p(out),
                                               _stream()), "chk_peer_gather_rows")
    _launched(1)


def _red_groups(groups):
    """groups: list of dicts(ids=, n_keys=, slots_per_rank=, world=, work=, single_row=, cols=[dict(param=, state0=, dense=,
    src=[(tensor, lo, hi, rank_stride), ...])]); tensors are kept alive by the caller."""
    arr = (_lib.RedGroup * len(groups))()
    for G, g in zip(arr, groups):
        _chk(g.get("ids"), g.get("work"))
        G.ids, G.n_keys, G.slots_per_rank = _p(g.get("ids")), g.get("n_keys", 1), g["slots_per_rank"]
        G.world, G.n_cols, G.single_row, G.work = g.get("world", 1), len(g["cols"]), int(bool(g.get("single_row"))), _p(g.get("work"))
        for C, c in zip(G.cols, g["cols"]):
            p = c["param"]
            _chk(p, c.get("state0"), c.get("dense"))
            C.param, C.state0, C.dense_grad = _p(p), _p(c.get("state0")), _p(c.get("dense"))
            C.width = p.shape[1] if p.dim() > 1 else 1
            for i, (t, lo, hi, rs) in enumerate(c["src"]):
                _chk(t)
                C.src[i], C.lo[i], C.hi[i], C.rank_stride[i] = _p(t), lo, hi, rs
            pc = c.get("pair")                                   # (coef tensor, pair_nt, coef_rank_stride): src[1] holds query rows
            if pc is not None:
                _chk(pc[0])
                C.pair_coef, C.pair_nt, C.coef_rank_stride = _p(pc[0]), pc[1], pc[2]
    return arr


def reduce_apply(dtype_of, opt, groups, hyper, finish=None):
    """Segment-reduce the contribution rows of every touched row and apply Adagrad in place / write the dense gradient.
    finish = (loss_part, loss_accum, step_id): the kernel's last block also ends the step (chk_step_finish's duties)."""
    _chk(hyper)
    arr = groups if not isinstance(groups, list) else _red_groups(groups)
    lp = la = sid = None
    if finish is not None:
        lp, la, sid = finish
        _chk(lp, la, sid)
    _lib.check(_lib.lib().chk_reduce_apply(_dt(dtype_of), opt, arr, len(arr), _p(hyper), int(finish is not None), _p(lp),
                                           0 if lp is None else lp.numel(), _p(la), _p(sid), _stream()), "chk_reduce_apply")
    _launched(1)


def step_finish(dtype_of, works, loss_part, loss_accum, step_id):
    import ctypes
    _chk(loss_part, loss_accum, step_id, *works)
    arr = (ctypes.c_void_p * max(len(works), 1))(*[w.data_ptr() for w in works])
    _lib.check(_lib.lib().chk_step_finish(_dt(dtype_of), arr, len(works), _p(loss_part), 0 if loss_part is None else loss_part.numel(),
                                          _p(loss_accum), _p(step_id), _stream()), "chk_step_finish")
    _launched(1)


def dense_apply(opt, tables, hyper, step_id):
    """tables: list of (param, grad, state0, state1|None) — torch.optim.Adagrad / Adam over whole tables; grads are cleared."""
    arr = (_lib.DenseTab * len(tables))()
    for d, (p, g, s0, s1) in zip(arr, tables):
        _chk(p, g, s0, s1)
        assert g.numel() == p.numel() and s0.numel() == p.numel()
        d.param, d.grad, d.state0, d.state1, d.n = _p(p), _p(g), _p(s0), _p(s1), p.numel()
    _chk(hyper, step_id)
    _lib.check(_lib.lib().chk_dense_apply(_dt(tables[0][0]), opt, arr, len(tables), _p(hyper), _p(step_id), _stream()), "chk_dense_apply")
    _launched(1)


def dp_fused_apply(dtype_of, opt, world, rank, peer_grad, peer_param, peer_s0, peer_s1, peer_sig, n, hyper, step_id, local_state):
    """Dense data-parallel step over peer memory: reduce-scatter of the flat gradients + optimizer + broadcast of the new
    values, then wait + clear (chk_dp_fused_apply).  peer_*: int64 device tensors of `world` base pointers."""
    _chk(peer_grad, peer_param, peer_s0, peer_s1, peer_sig, hyper, step_id, local_state)
    _lib.check(_lib.lib().chk_dp_fused_apply(_dt(dtype_of), opt, world, rank, _p(peer_grad), _p(peer_param), _p(peer_s0), _p(peer_s1),
                                             _p(peer_sig), n, _p(hyper), _p(step_id), _p(local_state), _stream()), "chk_dp_fused_apply")
    _launched(2)


def dp_all_gather(world, rank, peer_src, bytes_per_rank, dst, peer_sig, channel, local_state):
    """all_gather by peer reads behind a flag barrier (chk_dp_all_gather)."""
    _chk(peer_src, dst, peer_sig, local_state)
    _lib.check(_lib.lib().chk_dp_all_gather(world, rank, _p(peer_src), bytes_per_rank, _p(dst), _p(peer_sig), channel, _p(local_state),
                                            _stream()), "chk_dp_all_gather")
    _launched(1)


def rowsum_groups(src, B, nj, width, out):
    _chk(src, out)
    _lib.check(_lib.lib().chk_rowsum_groups(_dt(src), _p(src), B, nj, width, _p(out), _stream()), "chk_rowsum_groups")
    _launched(1)


def reg_factors(power, weight, hyper, B, entity, rel, heads, head_stride, rels, tails, tail_stride, g_ent, g_ent_stride, g_rel,
                g_rel_stride, g_tail, g_tail_stride, loss_part):
    _chk(hyper, entity, rel, heads, rels, tails, g_ent, g_rel, g_tail, loss_part)
    _lib.check(_lib.lib().chk_reg_factors(_dt(entity), power, float(weight), _p(hyper), B, _p(entity), entity.shape[1], _p(rel),
                                          rel.shape[1], _p(heads), head_stride, _p(rels), _p(tails), tail_stride, _p(g_ent),
                                          g_ent_stride, _p(g_rel), g_rel_stride, _p(g_tail), g_tail_stride, _p(loss_part), _stream()),
               "chk_reg_factors")
    _launched(1)


# ------------------------------------------------------------------------------------------- evaluation batch in one call
def eval_scratch(rank, b, dtype, device):
    n = _lib.lib().chk_eval_scratch_bytes(CHK_F32 if dtype == torch.float32 else CHK_F64, rank, b)
    if n <= 0:
        raise RuntimeError("chk_eval_scratch_bytes failed")
    return torch.empty((n,), dtype=torch.uint8, device=device)


def eval_batch(algo, kind, rank, multi_c, queries, entity, rel, rel_diag, ctx, c_table, bh, bt, hn_full, shard_entity, shard_hn,
               shard_bt, shard_offset, shadow, workspace, f_keys, f_indptr, f_vals, n_rel2, scratch, counts, target, flags):
    """One batch of get_ranking: K1 -> norms -> target -> rank counts -> device filter lookup -> filter pass (chk_eval_batch)."""
    _chk(queries, entity, rel, rel_diag, ctx, c_table, bh, bt, hn_full, shard_entity, shard_hn, shard_bt, shadow, workspace, f_keys,
         f_indptr, f_vals, scratch, counts, target, flags)
    a = _lib.EvalArgs()
    a.algo, a.kind, a.dtype, a.rank, a.multi_c, a.b = algo, kind, _dt(entity), rank, int(multi_c), queries.shape[0]
    a.queries = _p(queries)
    a.entity, a.rel, a.rel_diag, a.ctx, a.c_table, a.bh, a.bt = _p(entity), _p(rel), _p(rel_diag), _p(ctx), _p(c_table), _p(bh), _p(bt)
    a.hn_full, a.n_entities = _p(hn_full), entity.shape[0]
    a.shard_entity, a.shard_hn, a.shard_bt = _p(shard_entity), _p(shard_hn), _p(shard_bt)
    a.shard_rows, a.shard_offset = shard_entity.shape[0], shard_offset
    a.shadow, a.workspace = _p(shadow), _p(workspace)
    a.workspace_bytes = 0 if workspace is None else workspace.numel() * workspace.element_size()
    a.f_keys, a.f_indptr, a.f_vals, a.f_nkeys, a.n_rel2 = _p(f_keys), _p(f_indptr), _p(f_vals), f_keys.numel(), n_rel2
    a.scratch, a.scratch_bytes = _p(scratch), scratch.numel()
    a.counts, a.target, a.flags = _p(counts), _p(target), _p(flags)
    import ctypes
    _lib.check(_lib.lib().chk_eval_batch(ctypes.byref(a), _stream()), "chk_eval_batch")
    b = queries.shape[0]
    _launched(6 + (4 * ((b + 1023) // 1024) if algo == CHK_RANK_MMA else (shard_entity.shape[0] + 64 * 65535 - 1) // (64 * 65535)))
