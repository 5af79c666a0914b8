/* chk_b200.h — C ABI of the B200-native FFTRotH / FFTRefH / FFTAttH scoring hot path.
 *
 * Drop-in boundary for htmai-880/ComplexHyperbolicKGE (reference paths are relative to its root).
 * Every entry point replaces a piece of the reference's eager-PyTorch code that a binding
 * (ctypes) would call instead.  Conventions:
 *   - all pointers are DEVICE pointers (cudaMalloc'ed / torch CUDA tensors), contiguous, row-major;
 *   - `dtype` = CHK_F32 | CHK_F64 selects the scalar type of every `void*` table / vector
 *     (models/base.py:35-39 `--dtype float|double`); indices are int64 (torch.LongTensor);
 *   - `rank` r: entity rows are 2r wide = [Re X_0..X_{r-1} | Im X_0..X_{r-1}]; n = 2(r-1) must be a
 *     power of two with 16 <= n <= 512 (r in {9,17,33,65,129,257});
 *   - `stream` is a cudaStream_t (the caller's current stream); launches are asynchronous;
 *   - the library never allocates or frees memory it returns, holds no mutable global state besides per-thread
 *     items (the last-error string, the measurement events of chk_rank_mma_profile_events) and a cached SM count,
 *     and never calls back into the host language;
 *   - return value 0 = success, otherwise a CHK_E* code; chk_last_error() gives the text.
 * There is NO CPU fallback: without a CUDA device every compute entry returns CHK_ECUDA.
 */
#ifndef CHK_B200_H
#define CHK_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { CHK_ROT = 0, CHK_REF = 1, CHK_ATT = 2 };          /* FFTRotH / FFTRefH / FFTAttH */
enum { CHK_F32 = 0, CHK_F64 = 1 };
enum { CHK_OK = 0, CHK_EINVAL = 1, CHK_ECUDA = 2, CHK_EUNSUPPORTED = 3, CHK_EOVERFLOW = 4 };
enum { CHK_RANK_FMA = 0, CHK_RANK_MMA = 1 };             /* chk_rank_counts algorithm */

int chk_abi_version(void);
const char* chk_last_error(void);

/* ---- K1: query transform (get_queries) --------------------------------------------------------
 * Replaces FFTRotH/FFTRefH/FFTAttH.get_queries (models/complexhyperbolic.py:79-101,107-127,144-171):
 * irfft -> expmap0 -> Moebius translation -> Givens rotation / "reflection" / attention -> rfft, with
 * utils/complexhyperbolic.py:41-54,72-106 and utils/euclidean.py:26-75 fused, one warp per query.
 * c_table is c.weight: (R2,1) raw (softplus applied inside) when multi_c, else (1,1) used RAW.
 * out_q [nq,2r], out_c [nq] (the curvature actually used).  ctx may be NULL unless kind==CHK_ATT. */
int chk_query_fwd(int kind, int dtype, int rank, int64_t nq, int multi_c,
                  const void* entity, const void* rel, const void* rel_diag, const void* ctx,
                  const void* c_table, const int64_t* head_idx, const int64_t* rel_idx,
                  void* out_q, void* out_c, void* stream);

/* Same result as chk_query_fwd, THROUGHPUT variant for large batches (fp32, rank in {9,17,33}): one THREAD per query
 * (register FFT with compile-time twiddles, thread-local norms, entity rows staged through shared memory with
 * asynchronous copies), 2x the lane-group kernel at 2^20 queries when the 32 queries of a warp share their relation.
 * perm (int32 [nq], may be NULL) is the order in which queries are processed — pass argsort(rel_idx) to group them by
 * relation; outputs are written at the original query positions.  Latency at small nq is worse than chk_query_fwd. */
/* perm = the positions 0..n-1 grouped by keys[i] in [0, n_keys) (counting sort; order inside a group unspecified).
 * counts_scratch: int32 [n_keys].  n < 2^31, n_keys <= 12288. */
int chk_group_by_key(const int64_t* keys, int64_t n, int n_keys, int32_t* perm, int32_t* counts_scratch, void* stream);
int chk_query_fwd_grouped(int kind, int dtype, int rank, int64_t nq, int multi_c,
                          const void* entity, const void* rel, const void* rel_diag, const void* ctx,
                          const void* c_table, const int64_t* head_idx, const int64_t* rel_idx, const int32_t* perm,
                          void* out_q, void* out_c, void* stream);

/* Adjoint of chk_query_fwd (what autograd does through the ~35 eager ops of get_queries).
 * grad_q [nq,2r] in; per-query gradient ROWS out (the caller scatters them with chk_scatter_add_rows):
 * g_entity_rows [nq,2r], g_rel_rows [nq,2n], g_rel_diag_rows [nq,n] ([nq,2n] for ATT),
 * g_ctx_rows [nq,n] (ATT only, else NULL), g_c [nq] (w.r.t. the RAW c parameter). */
int chk_query_bwd(int kind, int dtype, int rank, int64_t nq, int multi_c,
                  const void* entity, const void* rel, const void* rel_diag, const void* ctx,
                  const void* c_table, const int64_t* head_idx, const int64_t* rel_idx,
                  const void* grad_q,
                  void* g_entity_rows, void* g_rel_rows, void* g_rel_diag_rows, void* g_ctx_rows,
                  void* g_c, void* stream);

/* ---- K3: scoring of gathered tails, forward + backward (training / similarity_score) -----------
 * Replaces KGModel.get_rhs + score + FFTUnitBall.similarity_score + Distance.forward/backward
 * (models/base.py:108-133,148-173; models/complexhyperbolic.py:45-59;
 *  utils/complexhyperbolic.py:176-254, lift=True semantics) for B queries x nt tails.
 * q row of pair (b,j) = q + (b*q_stride_b + j*q_stride_j)*2r   (q_stride_j = 0: one query per b).
 * tail row of pair (b,j) = table[tail_idx[b*nt+j]] if tail_idx else table[b*row_stride_b + j].
 * scores[b*nt+j] = (bh_vals[b*bh_stride_b + j*bh_stride_j] + bt[row]) + (-acosh(x)^2); bh_vals/bt may
 * both be NULL (bias none). */
int chk_score_gather_fwd(int dtype, int rank, int64_t B, int64_t nt,
                         const void* q, int64_t q_stride_b, int64_t q_stride_j,
                         const void* table, const int64_t* tail_idx, int64_t row_stride_b,
                         const void* bh_vals, int64_t bh_stride_b, int64_t bh_stride_j, const void* bt,
                         void* scores, void* stream);

/* grad_scores [B,nt] -> grad_q rows (same strides as q; must be pre-zeroed by the caller only if
 * q_stride_j==0 is NOT used — when q_stride_j==0 the kernel reduces over j and overwrites row b) and
 * grad_rows [B*nt,2r] (one row per pair, to be scattered by the caller).  Straight-through at the
 * clamps exactly as Distance.backward (utils/complexhyperbolic.py:202-203,231-234). */
int chk_score_gather_bwd(int dtype, int rank, int64_t B, int64_t nt,
                         const void* q, int64_t q_stride_b, int64_t q_stride_j,
                         const void* table, const int64_t* tail_idx, int64_t row_stride_b,
                         const void* grad_scores,
                         void* grad_q, void* grad_rows, void* stream);

/* Same adjoint, but the tail-row gradients are accumulated (atomic adds) straight into the DENSE gradient of the
 * table, grad_table_dense[tail_idx[b*nt+j], :] += row gradient — no [B*nt,2r] temporary, no second pass. */
int chk_score_gather_bwd_scatter(int dtype, int rank, int64_t B, int64_t nt,
                                 const void* q, int64_t q_stride_b, int64_t q_stride_j,
                                 const void* table, const int64_t* tail_idx,
                                 const void* grad_scores, void* grad_q, void* grad_table_dense, void* stream);

/* dense[idx[i], :] += rows[i, :]   (embedding_dense_backward of entity/rel/bias tables). */
int chk_scatter_add_rows(int dtype, void* dense, const int64_t* idx, const void* rows,
                         int64_t n_rows, int64_t width, void* stream);

/* ---- training-loop kernels next to the path ------------------------------
 * Negative-sampling loss of KGOptimizer.neg_sampling_loss (optimizers/kg_optimizer.py:115-122) on scores [B,nt]
 * whose column 0 is the positive tail: *loss_accum += -mean(cat[logsigmoid(s[:,0]), logsigmoid(-s[:,1:])]) over
 * B*nt terms (the caller zeroes loss_accum), grad_scores [B,nt] = d loss / d scores. */
int chk_nsloss(int dtype, int64_t B, int64_t nt, const void* scores, void* loss_accum, void* grad_scores, void* stream);
/* Row-sparse torch.optim.Adagrad step (lr_decay = 0, weight_decay = 0; run.py:205 hands the dense tables to
 * torch.optim): for every row listed in rows[m] — once per step even if listed several times —
 * sum += g*g; param -= lr * g / (sqrt(sum) + eps); g = 0.  Zero-gradient rows are no-ops in Adagrad, so this equals
 * the dense update.  stamp is an int32 [n_rows] scratch (zero-initialised once), *step_id a device counter that
 * differs from every earlier step's value (chk_step_counter_bump). */
int chk_sparse_adagrad(int dtype, void* param, void* grad, void* state_sum, const int64_t* rows, int64_t m,
                       int64_t width, double lr, double eps, int32_t* stamp, const int32_t* step_id, void* stream);
int chk_step_counter_bump(int32_t* counter, void* stream);
/* Data-parallel training, send side of the sparse embedding-gradient exchange (the reference has no
 * distributed code): out_rows[t, :] = the accumulated gradient row rows[t] of the dense `grad` if slot t is the first
 * slot of this step that names the row, zeros otherwise (so duplicates are sent once); claimed rows of `grad` are
 * cleared.  stamp / step_id are the ones of chk_sparse_adagrad (a different token is used, call this BEFORE it).
 * Fixed sizes, no host synchronisation: the ranks all_gather out_rows + rows and add them back rank by rank with
 * chk_multi_scatter_add, which makes the summed gradient bit-identical on every replica. */
int chk_claim_gather_rows(int dtype, void* grad, const int64_t* rows, int64_t m, int64_t width, int32_t* stamp,
                          const int32_t* step_id, void* out_rows, void* stream);
/* The same two operations over several tables in ONE launch each (entity, rel, rel_diag, context_vec, c, bh, bt):
 *   chk_multi_scatter_add:     grad[rows[i], :] += src_rows[i, :]            (tables with src_rows == NULL are skipped)
 *   chk_multi_sparse_adagrad:  the chk_sparse_adagrad update on rows[0..m) of every table. */
#define CHK_MAX_TABLES 8
typedef struct chk_table_desc {
    void* param; void* grad; void* state_sum;   /* dense [n_rows, width] tensors (param / state_sum unused by the scatter) */
    const int64_t* rows; int64_t m;             /* row ids touched in this step (duplicates allowed) */
    const void* src_rows;                       /* scatter only: [m, width] rows to add */
    int64_t width;
    int32_t* stamp;                             /* adagrad only: int32 [n_rows] scratch */
} chk_table_desc;
int chk_multi_scatter_add(int dtype, const chk_table_desc* tabs, int n_tables, void* stream);
int chk_multi_sparse_adagrad(int dtype, const chk_table_desc* tabs, int n_tables, double lr, double eps,
                             const int32_t* step_id, void* stream);

/* ---- fused training step: sampler, fused loss, segment-reduce + optimizer -------------------
 * One optimisation step of KGOptimizer.epoch (optimizers/kg_optimizer.py:239-277) = chk_train_prep -> chk_query_fwd ->
 * chk_score_gather_train -> chk_query_bwd -> [chk_group_build beside them] -> chk_reduce_apply -> chk_step_finish, with no
 * host synchronisation, no floating-point atomics (bit-reproducible) and no dense N x 2r gradient.
 *
 * Device scalars shared by the step kernels, `hyper` (double[CHK_HYPER_LEN], written by the host between steps):
 *   [0] lr  [1] eps  [2] 1/(loss terms of the GLOBAL batch = rows*(1+neg))  [3] valid rows of this rank's batch (rows beyond
 *   it are padding: zero loss, zero gradient)  [4] beta1  [5] beta2  [6] 1/(rows of the global batch)  [7] reserved. */
#define CHK_HYPER_LEN 8
enum { CHK_OPT_NONE = 0, CHK_OPT_ADAGRAD = 1, CHK_OPT_ADAM = 2 };

/* KGOptimizer.get_neg_samples (optimizers/kg_optimizer.py:92-99) on the device: tails[b,0] = batch[b,2], tails[b,j>=1] uniform
 * over the entities != batch[b,2] (Philox4x32-10 keyed by `seed`, counter = (b*neg+j-1, *step_id, stream_id)), or copied from
 * injected_tails [B,neg] when given (an overridden sampler).  double_neg = 0: heads [B], rels [B] are columns 0, 1 of the batch.
 * double_neg = 1 (semantics of the commented-out original, :78-91: every negative also corrupts the head): heads, rels are
 * [B,1+neg]; heads[b,0] = batch[b,0], heads[b,j>=1] uniform over the entities != batch[b,0] (or injected_heads), rels[b,:] = r. */
int chk_train_prep(const int64_t* batch, int64_t B, int64_t neg, int64_t n_entities, int double_neg,
                   const int64_t* injected_tails, const int64_t* injected_heads, uint64_t seed,
                   const int32_t* step_id, uint32_t stream_id, int64_t* heads, int64_t* rels, int64_t* tails, void* stream);

/* K3 training pass: scores of the (B, nt) gathered tails, the loss terms of KGOptimizer.neg_sampling_loss (:115-122; column 0
 * positive: -logsigmoid(s), others -logsigmoid(-s), scaled by hyper[2]) and the adjoint, tail rows gathered once.
 * Out: loss_part [B] (row sums of the scaled terms), grad_scores [B,nt] (= the bt gradient of every pair), grad_q (strides of q;
 * reduced over j when q_stride_j == 0), grad_rows [B*nt,2r], g_bh [B] (sum_j grad_scores; NULL with per-pair queries).
 * pair_coef (optional, [B*nt,4]): when given, the tail-row gradient of a pair is NOT stored (grad_rows may be NULL); the kernel
 * writes its three pair scalars (c1, c2, c3, 0) and chk_reduce_apply rebuilds the row from them, the query row z and the tail
 * row w it updates:  g_re[k] = c1 z_re[k] + c2 z_im[k] - c3 w_re[k],  g_im[k] = c1 z_im[k] - c2 z_re[k] - c3 w_im[k]
 * (same operations, same order as the stored row: bit-identical).  16 bytes per pair instead of 8r (4r in fp32).
 * bias: bh value of pair (b,j) = bh[head_idx[b*head_stride_b + j*head_stride_j]]; bh/bt both NULL for bias 'none'. */
int chk_score_gather_train(int dtype, int rank, int64_t B, int64_t nt,
                           const void* q, int64_t q_stride_b, int64_t q_stride_j,
                           const void* table, const int64_t* tail_idx,
                           const int64_t* head_idx, int64_t head_stride_b, int64_t head_stride_j,
                           const void* bh, const void* bt, const double* hyper,
                           void* loss_part, void* grad_scores, void* grad_q, void* grad_rows, void* pair_coef, void* g_bh,
                           void* stream);

/* Owner-sharded tables (data parallel; no reference counterpart — optimizers/kg_optimizer.py:51 is single-device): every rank
 * holds a full-layout copy of the entity / bt tables in peer-accessible (symmetric) memory, but only the rows it OWNS
 * (row / rows_per_owner == rank) are current.  chk_score_gather_train_peer reads every tail row and bt value from its owner's
 * copy over NVLink (peer_tables / peer_bt: device arrays of `world` <= 32 base pointers); chk_peer_gather_rows copies the rows `ids`
 * of such a table into a local buffer (head rows for K1 and its adjoint, head biases).  bh: local [B]-indexable values. */
int chk_score_gather_train_peer(int dtype, int rank, int64_t B, int64_t nt,
                                const void* q, int64_t q_stride_b, int64_t q_stride_j,
                                const void* const* peer_tables, const void* const* peer_bt, int64_t rows_per_owner, int world,
                                const int64_t* tail_idx,
                                const int64_t* head_idx, int64_t head_stride_b, int64_t head_stride_j,
                                const void* bh, const double* hyper,
                                void* loss_part, void* grad_scores, void* grad_q, void* grad_rows, void* pair_coef, void* g_bh,
                                void* stream);
int chk_peer_gather_rows(int dtype, const void* const* peer_tables, int64_t rows_per_owner, const int64_t* ids, int64_t n,
                         int64_t width, void* out, void* stream);

/* Grouping of `total_slots` slots by the table row they name (ids[s] in [0, n_keys)): fills the workspace (int32, size
 * chk_group_workspace_bytes, zero-initialised once by the caller; chk_reduce_apply / chk_step_finish leave it ready for
 * the next step) with per-row counts, segment bases, the slot order and the list of touched rows.  Only slots whose row lies in
 * [own_lo, own_hi) are grouped (0, n_keys: all; an owner-sharded table: the rank's own row block) — chk_reduce_apply then
 * touches nothing else. */
int64_t chk_group_workspace_bytes(int64_t n_keys, int64_t total_slots);
int chk_group_build(const int64_t* ids, int64_t total_slots, int64_t n_keys, int64_t own_lo, int64_t own_hi, void* work, void* stream);

/* Segment-reduce + apply.  A group = one key space (entity ids; relation ids) with world*slots_per_rank slots numbered
 * rank-major (slot = k*slots_per_rank + s); a column = one table fed by up to two contribution sources: local slot s in
 * [lo[i], hi[i]) of rank k contributes the row src[i] + k*rank_stride[i] + (s-lo[i])*width (elements).  Every touched row's
 * contributions are summed in ascending slot order by one warp, then either the torch.optim.Adagrad update (lr_decay = 0,
 * weight_decay = 0: sum += g*g; p -= lr*g/(sqrt(sum)+eps), hyper[0..1]) is applied to the row in place (dense_grad NULL,
 * opt = CHK_OPT_ADAGRAD, state0 = the optimizer's `sum`), or the row sum is WRITTEN to dense_grad[row] (other rows untouched).
 * single_row: every slot names row 0 (the (1,1) curvature table), no grouping workspace.  width must be 1 or even. */
#define CHK_RED_MAX_COLS 4
#define CHK_RED_MAX_GROUPS 3
typedef struct chk_red_col {
    void* param; void* state0; void* dense_grad; int64_t width;
    const void* src[2]; int64_t lo[2], hi[2], rank_stride[2];
    /* Computed source (pair_coef NULL = off; first column of group 0 only, rank in {9,17,33,65,129,257}): the slots
     * [lo[1], hi[1]) are (query, tail) pairs whose contribution row is rebuilt, not read: src[1] holds the QUERY rows
     * ([re | im], `width` elements, rank_stride[1] between ranks) and pair_coef the (c1, c2, c3, pad) chk_score_gather_train
     * wrote per pair (coef_rank_stride elements between ranks).  Pair p = slot - lo[1] uses query row p / pair_nt
     * (pair_nt >= 1: one query per pair_nt pairs) or p (pair_nt == 0: per-pair queries, double_neg). */
    const void* pair_coef; int64_t pair_nt; int64_t coef_rank_stride;
} chk_red_col;
typedef struct chk_red_group {
    const int64_t* ids; int64_t n_keys; int64_t slots_per_rank; int32_t world; int32_t n_cols; int32_t single_row; int32_t pad_;
    void* work; chk_red_col cols[CHK_RED_MAX_COLS];
} chk_red_group;
/* finish_step != 0: the last block of the launch also does chk_step_finish's duties (headers reset, loss partials summed
 * in a fixed order into *loss_accum, *step_id bumped; loss_part / step_id may be NULL), saving that launch. */
int chk_reduce_apply(int dtype, int opt, const chk_red_group* groups, int n_groups, const double* hyper,
                     int finish_step, const void* loss_part, int64_t n_loss, void* loss_accum, int32_t* step_id, void* stream);
/* End of a step: resets the grouping headers, adds the per-row loss partials (fixed order) to *loss_accum, bumps *step_id.
 * loss_part / step_id may be NULL. */
int chk_step_finish(int dtype, void* const* group_works, int n_groups, const void* loss_part, int64_t n_loss, void* loss_accum,
                    int32_t* step_id, void* stream);
/* torch.optim.Adagrad / torch.optim.Adam (defaults: no weight decay, no amsgrad; hyper[0,1,4,5], *step_id = 1-based step) over
 * whole tables from a dense gradient, which is cleared (run.py:205 hands the dense tables to torch.optim). */
typedef struct chk_dense_tab { void* param; void* grad; void* state0; void* state1; int64_t n; } chk_dense_tab;
int chk_dense_apply(int dtype, int opt, const chk_dense_tab* tabs, int n_tables, const double* hyper, const int32_t* step_id, void* stream);
/* Data-parallel dense-gradient step over NVLink peer memory (no NCCL call; no reference counterpart): the flat gradient, parameter
 * and optimizer-state buffers of every rank live in symmetric memory with the same layout (peer_* = device arrays of `world`
 * base pointers, n elements each, n a multiple of 4 * world, 16-byte aligned).  Every rank sums its 1/world slice of the `world` gradient buffers in ascending rank order,
 * applies torch.optim.Adagrad (state0 = sum) / Adam (state0, state1 = exp_avg, exp_avg_sq; *step_id = 1-based step) to it and
 * stores the new parameter / state values into EVERY replica; a second kernel waits for all slices and clears the local
 * gradient buffer.  peer_signal: int32[4 * world] per rank (symmetric), zero-initialised; local_state: int32[8], zero-initialised
 * ([2] becomes non-zero if a peer did not arrive within ~4 s).  Every rank must make the same sequence of calls. */
int chk_dp_fused_apply(int dtype, int opt, int world, int rank, const void* const* peer_grad, void* const* peer_param,
                       void* const* peer_state0, void* const* peer_state1, int32_t* const* peer_signal, int64_t n,
                       const double* hyper, const int32_t* step_id, int32_t* local_state, void* stream);

/* all_gather by peer reads (no NCCL call): a flag barrier ("my block is complete"; it also orders everything before the call on
 * every rank against everything after it on every other rank), then dst[k] = the bytes_per_rank bytes at peer_src[k] for every
 * k.  peer_src: device array of `world` base pointers of a symmetric buffer; bytes_per_rank a multiple of 8; channel 0 / 1 = two
 * independent flag sets in the peer_signal / local_state arrays of chk_dp_fused_apply. */
int chk_dp_all_gather(int world, int rank, const void* const* peer_src, int64_t bytes_per_rank, void* dst,
                      int32_t* const* peer_signal, int channel, int32_t* local_state, void* stream);

/* out[b,:] = sum_j in[b,j,:] in a fixed order — eight interleaved partial sums (j mod 8), added in ascending order — so the
 * result is bit-reproducible (double_neg: the nt per-pair relation-row gradients of a triple share a row). */
int chk_rowsum_groups(int dtype, const void* in, int64_t B, int64_t nj, int64_t width, void* out, void* stream);
/* N3 (power 3) / F2 (power 2) of the positive call's factors entity[h], rel[r], entity[t] (optimizers/regularizers.py:21-58):
 * value weight*sum|f|^p*hyper[6] added to loss_part[b], gradient rows added to the three contribution rows of triple b. */
int chk_reg_factors(int dtype, int power, double weight, const double* hyper, int64_t B,
                    const void* entity, int64_t ent_width, const void* rel, int64_t rel_width,
                    const int64_t* heads, int64_t head_stride, const int64_t* rels, const int64_t* tails, int64_t tail_stride,
                    void* g_ent_rows, int64_t g_ent_stride, void* g_rel_rows, int64_t g_rel_stride,
                    void* g_tail_rows, int64_t g_tail_stride, void* loss_part, void* stream);

/* ---- K2: scoring against the whole entity table + filtered rank counts (evaluation) -------------
 * All K2 entry points share ONE canonical pair-score arithmetic (ascending-k FMA chain, see
 * csrc/chk_common.cuh) so that target scores, tile counts, the filter pass and the exact re-check of the
 * tensor-core tier agree bit-for-bit; exact ties (clamp regime) therefore behave like the reference's.
 *
 * Clamped Hermitian norm of every row: hn[i] = clamp(sum|w_i|^2 - 1, -1, -eps)
 * (utils/complexhyperbolic.py:187-188,229-230).  Used for the entity table (once per evaluation pass)
 * and for the query batch (qn). */
int chk_row_hnorm(int dtype, int rank, int64_t n_rows, const void* table, void* hn, void* stream);

/* scores[i, e] for all e in the shard (models/base.py:255 `score(q, candidates)`).  q [b,2r], qn [b],
 * hn/bt [n_rows]; bh_vals/bt both NULL for bias 'none'.  scores [b, n_rows]. */
int chk_score_all(int dtype, int rank, int64_t b, const void* q, const void* qn, const void* bh_vals,
                  const void* entity, const void* hn, const void* bt, int64_t n_rows,
                  void* scores, void* stream);

/* target[i] = score(q_i, tail_rows_i) (models/base.py:256).  tail_rows [b,2r], tail_hn [b], tail_bt [b]
 * are the gathered rows of the true tails (gathered by the caller so that a sharded table works). */
int chk_target_scores(int dtype, int rank, int64_t b, const void* q, const void* qn, const void* bh_vals,
                      const void* tail_rows, const void* tail_hn, const void* tail_bt,
                      void* target, void* stream);

/* Filtered rank counts of KGModel.get_ranking (models/base.py:255-271), fused: for query i
 *   counts[i] += #{ e in shard, e not in filter_i : score(i,e) >= target[i] }
 * filter_i = filter_idx[filter_indptr[i] .. filter_indptr[i+1]) holds GLOBAL entity ids, UNIQUE per
 * query, and must contain the true tail (the reference appends it, base.py:267); filter_total =
 * filter_indptr[b].  Shard rows are global ids [shard_offset, shard_offset + n_rows).  counts is int64 [b],
 * accumulated (caller zeroes it); summing it over shards / GPUs gives rank - 1.
 * algo = CHK_RANK_FMA: exact fp32/fp64 FMA tiles.
 * algo = CHK_RANK_MMA (fp32 AND fp64 models): the contraction runs on tcgen05 (bf16x3 split, fp32 accumulation in
 *   TMEM) from the pre-tiled bf16 shadow of chk_entity_shadow_build; the epilogue decides every pair whose
 *   approximate score is outside a proven error band around the target, and the pairs inside the band are
 *   re-scored with the exact chain IN THE MODEL'S dtype, so the counts EQUAL CHK_RANK_FMA's (for --dtype double:
 *   fp64-exact ranks at tensor-core speed; tcgen05 has no f64 kind, the fp64 work is only the re-check).  Needs `shadow` (built for the same entity/n_rows) and
 *   a workspace of chk_rank_mma_workspace_bytes(rank, b) bytes.  If the band list overflows, a sticky flag is
 *   raised in the workspace (read it with chk_rank_mma_status after the pass; the counts of that pass are then
 *   invalid and the caller re-runs it with CHK_RANK_FMA). */
int chk_rank_counts(int algo, int dtype, int rank, int64_t b, const void* q, const void* qn,
                    const void* bh_vals, const void* target, const void* entity, const void* hn,
                    const void* bt, int64_t n_rows, int64_t shard_offset,
                    const int64_t* filter_indptr, const int64_t* filter_idx, int64_t filter_total,
                    const void* shadow, void* workspace, int64_t workspace_bytes,
                    int64_t* counts, void* stream);
int64_t chk_entity_shadow_bytes(int rank, int64_t n_rows);
/* Shadow of an entity shard for CHK_RANK_MMA: bf16 hi/lo operand blocks (pre-tiled for 1-D TMA bulk copies) plus the
 * fp32 per-entity epilogue inputs (||w||, Nyquist coefficient, 1/hn, bt).  entity / hn / bt are in `dtype`
 * (fp32 or fp64 models; bt may be NULL); rebuild it whenever one of them changes. */
int chk_entity_shadow_build(int dtype, int rank, int64_t n_rows, const void* entity, const void* hn, const void* bt,
                            void* shadow, void* stream);
int64_t chk_rank_mma_workspace_bytes(int rank, int64_t b);
/* Zero the workspace header (list length + sticky overflow flag); call once before a ranking pass. */
int chk_rank_mma_reset(void* workspace, void* stream);
/* Synchronises `stream` and reads the header back: length of the last re-check list, sticky overflow flag. */
int chk_rank_mma_status(const void* workspace, int64_t* last_list_len, int* overflowed, void* stream);
/* Measurement support: while armed (non-NULL cudaEvent_t handles, per calling thread), every CHK_RANK_MMA launch of
 * this thread records ev_start / ev_stop on its stream immediately around rank_mma_kernel (the dominant kernel), so
 * a benchmark can time that kernel alone inside a whole chk_rank_counts call.  Pass NULL, NULL to disarm. */
int chk_rank_mma_profile_events(void* ev_start, void* ev_stop);
/* Test support (like chk_score_all): the tensor-core tier's approximate scores [b,n_rows] and error bands
 * [b,n_rows] (band 0 = decided exactly in the clamp regime), plus its counts (no filter pass). */
int chk_score_all_mma(int dtype, int rank, int64_t b, const void* q, const void* qn, const void* bh_vals,
                      const void* target, const void* entity, const void* hn, const void* bt, int64_t n_rows,
                      const void* shadow, void* workspace, int64_t workspace_bytes, int64_t* counts,
                      void* scores, void* band, void* stream);

/* ---- evaluation batch in one call, filter index resident on the device ---------------------------------
 * The reference keeps filters[(entity, relation)] -> python list (datasets/process.py:55-77) and walks it per query on the host
 * (models/base.py:264-268).  Here the whole index lives in HBM as a sorted key table (code = entity * n_rel2 + relation) + CSR
 * (indptr [n_keys+1], vals: ids sorted and unique inside every list); chk_filter_lookup finds, per query of a batch, its list
 * (flt_start[i] = start in vals, flt_indptr = prefix sums of the list lengths), subtracts the true tail's own contribution
 * (counts[i] -= 1 when the tail is NOT in the list and lives in the shard [shard_offset, shard_offset + n_rows); list entries are
 * the filter pass's job) and sets bit 0 of *flags when a key is missing (the reference raises KeyError, base.py:266). */
int chk_filter_lookup(const int64_t* queries, int64_t b, int64_t n_rel2, const int64_t* keys, int64_t n_keys,
                      const int64_t* indptr, const int64_t* vals, int64_t shard_offset, int64_t n_rows,
                      int64_t* flt_indptr, int64_t* flt_start, int64_t* counts, int32_t* flags, void* stream);
/* One evaluation batch of KGModel.get_ranking (models/base.py:243-271) enqueued by ONE host call:
 * ids split -> chk_query_fwd -> chk_row_hnorm -> target scores from the full table -> chk_rank_counts tier `algo` over the shard
 * -> chk_filter_lookup -> filter pass.  counts [b] (zeroed here) ends as rank-1 of this shard's contribution (sum over shards,
 * then +1); target [b] holds the true tails' scores (NaN check, base.py:259-260). */
typedef struct chk_eval_args {
    int32_t algo, kind, dtype, rank, multi_c, pad_;
    int64_t b;
    const int64_t* queries;                               /* [b,3] (head, relation, tail) */
    const void *entity, *rel, *rel_diag, *ctx, *c_table, *bh, *bt;    /* the model's FULL tables (bh, bt NULL for bias 'none') */
    const void* hn_full; int64_t n_entities;              /* chk_row_hnorm of the full entity table */
    const void *shard_entity, *shard_hn, *shard_bt; int64_t shard_rows, shard_offset;   /* this rank's row range */
    const void* shadow; void* workspace; int64_t workspace_bytes;                       /* CHK_RANK_MMA only */
    const int64_t *f_keys, *f_indptr, *f_vals; int64_t f_nkeys, n_rel2;                 /* filter index */
    void* scratch; int64_t scratch_bytes;                 /* chk_eval_scratch_bytes(dtype, rank, b) */
    int64_t* counts; void* target; int32_t* flags;        /* outputs */
} chk_eval_args;
int64_t chk_eval_scratch_bytes(int dtype, int rank, int64_t b);
int chk_eval_batch(const chk_eval_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif
