// K2 (exact tier) — scores of a query batch against an entity shard, and fused filtered rank counts.
//
// Replaces the evaluation inner loop of KGModel.get_ranking (reference models/base.py:243-271):
// `score(q, candidates)` — which the reference evaluates by materialising z*conj(w) as a (b, N, r) complex
// temporary inside Distance.forward (utils/complexhyperbolic.py:222-237) —, the per-query Python filter
// loop (:264-268) and `sum(scores >= targets)` (:269-271).
//
// The Hermitian contraction is a register-tiled FMA GEMM: A = queries (complex, r), B = entity rows;
// each thread owns TQ x TE pairs with (Re, Im) accumulators and walks k in ascending order with the
// canonical chain dot_step(), so the tile kernel, the target score, the filter pass and the exact
// re-check of the tensor-core tier (chk_rank_mma.cu) produce bit-identical pair scores.  The epilogue
// (Hermitian form -> clamped x -> acosh -> -d^2 -> +bias -> compare / store) is fused; only (b) int64
// counters or the requested (b, n_rows) score matrix reach HBM.
//
// Filtering is subtractive: counts += #{e in shard: s >= target} - #{e in filter_i ∩ shard: s >= target}.
#include "chk_common.cuh"

namespace {

// ---- clamped Hermitian norm per row: one warp per row, fixed reduction tree -----------------------
template <typename T>
__global__ void row_hnorm_kernel(const T* __restrict__ table, int64_t n_rows, int r, T* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t wpb = blockDim.x >> 5;
    for (int64_t row = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); row < n_rows; row += (int64_t)gridDim.x * wpb) {
        const T* p = table + row * 2 * r;
        T s = T(0);
        for (int k = lane; k < r; k += 32) { s = Sc<T>::fma_(p[k], p[k], s); s = Sc<T>::fma_(p[r + k], p[r + k], s); }
        s = warp_sum<T>(s);
        if (lane == 0) out[row] = clamp_hnorm<T>(s);
    }
}

// ---- tile kernel ---------------------------------------------------------------------------------
template <typename T> struct TileCfg;
template <> struct TileCfg<float>  { static constexpr int TQ = 8, TE = 4, KC = 16; };
template <> struct TileCfg<double> { static constexpr int TQ = 4, TE = 4, KC = 16; };

template <typename T, int MODE>
__global__ void __launch_bounds__(256) rank_tile_kernel(RArgs<T> A) {
    using C = TileCfg<T>;
    constexpr int TQ = C::TQ, TE = C::TE, KC = C::KC;
    constexpr int BQ = 16 * TQ, BE = 16 * TE;
    constexpr int PADQ = BQ + 4, PADE = BE + 4;
    __shared__ __align__(16) T sQ[2][KC][PADQ];       // [plane][k][query]
    __shared__ __align__(16) T sE[2][KC][PADE];       // [plane][k][entity]
    __shared__ int sCnt[BQ];
    const int tid = threadIdx.x;
    const int tq = tid >> 4, te = tid & 15;           // 16 x 16 thread grid
    const int r = A.r;
    const int64_t q0 = (int64_t)blockIdx.x * BQ;
    const int64_t e0 = (int64_t)blockIdx.y * BE;

    static_assert(Chain<T>::BLK == 0 || Chain<T>::BLK == KC, "the K chunk is the canonical summation block");
    constexpr bool BLOCKED = Chain<T>::BLK != 0;
    T re[TQ][TE], im[TQ][TE];                      // running totals
#pragma unroll
    for (int i = 0; i < TQ; ++i)
#pragma unroll
        for (int j = 0; j < TE; ++j) { re[i][j] = T(0); im[i][j] = T(0); }
    if (MODE == 1) { for (int i = tid; i < BQ; i += 256) sCnt[i] = 0; }

    for (int k0 = 0; k0 < r; k0 += KC) {
        const int kc = min(KC, r - k0);
        __syncthreads();
        // cooperative, k-contiguous global reads; transposed smem writes
        for (int idx = tid; idx < BQ * KC; idx += 256) {
            int row = idx / KC, kk = idx - row * KC;
            int64_t gq = q0 + row;
            bool ok = (gq < A.b) && (kk < kc);
            const T* p = A.q + gq * 2 * r + k0 + kk;
            sQ[0][kk][row] = ok ? p[0] : T(0);
            sQ[1][kk][row] = ok ? p[r] : T(0);
        }
        for (int idx = tid; idx < BE * KC; idx += 256) {
            int row = idx / KC, kk = idx - row * KC;
            int64_t ge = e0 + row;
            bool ok = (ge < A.n_rows) && (kk < kc);
            const T* p = A.entity + ge * 2 * r + k0 + kk;
            sE[0][kk][row] = ok ? p[0] : T(0);
            sE[1][kk][row] = ok ? p[r] : T(0);
        }
        __syncthreads();
        // fp32: this chunk's partial sums start from zero and are added to the totals afterwards (blocked canonical
        // order); fp64: the chunk continues the single chain in the totals.
        T pre[BLOCKED ? TQ : 1][BLOCKED ? TE : 1], pim[BLOCKED ? TQ : 1][BLOCKED ? TE : 1];
        if (BLOCKED) {
#pragma unroll
            for (int i = 0; i < TQ; ++i)
#pragma unroll
                for (int j = 0; j < TE; ++j) { pre[BLOCKED ? i : 0][BLOCKED ? j : 0] = T(0); pim[BLOCKED ? i : 0][BLOCKED ? j : 0] = T(0); }
        }
        auto kstep = [&](int kk) {
            T zr[TQ], zi[TQ], wr[TE], wi[TE];
#pragma unroll
            for (int i = 0; i < TQ; ++i) { zr[i] = sQ[0][kk][tq * TQ + i]; zi[i] = sQ[1][kk][tq * TQ + i]; }
#pragma unroll
            for (int j = 0; j < TE; ++j) { wr[j] = sE[0][kk][te * TE + j]; wi[j] = sE[1][kk][te * TE + j]; }
#pragma unroll
            for (int i = 0; i < TQ; ++i)
#pragma unroll
                for (int j = 0; j < TE; ++j) {
                    if (BLOCKED) dot_step<T>(zr[i], zi[i], wr[j], wi[j], pre[BLOCKED ? i : 0][BLOCKED ? j : 0], pim[BLOCKED ? i : 0][BLOCKED ? j : 0]);
                    else dot_step<T>(zr[i], zi[i], wr[j], wi[j], re[i][j], im[i][j]);
                }
        };
        if (kc == KC) {
#pragma unroll
            for (int kk = 0; kk < KC; ++kk) kstep(kk);
        } else {
            for (int kk = 0; kk < kc; ++kk) kstep(kk);
        }
        if (BLOCKED) {
#pragma unroll
            for (int i = 0; i < TQ; ++i)
#pragma unroll
                for (int j = 0; j < TE; ++j) {
                    re[i][j] = Sc<T>::add_(re[i][j], pre[BLOCKED ? i : 0][BLOCKED ? j : 0]);
                    im[i][j] = Sc<T>::add_(im[i][j], pim[BLOCKED ? i : 0][BLOCKED ? j : 0]);
                }
        }
    }
    // ---- fused epilogue ----
    const bool has_bias = A.bt != nullptr;
    T wn[TE], btv[TE];
#pragma unroll
    for (int j = 0; j < TE; ++j) {
        int64_t ge = e0 + te * TE + j;
        bool ok = ge < A.n_rows;
        wn[j] = ok ? A.hn[ge] : T(-1);
        btv[j] = (ok && has_bias) ? A.bt[ge] : T(0);
    }
#pragma unroll
    for (int i = 0; i < TQ; ++i) {
        int64_t gq = q0 + tq * TQ + i;
        if (gq >= A.b) continue;
        const T zn = A.qn[gq];
        const T bh = has_bias ? A.bh_vals[gq] : T(0);
        int cnt = 0;
        const T tgt = (MODE == 1) ? A.target[gq] : T(0);
#pragma unroll
        for (int j = 0; j < TE; ++j) {
            int64_t ge = e0 + te * TE + j;
            if (ge >= A.n_rows) continue;
            T s = pair_score<T>(re[i][j], im[i][j], zn, wn[j], has_bias, bh, btv[j]);
            if (MODE == 0) A.scores[gq * A.n_rows + ge] = s;
            else cnt += (s >= tgt) ? 1 : 0;
        }
        if (MODE == 1 && cnt) atomicAdd(&sCnt[tq * TQ + i], cnt);
    }
    if (MODE == 1) {
        __syncthreads();
        for (int i = tid; i < BQ; i += 256) {
            int64_t gq = q0 + i;
            if (gq < A.b && sCnt[i]) atomicAdd(A.counts + gq, (unsigned long long)sCnt[i]);
        }
    }
}

// ---- per-pair exact kernels (canonical chain exact_pair(), one thread per pair) ---------------------

// target[i] = score(q_i, tail_rows_i): pair (i, i) of the gathered tail rows, one lane per query, rows staged by the warp
template <typename T>
__global__ void __launch_bounds__(32) target_kernel(RArgs<T> A, T* __restrict__ target) {
    __shared__ PairTiles<T> S;
    for (int64_t i0 = (int64_t)blockIdx.x * 32; i0 < A.b; i0 += (int64_t)gridDim.x * 32) {
        const int64_t i = i0 + threadIdx.x;
        const bool valid = i < A.b;
        const T s = warp_exact_pairs<T>(A, (unsigned)(valid ? i : 0), (unsigned)(valid ? i : 0), valid, S);
        if (valid) target[i] = s;
    }
}

// counts[i] -= #{ e in filter_i within the shard : score(i,e) >= target[i] }; one lane per filter entry, rows
// staged by the warp (warp_exact_pairs).
template <typename T>
__global__ void __launch_bounds__(32) filter_sub_kernel(RArgs<T> A, const int64_t* __restrict__ indptr,
                                                        const int64_t* __restrict__ fidx, int64_t shard_offset, int64_t total) {
    __shared__ PairTiles<T> S;
    for (int64_t t0 = (int64_t)blockIdx.x * 32; t0 < total; t0 += (int64_t)gridDim.x * 32) {
        const int64_t t = t0 + threadIdx.x;
        bool valid = t < total;
        int64_t i = 0, e = 0;
        if (valid) {
            // binary search the owning query: largest i with indptr[i] <= t
            int64_t lo = 0, hi = A.b;
            while (hi - lo > 1) { int64_t mid = (lo + hi) >> 1; if (indptr[mid] <= t) lo = mid; else hi = mid; }
            i = lo;
            e = fidx[t] - shard_offset;
            valid = e >= 0 && e < A.n_rows;
        }
        if (!__any_sync(CHK_FULL, valid)) continue;
        const T s = warp_exact_pairs<T>(A, (unsigned)i, (unsigned)e, valid, S);
        if (valid && s >= A.target[i]) atomicAdd(A.counts + i, (unsigned long long)(-1LL));
    }
}

template <typename T>
int run_tiles(const RArgs<T>& A, int mode, cudaStream_t st) {
    using C = TileCfg<T>;
    constexpr int BQ = 16 * C::TQ, BE = 16 * C::TE;
    int64_t gq = (A.b + BQ - 1) / BQ, ge = (A.n_rows + BE - 1) / BE;
    if (ge > 65535 * 32LL) { chk_set_error("shard too large for one launch"); return CHK_EUNSUPPORTED; }
    // grid.y is limited to 65535: walk entity tiles in slabs
    for (int64_t y0 = 0; y0 < ge; y0 += 65535) {
        int64_t ny = ge - y0 < 65535 ? ge - y0 : 65535;
        RArgs<T> S = A;
        S.entity = A.entity + y0 * BE * 2 * A.r;
        S.hn = A.hn + y0 * BE;
        if (A.bt) S.bt = A.bt + y0 * BE;
        S.n_rows = A.n_rows - y0 * BE;
        if (S.n_rows > ny * BE) S.n_rows = ny * BE;
        if (mode == 0) {
            // scores are addressed with the full row pitch: only a single slab is supported for MODE 0
            if (y0 != 0) { chk_set_error("chk_score_all: more than 65535 entity tiles"); return CHK_EUNSUPPORTED; }
            S.n_rows = A.n_rows;
            rank_tile_kernel<T, 0><<<dim3((unsigned)gq, (unsigned)ny), 256, 0, st>>>(S);
        } else {
            rank_tile_kernel<T, 1><<<dim3((unsigned)gq, (unsigned)ny), 256, 0, st>>>(S);
        }
        CHK_CUDA_LAUNCH_CHECK("rank_tile_kernel");
    }
    return CHK_OK;
}

}  // namespace

extern "C" int chk_row_hnorm(int dtype, int rank, int64_t n_rows, const void* table, void* hn, void* stream) {
    if (n_rows == 0) return CHK_OK;
    if (n_rows < 0 || rank < 2 || !table || !hn) { chk_set_error("chk_row_hnorm: bad argument"); return CHK_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    int64_t blocks = (n_rows + 7) / 8;
    if (blocks > 148 * 32) blocks = 148 * 32;
    if (dtype == CHK_F32) row_hnorm_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((const float*)table, n_rows, rank, (float*)hn);
    else if (dtype == CHK_F64) row_hnorm_kernel<double><<<(unsigned)blocks, 256, 0, st>>>((const double*)table, n_rows, rank, (double*)hn);
    else { chk_set_error("unknown dtype %d", dtype); return CHK_EINVAL; }
    CHK_CUDA_LAUNCH_CHECK("row_hnorm_kernel");
    return CHK_OK;
}

template <typename T>
static RArgs<T> make_rargs(int rank, int64_t b, const void* q, const void* qn, const void* bh_vals, const void* target,
                           const void* entity, const void* hn, const void* bt, int64_t n_rows) {
    RArgs<T> A{};
    A.q = (const T*)q; A.qn = (const T*)qn; A.bh_vals = (const T*)bh_vals; A.target = (const T*)target;
    A.entity = (const T*)entity; A.hn = (const T*)hn; A.bt = (const T*)bt; A.b = b; A.n_rows = n_rows; A.r = rank;
    return A;
}

extern "C" int chk_score_all(int dtype, int rank, int64_t b, const void* q, const void* qn, const void* bh_vals,
                             const void* entity, const void* hn, const void* bt, int64_t n_rows,
                             void* scores, void* stream) {
    if (b == 0 || n_rows == 0) return CHK_OK;
    if (b < 0 || n_rows < 0 || rank < 2 || !q || !qn || !entity || !hn || !scores || ((bh_vals == nullptr) != (bt == nullptr))) {
        chk_set_error("chk_score_all: bad argument"); return CHK_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CHK_F32) { auto A = make_rargs<float>(rank, b, q, qn, bh_vals, nullptr, entity, hn, bt, n_rows); A.scores = (float*)scores; return run_tiles<float>(A, 0, st); }
    if (dtype == CHK_F64) { auto A = make_rargs<double>(rank, b, q, qn, bh_vals, nullptr, entity, hn, bt, n_rows); A.scores = (double*)scores; return run_tiles<double>(A, 0, st); }
    chk_set_error("unknown dtype %d", dtype); return CHK_EINVAL;
}

extern "C" int chk_target_scores(int dtype, int rank, int64_t b, const void* q, const void* qn, const void* bh_vals,
                                 const void* tail_rows, const void* tail_hn, const void* tail_bt,
                                 void* target, void* stream) {
    if (b == 0) return CHK_OK;
    if (b < 0 || rank < 2 || !q || !qn || !tail_rows || !tail_hn || !target || ((bh_vals == nullptr) != (tail_bt == nullptr))) {
        chk_set_error("chk_target_scores: bad argument"); return CHK_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    unsigned blocks = (unsigned)((b + 31) / 32);
    if (blocks > 148 * 16) blocks = 148 * 16;
    // the gathered tail rows play the entity table: pair (i, i), hn = tail_hn, bt = tail_bt
    if (dtype == CHK_F32) { auto A = make_rargs<float>(rank, b, q, qn, bh_vals, nullptr, tail_rows, tail_hn, tail_bt, b); target_kernel<float><<<blocks, 32, 0, st>>>(A, (float*)target); }
    else if (dtype == CHK_F64) { auto A = make_rargs<double>(rank, b, q, qn, bh_vals, nullptr, tail_rows, tail_hn, tail_bt, b); target_kernel<double><<<blocks, 32, 0, st>>>(A, (double*)target); }
    else { chk_set_error("unknown dtype %d", dtype); return CHK_EINVAL; }
    CHK_CUDA_LAUNCH_CHECK("target_kernel");
    return CHK_OK;
}

// exact tier of chk_rank_counts; the tensor-core tier lives in chk_rank_mma.cu
int chk_rank_counts_fma(int dtype, int rank, int64_t b, const void* q, const void* qn, const void* bh_vals,
                        const void* target, const void* entity, const void* hn, const void* bt,
                        int64_t n_rows, int64_t* counts, cudaStream_t st) {
    if (dtype == CHK_F32) { auto A = make_rargs<float>(rank, b, q, qn, bh_vals, target, entity, hn, bt, n_rows); A.counts = (unsigned long long*)counts; return run_tiles<float>(A, 1, st); }
    if (dtype == CHK_F64) { auto A = make_rargs<double>(rank, b, q, qn, bh_vals, target, entity, hn, bt, n_rows); A.counts = (unsigned long long*)counts; return run_tiles<double>(A, 1, st); }
    chk_set_error("unknown dtype %d", dtype); return CHK_EINVAL;
}

int chk_filter_subtract(int dtype, int rank, int64_t b, const void* q, const void* qn, const void* bh_vals,
                        const void* target, const void* entity, const void* hn, const void* bt,
                        int64_t n_rows, int64_t shard_offset, const int64_t* indptr, const int64_t* fidx,
                        int64_t total, int64_t* counts, cudaStream_t st) {
    if (total <= 0) return CHK_OK;
    int64_t blocks = (total + 31) / 32;
    if (blocks > 148 * 32) blocks = 148 * 32;
    if (dtype == CHK_F32) { auto A = make_rargs<float>(rank, b, q, qn, bh_vals, target, entity, hn, bt, n_rows); A.counts = (unsigned long long*)counts; filter_sub_kernel<float><<<(unsigned)blocks, 32, 0, st>>>(A, indptr, fidx, shard_offset, total); }
    else if (dtype == CHK_F64) { auto A = make_rargs<double>(rank, b, q, qn, bh_vals, target, entity, hn, bt, n_rows); A.counts = (unsigned long long*)counts; filter_sub_kernel<double><<<(unsigned)blocks, 32, 0, st>>>(A, indptr, fidx, shard_offset, total); }
    else { chk_set_error("unknown dtype %d", dtype); return CHK_EINVAL; }
    CHK_CUDA_LAUNCH_CHECK("filter_sub_kernel");
    return CHK_OK;
}
