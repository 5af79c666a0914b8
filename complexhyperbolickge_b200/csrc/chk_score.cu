// K3 — scoring of gathered tails, forward and fused backward (training path), plus the row scatter-add.
//
// Replaces, for B queries x nt tails: KGModel.get_rhs gather (reference models/base.py:128-133),
// FFTUnitBall.similarity_score -> Distance.forward (models/complexhyperbolic.py:45-59,
// utils/complexhyperbolic.py:212-237, lift=True Hermitian form :176-178), the bias add
// (models/base.py:171) and Distance.backward/grad (utils/complexhyperbolic.py:192-210,239-254).
//
// One CTA per query row b; its warps split the nt tails.  Lane l owns complex coefficients
// k = l, l+32, ... of both rows, so the 2r-wide tail row is read with coalesced loads exactly once and
// nothing but the scores (fwd) / gradient rows (bwd) is written.  The backward recomputes the five pair
// scalars (re, im, zn, wn, x) instead of saving the reference's seven (b, nt, r) tensors.
#include "chk_common.cuh"

namespace {

constexpr int kWarps = 8;

template <typename T> struct SArgs {
    const T* q; int64_t q_stride_b, q_stride_j;
    const T* table; const int64_t* tail_idx; int64_t row_stride_b;
    const T* bh_vals; int64_t bh_stride_b, bh_stride_j; const T* bt;
    int64_t B, nt; int r;
    T* scores;                       // fwd
    const T* grad_scores; T* grad_q; T* grad_rows;   // bwd
    T* grad_dense;                   // bwd, optional: accumulate tail-row gradients straight into the dense table gradient
    // MODE 2 (training: forward + negative-sampling loss + backward in one pass, chk_score_gather_train)
    const int64_t* head_idx; int64_t head_stride_b, head_stride_j; const T* bh_table;   // bh value of pair (b,j) = bh_table[head_idx[..]]
    const double* hyper;             // device scalars: [2] = 1/(number of loss terms of the GLOBAL batch), [3] = valid rows of this batch
    T* loss_part;                    // [B] per-row loss partial (already scaled by hyper[2])
    T* gscores;                      // [B, nt] d loss / d score (feeds the bt gradient); g_bh[b] = sum_j of it
    T* g_bh;                         // [B] (one query per row) or NULL (per-pair queries: the bh gradient of a pair is gscores itself)
    T* pair_coef;                    // MODE 2, optional: [B*nt, 4] = (B1, B2, B3, 0) of every pair INSTEAD of its 2r-wide gradient row
                                     // (chk_reduce_apply rebuilds grad_w = B1 z + B2 (-i z) - B3 w from the query row and the tail row)
};

template <typename T> __device__ __forceinline__ void store4(T* p, T a, T b, T c, T d);
template <> __device__ __forceinline__ void store4<float>(float* p, float a, float b, float c, float d) {
    *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
template <> __device__ __forceinline__ void store4<double>(double* p, double a, double b, double c, double d) {
    reinterpret_cast<double2*>(p)[0] = make_double2(a, b); reinterpret_cast<double2*>(p)[1] = make_double2(c, d);
}

template <typename T>
__device__ __forceinline__ T logsigmoid_t(T x) {          // min(x,0) - log1p(exp(-|x|)), as ATen
    return Sc<T>::min_(x, T(0)) - Sc<T>::log1p_(Sc<T>::exp_(-Sc<T>::abs_(x)));
}

// A group of L = 2^LOGL lanes owns one (query, tail) pair: lane gl of the group holds complex coefficients
// k = gl, gl+L, ... (P per lane), so a warp works on 32/L pairs at once and the per-pair scalar section
// (Hermitian form, acosh, gradient coefficients) is amortised over them.  L = 8 at rank 33 (4 pairs per warp).
template <typename T, int LOGL, int P>
__device__ __forceinline__ void load_row(const T* __restrict__ row, int r, int gl, T (&re)[P], T (&im)[P]) {
#pragma unroll
    for (int i = 0; i < P; ++i) {
        const int k = gl + (i << LOGL);
        const bool ok = k < r;
        re[i] = ok ? row[k] : T(0);
        im[i] = ok ? row[r + k] : T(0);
    }
}
template <typename T, int LOGL>
__device__ __forceinline__ T group_sum(T v) {
#pragma unroll
    for (int o = (1 << LOGL) >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(CHK_FULL, v, o);
    return v;
}
template <typename T, int LOGL>
__device__ __forceinline__ void group_sum3(T& a, T& b, T& c) {
#pragma unroll
    for (int o = (1 << LOGL) >> 1; o > 0; o >>= 1) {
        T ta = __shfl_xor_sync(CHK_FULL, a, o), tb = __shfl_xor_sync(CHK_FULL, b, o), tc = __shfl_xor_sync(CHK_FULL, c, o);
        a += ta; b += tb; c += tc;
    }
}

// MODE 0: scores; MODE 1: adjoint from given d/dscores; MODE 2: training pass — scores, the negative-sampling loss terms
// -logsigmoid(+s) (column 0) / -logsigmoid(-s) (columns >= 1) of KGOptimizer.neg_sampling_loss (reference
// optimizers/kg_optimizer.py:115-122), their derivative and the adjoint, with every tail row gathered ONCE.
template <typename T, int LOGL, int P, int MODE>
__global__ void __launch_bounds__(kWarps * 32, (sizeof(T) == 4 && P <= 5) ? 4 : 1) score_gather_kernel(SArgs<T> A) {   // fp32, rank <= 129: 64 registers, so a
    // 500-row batch (4 CTAs x 148 SMs = 592 slots) is ONE wave instead of 1.13 (ncu r2: 80 registers, 3 CTAs per SM, a second wave of 56 CTAs)
    constexpr bool BWD = MODE >= 1, TRAIN = MODE == 2;
    constexpr int L = 1 << LOGL, G = 32 / L;          // lanes per pair, pairs per warp
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane & (L - 1), grp = lane >> LOGL;
    const int r = A.r;
    extern __shared__ unsigned char smem_raw[];
    T* red = reinterpret_cast<T*>(smem_raw);          // [kWarps][2r] for the grad_q reduction (BWD)
    __shared__ T red2[2][kWarps];                     // TRAIN: per-warp loss / bh-gradient partials
    T inv_total = T(0); int64_t n_valid = A.B;
    if (TRAIN) { inv_total = (T)A.hyper[2]; n_valid = (int64_t)A.hyper[3]; }
    for (int64_t b = blockIdx.x; b < A.B; b += gridDim.x) {
        T zr[P], zi[P], gzr[P], gzi[P];
        T zn = T(0);
        T loss_acc = T(0), gbh_acc = T(0);            // TRAIN: held by the gl == 0 lane of every pair group
        const bool per_pair_q = A.q_stride_j != 0;
        if (!per_pair_q) {
            load_row<T, LOGL, P>(A.q + b * A.q_stride_b * 2 * r, r, gl, zr, zi);
            T s = T(0);
#pragma unroll
            for (int i = 0; i < P; ++i) { s = Sc<T>::fma_(zr[i], zr[i], s); s = Sc<T>::fma_(zi[i], zi[i], s); }
            zn = clamp_hnorm<T>(group_sum<T, LOGL>(s));
        }
#pragma unroll
        for (int i = 0; i < P; ++i) { gzr[i] = T(0); gzi[i] = T(0); }
        // U pair-groups per warp iteration: their (independent) row gathers are all issued before any is reduced
        constexpr int U = P <= 5 ? 2 : 1;
        auto pair_body = [&](const int64_t j, const bool valid, const int64_t row, const T (&wr)[P], const T (&wi)[P]) {
            const int64_t pair = b * A.nt + j;
            if (per_pair_q) {
                load_row<T, LOGL, P>(A.q + (b * A.q_stride_b + j * A.q_stride_j) * 2 * r, r, gl, zr, zi);
                T s = T(0);
#pragma unroll
                for (int i = 0; i < P; ++i) { s = Sc<T>::fma_(zr[i], zr[i], s); s = Sc<T>::fma_(zi[i], zi[i], s); }
                zn = clamp_hnorm<T>(group_sum<T, LOGL>(s));
            }
            T re = T(0), im = T(0), ws = T(0);
#pragma unroll
            for (int i = 0; i < P; ++i) {
                dot_step<T>(zr[i], zi[i], wr[i], wi[i], re, im);
                ws = Sc<T>::fma_(wr[i], wr[i], ws); ws = Sc<T>::fma_(wi[i], wi[i], ws);
            }
            group_sum3<T, LOGL>(re, im, ws);
            const T wn = clamp_hnorm<T>(ws);
            const T x = clamped_x<T>(re, im, zn, wn);
            const T d = acosh_x<T>(x);
            if (!BWD) {
                if (gl == 0 && valid) {
                    T s = -Sc<T>::mul_(d, d);
                    A.scores[pair] = A.bt ? Sc<T>::add_(Sc<T>::add_(A.bh_vals[b * A.bh_stride_b + j * A.bh_stride_j], A.bt[row]), s) : s;
                }
            } else {
                T gsc;
                if (TRAIN) {
                    T s = -Sc<T>::mul_(d, d);
                    if (A.bt) s = Sc<T>::add_(Sc<T>::add_(A.bh_table[A.head_idx[b * A.head_stride_b + j * A.head_stride_j]], A.bt[row]), s);
                    const bool pos = j == 0;
                    const T xs = pos ? s : -s;                               // term = -logsigmoid(xs)
                    const bool live = valid && b < n_valid;                 // padding rows of a ragged batch carry no loss
                    const T gterm = -(T(1) / (T(1) + Sc<T>::exp_(xs))) * inv_total;   // d(-logsigmoid(xs))/dxs / total = -sigmoid(-xs)/total
                    gsc = live ? (pos ? gterm : -gterm) : T(0);
                    if (gl == 0 && valid) {
                        A.gscores[pair] = gsc;
                        if (live) { loss_acc -= logsigmoid_t<T>(xs) * inv_total; gbh_acc += gsc; }
                    }
                } else {
                    gsc = valid ? A.grad_scores[pair] : T(0);
                }
                const T gd = T(-2) * d * gsc;
                const T re1 = re - T(1);
                const T mod2 = Sc<T>::fma_(re1, re1, im * im);
                const T sq = Sc<T>::sqrt_(x * x - T(1));
                const T pz = Sc<T>::min_(sq * zn * zn * wn, -Sc<T>::ball_eps);
                const T pw = Sc<T>::min_(sq * wn * wn * zn, -Sc<T>::ball_eps);
                const T cz = T(4) * gd / pz, cw = T(4) * gd / pw;
                // grad_w = cw (wn (zw) z~ - |zw|^2 w), grad_z = cz (zn (zw) w~ - |zw|^2 z) with the pair scalars folded into six
                // coefficients: 12 flops per complex coefficient instead of 28, and no data-dependent branch inside the loops
                const T B1 = cw * wn * re1, B2 = cw * wn * im, B3 = cw * mod2;
                const T A1 = cz * zn * re1, A2 = cz * zn * im, A3 = cz * mod2;
                if (TRAIN && A.pair_coef) {
                    if (valid && gl == 0) {
                        store4<T>(A.pair_coef + pair * 4, B1, B2, B3, T(0));
                    }
                } else if (valid) {
                    if (A.grad_dense == nullptr) {
                        T* grow = A.grad_rows + pair * 2 * r;
#pragma unroll
                        for (int i = 0; i < P; ++i) {
                            const int k = gl + (i << LOGL);
                            if (k < r) {
                                grow[k] = Sc<T>::fma_(B1, zr[i], Sc<T>::fma_(B2, zi[i], -B3 * wr[i]));
                                grow[r + k] = Sc<T>::fma_(B1, zi[i], Sc<T>::fma_(-B2, zr[i], -B3 * wi[i]));
                            }
                        }
                    } else {
                        T* grow = A.grad_dense + row * 2 * r;
#pragma unroll
                        for (int i = 0; i < P; ++i) {
                            const int k = gl + (i << LOGL);
                            if (k < r) {
                                atomicAdd(grow + k, Sc<T>::fma_(B1, zr[i], Sc<T>::fma_(B2, zi[i], -B3 * wr[i])));
                                atomicAdd(grow + r + k, Sc<T>::fma_(B1, zi[i], Sc<T>::fma_(-B2, zr[i], -B3 * wi[i])));
                            }
                        }
                    }
                }
                if (per_pair_q) {
                    if (valid) {
                        T* gqrow = A.grad_q + (b * A.q_stride_b + j * A.q_stride_j) * 2 * r;
#pragma unroll
                        for (int i = 0; i < P; ++i) {
                            const int k = gl + (i << LOGL);
                            if (k < r) {
                                gqrow[k] = Sc<T>::fma_(A1, wr[i], Sc<T>::fma_(-A2, wi[i], -A3 * zr[i]));
                                gqrow[r + k] = Sc<T>::fma_(A1, wi[i], Sc<T>::fma_(A2, wr[i], -A3 * zi[i]));
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < P; ++i) {                  // gd == 0 for padding groups: they add zeros
                        gzr[i] += Sc<T>::fma_(A1, wr[i], Sc<T>::fma_(-A2, wi[i], -A3 * zr[i]));
                        gzi[i] += Sc<T>::fma_(A1, wi[i], Sc<T>::fma_(A2, wr[i], -A3 * zi[i]));
                    }
                }
            }
        };
        for (int64_t j0 = (int64_t)warp * G; j0 < A.nt; j0 += (int64_t)kWarps * G * U) {
            int64_t rows[U];
            T wr[U][P], wi[U][P];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t j = j0 + (int64_t)u * kWarps * G + grp;
                const int64_t jj = j < A.nt ? j : 0;
                rows[u] = A.tail_idx ? A.tail_idx[b * A.nt + jj] : (b * A.row_stride_b + jj);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) load_row<T, LOGL, P>(A.table + rows[u] * 2 * r, r, gl, wr[u], wi[u]);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t jb = j0 + (int64_t)u * kWarps * G;           // warp-uniform: every lane runs the shuffles
                if (jb < A.nt) {
                    const int64_t j = jb + grp;
                    pair_body(j < A.nt ? j : 0, j < A.nt, rows[u], wr[u], wi[u]);
                }
            }
        }
        if (TRAIN) {
            // fixed-order block reduction of the row's loss terms and bh gradient (bit-reproducible: no atomics)
            loss_acc = warp_sum<T>(loss_acc); gbh_acc = warp_sum<T>(gbh_acc);
            __syncthreads();
            if (lane == 0) { red2[0][warp] = loss_acc; red2[1][warp] = gbh_acc; }
            __syncthreads();
            if (threadIdx.x == 0) {
                T l = T(0), g = T(0);
#pragma unroll
                for (int w = 0; w < kWarps; ++w) { l += red2[0][w]; g += red2[1][w]; }
                A.loss_part[b] = l;
                if (A.g_bh) A.g_bh[b] = g;
            }
        }
        if (BWD && !per_pair_q) {
            // sum the per-lane partial gradients over the warp's pair groups, then over the warps
#pragma unroll
            for (int i = 0; i < P; ++i)
#pragma unroll
                for (int o = L; o < 32; o <<= 1) {
                    gzr[i] += __shfl_xor_sync(CHK_FULL, gzr[i], o);
                    gzi[i] += __shfl_xor_sync(CHK_FULL, gzi[i], o);
                }
            __syncthreads();
            if (grp == 0) {
#pragma unroll
                for (int i = 0; i < P; ++i) {
                    const int k = gl + (i << LOGL);
                    if (k < r) { red[warp * 2 * r + k] = gzr[i]; red[warp * 2 * r + r + k] = gzi[i]; }
                }
            }
            __syncthreads();
            T* out = A.grad_q + b * A.q_stride_b * 2 * r;
            for (int c = threadIdx.x; c < 2 * r; c += blockDim.x) {
                T s = T(0);
#pragma unroll
                for (int w = 0; w < kWarps; ++w) s += red[w * 2 * r + c];
                out[c] = s;
            }
        }
    }
}

template <typename T, int MODE>
int launch_gather(const SArgs<T>& A, cudaStream_t st) {
    constexpr bool BWD = MODE >= 1;
    int64_t blocks = A.B < 148 * 32 ? A.B : 148 * 32;
    size_t smem = BWD ? (size_t)kWarps * 2 * A.r * sizeof(T) : 0;
#define CHK_LAUNCH(LOGL, P)                                                                         \
    score_gather_kernel<T, LOGL, P, MODE><<<(unsigned)blocks, kWarps * 32, smem, st>>>(A)
    const int r = A.r;
    if (r <= 10) CHK_LAUNCH(1, 5);            // rank 9:   2 lanes x 5
    else if (r <= 20) CHK_LAUNCH(2, 5);       // rank 17:  4 lanes x 5
    else if (r <= 40) CHK_LAUNCH(3, 5);       // rank 33:  8 lanes x 5 (4 pairs per warp)
    else if (r <= 80) CHK_LAUNCH(4, 5);       // rank 65: 16 lanes x 5
    else if (r <= 160) CHK_LAUNCH(5, 5);      // rank 129
    else if (r <= 288) CHK_LAUNCH(5, 9);      // rank 257
    else { chk_set_error("rank %d too large", A.r); return CHK_EUNSUPPORTED; }
#undef CHK_LAUNCH
    CHK_CUDA_LAUNCH_CHECK("score_gather_kernel");
    return CHK_OK;
}

template <typename T>
__global__ void scatter_add_rows_kernel(T* __restrict__ dense, const int64_t* __restrict__ idx,
                                        const T* __restrict__ rows, int64_t n_rows, int64_t width) {
    const int64_t total = n_rows * width;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t rrow = i / width, c = i - rrow * width;
        atomicAdd(dense + idx[rrow] * width + c, rows[i]);
    }
}

}  // namespace

extern "C" int chk_score_gather_fwd(int dtype, int rank, int64_t B, int64_t nt,
                                    const void* q, int64_t q_stride_b, int64_t q_stride_j,
                                    const void* table, const int64_t* tail_idx, int64_t row_stride_b,
                                    const void* bh_vals, int64_t bh_stride_b, int64_t bh_stride_j, const void* bt,
                                    void* scores, void* stream) {
    if (B == 0 || nt == 0) return CHK_OK;
    if (B < 0 || nt < 0 || rank < 2 || !q || !table || !scores) { chk_set_error("chk_score_gather_fwd: bad argument"); return CHK_EINVAL; }
    if ((bh_vals == nullptr) != (bt == nullptr)) { chk_set_error("bh_vals and bt must both be given or both be NULL"); return CHK_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CHK_F32) {
        SArgs<float> A{(const float*)q, q_stride_b, q_stride_j, (const float*)table, tail_idx, row_stride_b,
                       (const float*)bh_vals, bh_stride_b, bh_stride_j, (const float*)bt, B, nt, rank, (float*)scores, nullptr, nullptr, nullptr, nullptr};
        return launch_gather<float, 0>(A, st);
    } else if (dtype == CHK_F64) {
        SArgs<double> A{(const double*)q, q_stride_b, q_stride_j, (const double*)table, tail_idx, row_stride_b,
                        (const double*)bh_vals, bh_stride_b, bh_stride_j, (const double*)bt, B, nt, rank, (double*)scores, nullptr, nullptr, nullptr, nullptr};
        return launch_gather<double, 0>(A, st);
    }
    chk_set_error("unknown dtype %d", dtype); return CHK_EINVAL;
}

static int score_gather_bwd_impl(int dtype, int rank, int64_t B, int64_t nt,
                                 const void* q, int64_t q_stride_b, int64_t q_stride_j,
                                 const void* table, const int64_t* tail_idx, int64_t row_stride_b,
                                 const void* grad_scores, void* grad_q, void* grad_rows, void* grad_dense, void* stream) {
    if (B == 0 || nt == 0) return CHK_OK;
    if (B < 0 || nt < 0 || rank < 2 || !q || !table || !grad_scores || !grad_q || (!grad_rows && !grad_dense) ||
        (grad_dense && !tail_idx)) {
        chk_set_error("chk_score_gather_bwd: bad argument"); return CHK_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CHK_F32) {
        SArgs<float> A{(const float*)q, q_stride_b, q_stride_j, (const float*)table, tail_idx, row_stride_b,
                       nullptr, 0, 0, nullptr, B, nt, rank, nullptr, (const float*)grad_scores, (float*)grad_q, (float*)grad_rows, (float*)grad_dense};
        return launch_gather<float, 1>(A, st);
    } else if (dtype == CHK_F64) {
        SArgs<double> A{(const double*)q, q_stride_b, q_stride_j, (const double*)table, tail_idx, row_stride_b,
                        nullptr, 0, 0, nullptr, B, nt, rank, nullptr, (const double*)grad_scores, (double*)grad_q, (double*)grad_rows, (double*)grad_dense};
        return launch_gather<double, 1>(A, st);
    }
    chk_set_error("unknown dtype %d", dtype); return CHK_EINVAL;
}

extern "C" int chk_score_gather_bwd(int dtype, int rank, int64_t B, int64_t nt,
                                    const void* q, int64_t q_stride_b, int64_t q_stride_j,
                                    const void* table, const int64_t* tail_idx, int64_t row_stride_b,
                                    const void* grad_scores, void* grad_q, void* grad_rows, void* stream) {
    return score_gather_bwd_impl(dtype, rank, B, nt, q, q_stride_b, q_stride_j, table, tail_idx, row_stride_b, grad_scores,
                                 grad_q, grad_rows, nullptr, stream);
}

extern "C" int chk_score_gather_bwd_scatter(int dtype, int rank, int64_t B, int64_t nt,
                                            const void* q, int64_t q_stride_b, int64_t q_stride_j,
                                            const void* table, const int64_t* tail_idx,
                                            const void* grad_scores, void* grad_q, void* grad_table_dense, void* stream) {
    return score_gather_bwd_impl(dtype, rank, B, nt, q, q_stride_b, q_stride_j, table, tail_idx, 0, grad_scores,
                                 grad_q, nullptr, grad_table_dense, stream);
}

extern "C" int chk_scatter_add_rows(int dtype, void* dense, const int64_t* idx, const void* rows,
                                    int64_t n_rows, int64_t width, void* stream) {
    if (n_rows == 0 || width == 0) return CHK_OK;
    if (n_rows < 0 || width < 0 || !dense || !idx || !rows) { chk_set_error("chk_scatter_add_rows: bad argument"); return CHK_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    int64_t total = n_rows * width;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (dtype == CHK_F32) scatter_add_rows_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((float*)dense, idx, (const float*)rows, n_rows, width);
    else if (dtype == CHK_F64) scatter_add_rows_kernel<double><<<(unsigned)blocks, 256, 0, st>>>((double*)dense, idx, (const double*)rows, n_rows, width);
    else { chk_set_error("unknown dtype %d", dtype); return CHK_EINVAL; }
    CHK_CUDA_LAUNCH_CHECK("scatter_add_rows_kernel");
    return CHK_OK;
}

// Training pass (MODE 2): replaces, for one batch, the two model() calls of KGOptimizer.neg_sampling_loss plus
// F.logsigmoid / mean and their autograd backward (reference optimizers/kg_optimizer.py:101-123) — see the header.
template <typename T>
static int score_gather_train_t(int rank, int64_t B, int64_t nt, const void* q, int64_t q_stride_b, int64_t q_stride_j,
                                const void* table, const int64_t* tail_idx, const int64_t* head_idx, int64_t head_stride_b,
                                int64_t head_stride_j, const void* bh, const void* bt, const double* hyper, void* loss_part,
                                void* grad_scores, void* grad_q, void* grad_rows, void* pair_coef, void* g_bh, cudaStream_t st) {
    SArgs<T> A{};
    A.q = (const T*)q; A.q_stride_b = q_stride_b; A.q_stride_j = q_stride_j;
    A.table = (const T*)table; A.tail_idx = tail_idx; A.row_stride_b = 0;
    A.bt = (const T*)bt; A.B = B; A.nt = nt; A.r = rank;
    A.grad_q = (T*)grad_q; A.grad_rows = (T*)grad_rows; A.grad_dense = nullptr;
    A.head_idx = head_idx; A.head_stride_b = head_stride_b; A.head_stride_j = head_stride_j; A.bh_table = (const T*)bh;
    A.hyper = hyper; A.loss_part = (T*)loss_part; A.gscores = (T*)grad_scores; A.g_bh = (T*)g_bh; A.pair_coef = (T*)pair_coef;
    return launch_gather<T, 2>(A, st);
}

extern "C" int chk_score_gather_train(int dtype, int rank, int64_t B, int64_t nt,
                                      const void* q, int64_t q_stride_b, int64_t q_stride_j,
                                      const void* table, const int64_t* tail_idx,
                                      const int64_t* head_idx, int64_t head_stride_b, int64_t head_stride_j,
                                      const void* bh, const void* bt, const double* hyper,
                                      void* loss_part, void* grad_scores, void* grad_q, void* grad_rows, void* pair_coef, void* g_bh,
                                      void* stream) {
    if (B == 0 || nt == 0) return CHK_OK;
    if (B < 0 || nt < 0 || rank < 2 || !q || !table || !tail_idx || !hyper || !loss_part || !grad_scores || !grad_q || (!grad_rows && !pair_coef) ||
        ((bh == nullptr) != (bt == nullptr)) || (bh && !head_idx)) {
        chk_set_error("chk_score_gather_train: bad argument"); return CHK_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CHK_F32) return score_gather_train_t<float>(rank, B, nt, q, q_stride_b, q_stride_j, table, tail_idx, head_idx, head_stride_b, head_stride_j, bh, bt, hyper, loss_part, grad_scores, grad_q, grad_rows, pair_coef, g_bh, st);
    if (dtype == CHK_F64) return score_gather_train_t<double>(rank, B, nt, q, q_stride_b, q_stride_j, table, tail_idx, head_idx, head_stride_b, head_stride_j, bh, bt, hyper, loss_part, grad_scores, grad_q, grad_rows, pair_coef, g_bh, st);
    chk_set_error("unknown dtype %d", dtype); return CHK_EINVAL;
}
