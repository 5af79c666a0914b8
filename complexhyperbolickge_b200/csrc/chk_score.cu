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
    const T* const* peer_tab;        // MODE 2, optional (owner-sharded tables, data parallel): row `i` of the entity table is read from
    const T* const* peer_bt;         // peer_tab[i / rows_per_owner] (the owner's copy, peer memory over NVLink), bt likewise; NULL = local table
    int64_t rows_per_owner; int peer_n;   // peer_n = world (<= 32: the base pointers are staged in shared memory)
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

// asynchronous global -> shared copies (the row ring of the rank-257 training pass)
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
constexpr int kRingDepth = 4;                         // rows in flight per warp (rank-257 fp32 training pass)
__host__ __device__ inline int ring_slot_floats(int r) { return ((2 * r + 3) & ~3) + 4; }   // row | int64 row id | bt | pad (16-byte multiple)

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
// RC: the rank as a compile-time constant (0 = runtime A.r).  The training pass is instantiated for the BASELINE ranks 33 and 257:
// ncu r2 showed a third of its warp instructions were integer address / predicate work (IADD3, IMAD, LEA, ISETP) that
// constant row offsets and `k < r` tests fold away.
template <typename T, int LOGL, int P, int MODE, int RC = 0>
__global__ void __launch_bounds__(kWarps * 32, (sizeof(T) == 4 && P <= 5) ? 4 : 1) score_gather_kernel(SArgs<T> A) {   // fp32, rank <= 129: 64 registers, so a
    // 500-row batch (4 CTAs x 148 SMs = 592 slots) is ONE wave instead of 1.13 (ncu r2: 80 registers, 3 CTAs per SM, a second wave of 56 CTAs)
    constexpr bool BWD = MODE >= 1, TRAIN = MODE >= 2;          // MODE 3: the training pass on owner-sharded (peer) tables
    constexpr int L = 1 << LOGL, G = 32 / L;          // lanes per pair, pairs per warp
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gl = lane & (L - 1), grp = lane >> LOGL;
    const int r = RC > 0 ? RC : A.r;
    extern __shared__ unsigned char smem_raw[];
    T* red = reinterpret_cast<T*>(smem_raw);          // [kWarps][2r] for the grad_q reduction (BWD)
    __shared__ T red2[2][kWarps];                     // TRAIN: per-warp loss / bh-gradient partials
    __shared__ const T* s_tab[32];
    __shared__ const T* s_bt[32];
    T inv_total = T(0); int64_t n_valid = A.B;
    if (TRAIN) {
        inv_total = (T)A.hyper[2]; n_valid = (int64_t)A.hyper[3];
        if (A.peer_tab) {                             // owner-sharded tables: a row is read from its owner's copy
            if ((int)threadIdx.x < A.peer_n) { s_tab[threadIdx.x] = A.peer_tab[threadIdx.x]; s_bt[threadIdx.x] = A.peer_bt ? A.peer_bt[threadIdx.x] : nullptr; }
            __syncthreads();
        }
    }
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
        auto pair_body = [&](const int64_t j, const bool valid, const int64_t row, const T (&wr)[P], const T (&wi)[P], const T btv) {
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
                    if (A.bt) s = Sc<T>::add_(Sc<T>::add_(A.bh_table[A.head_idx[b * A.head_stride_b + j * A.head_stride_j]], btv), s);
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
        // local rows: the direct loads of the other branch are as fast (r2: 268 vs 276 us/step), so the ring is MODE 3 only
        constexpr bool RING = MODE == 3 && P == 9 && sizeof(T) == 4 && LOGL == 5;
        if constexpr (RING) {
            // Rank 257, fp32: one row per warp iteration would leave a warp's 13 rows strictly serial (row latency + pair
            // scalars each; the latency is a peer access on owner-sharded tables).  Rows are therefore streamed through a
            // per-warp ring in shared memory with 8-byte cp.async copies, kRingDepth - 1 rows ahead of the one being scored
            // (rows are 8-byte, not 16-byte, aligned: 2r = 514 floats, so no bulk/TMA copy); the row id and its bt value
            // ride in the slot's tail.
            const int slotf = ring_slot_floats(r);
            float* ring = reinterpret_cast<float*>(red) + kWarps * 2 * r + warp * kRingDepth * slotf;
            const int nrows = warp < A.nt ? (int)((A.nt - warp + kWarps - 1) / kWarps) : 0;
            auto issue = [&](const int i) {
                if (i < nrows) {
                    const int64_t row = A.tail_idx[b * A.nt + warp + (int64_t)i * kWarps];
                    const float* tab = reinterpret_cast<const float*>(A.table);
                    const float* btab = reinterpret_cast<const float*>(A.bt);
                    if (A.peer_tab) {
                        const unsigned o = (unsigned)row / (unsigned)A.rows_per_owner;
                        tab = reinterpret_cast<const float*>(s_tab[o]); btab = reinterpret_cast<const float*>(s_bt[o]);
                    }
                    float* dst = ring + (i % kRingDepth) * slotf;
                    const float* src = tab + row * 2 * r;
                    for (int c = lane; c < r; c += 32) cp_async8(dst + 2 * c, src + 2 * c);
                    if (lane == 0) {
                        *reinterpret_cast<int64_t*>(dst + slotf - 4) = row;
                        if (A.bt) cp_async4(dst + slotf - 2, btab + row);
                    }
                }
                cp_async_commit();
            };
#pragma unroll
            for (int i = 0; i < kRingDepth - 1; ++i) issue(i);
            for (int i = 0; i < nrows; ++i) {
                issue(i + kRingDepth - 1);
                cp_async_wait<kRingDepth - 1>();
                __syncwarp();
                const float* slot = ring + (i % kRingDepth) * slotf;
                T wr[P], wi[P];
#pragma unroll
                for (int q = 0; q < P; ++q) {
                    const int k = gl + (q << LOGL);
                    const bool ok = k < r;
                    wr[q] = ok ? (T)slot[k] : T(0);
                    wi[q] = ok ? (T)slot[r + k] : T(0);
                }
                const int64_t row = *reinterpret_cast<const int64_t*>(slot + slotf - 4);
                const T btv = A.bt ? (T)slot[slotf - 2] : T(0);
                pair_body(warp + (int64_t)i * kWarps, true, row, wr, wi, btv);
                __syncwarp();                                              // the slot is refilled two iterations from now
            }
        } else {
            // The tail ids of the NEXT iteration are fetched while this one computes, and a pair's bt value travels with its row:
            // the only exposed latency per iteration is the one of the rows themselves (it is a peer access on owner-sharded tables).
            constexpr int64_t STEP = (int64_t)kWarps * G * U;
            auto row_of = [&](const int64_t jb, const int u) -> int64_t {
                const int64_t j = jb + (int64_t)u * kWarps * G + grp;
                const int64_t jj = j < A.nt ? j : 0;
                return A.tail_idx ? A.tail_idx[b * A.nt + jj] : (b * A.row_stride_b + jj);
            };
            int64_t rows[U];
    #pragma unroll
            for (int u = 0; u < U; ++u) rows[u] = row_of((int64_t)warp * G, u);
            for (int64_t j0 = (int64_t)warp * G; j0 < A.nt; j0 += STEP) {
                T wr[U][P], wi[U][P], btv[U];
    #pragma unroll
                for (int u = 0; u < U; ++u) {
                    const T* tab = A.table;
                    btv[u] = T(0);
                    if (TRAIN) {
                        const T* btab = A.bt;
                        if (A.peer_tab) { const unsigned o = (unsigned)rows[u] / (unsigned)A.rows_per_owner; tab = s_tab[o]; btab = s_bt[o]; }
                        if (A.bt) btv[u] = btab[rows[u]];
                    }
                    load_row<T, LOGL, P>(tab + rows[u] * 2 * r, r, gl, wr[u], wi[u]);
                }
                int64_t nxt[U];
    #pragma unroll
                for (int u = 0; u < U; ++u) nxt[u] = j0 + STEP < A.nt ? row_of(j0 + STEP, u) : 0;
    #pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int64_t jb = j0 + (int64_t)u * kWarps * G;           // warp-uniform: every lane runs the shuffles
                    if (jb < A.nt) {
                        const int64_t j = jb + grp;
                        pair_body(j < A.nt ? j : 0, j < A.nt, rows[u], wr[u], wi[u], btv[u]);
                    }
                }
    #pragma unroll
                for (int u = 0; u < U; ++u) rows[u] = nxt[u];
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
    if (MODE == 3 && sizeof(T) == 4 && A.r > 160) {                       // row ring of the rank-257 fp32 training pass on peer tables
        smem += (size_t)kWarps * kRingDepth * ring_slot_floats(A.r) * sizeof(float);
        static bool attr_set = false;
        if (!attr_set) {
            if (cudaFuncSetAttribute((const void*)score_gather_kernel<T, 5, 9, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess ||
                cudaFuncSetAttribute((const void*)score_gather_kernel<T, 5, 9, MODE, 257>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess) {
                chk_set_error("score_gather_kernel: cannot raise the dynamic shared memory limit"); return CHK_ECUDA;
            }
            attr_set = true;
        }
    }
#define CHK_LAUNCH(LOGL, P)                                                                         \
    score_gather_kernel<T, LOGL, P, MODE><<<(unsigned)blocks, kWarps * 32, smem, st>>>(A)
    const int r = A.r;
    if (MODE >= 2 && r == 33) score_gather_kernel<T, 3, 5, MODE, 33><<<(unsigned)blocks, kWarps * 32, smem, st>>>(A);
    else if (MODE >= 2 && r == 257) score_gather_kernel<T, 5, 9, MODE, 257><<<(unsigned)blocks, kWarps * 32, smem, st>>>(A);
    else if (r <= 10) CHK_LAUNCH(1, 5);       // rank 9:   2 lanes x 5
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
                                void* grad_scores, void* grad_q, void* grad_rows, void* pair_coef, void* g_bh,
                                const void* const* peer_tab, const void* const* peer_bt, int64_t rows_per_owner, int peer_n, cudaStream_t st) {
    SArgs<T> A{};
    A.q = (const T*)q; A.q_stride_b = q_stride_b; A.q_stride_j = q_stride_j;
    A.table = (const T*)table; A.tail_idx = tail_idx; A.row_stride_b = 0;
    A.bt = (const T*)bt; A.B = B; A.nt = nt; A.r = rank;
    A.grad_q = (T*)grad_q; A.grad_rows = (T*)grad_rows; A.grad_dense = nullptr;
    A.head_idx = head_idx; A.head_stride_b = head_stride_b; A.head_stride_j = head_stride_j; A.bh_table = (const T*)bh;
    A.hyper = hyper; A.loss_part = (T*)loss_part; A.gscores = (T*)grad_scores; A.g_bh = (T*)g_bh; A.pair_coef = (T*)pair_coef;
    A.peer_tab = (const T* const*)peer_tab; A.peer_bt = (const T* const*)peer_bt; A.rows_per_owner = rows_per_owner > 0 ? rows_per_owner : 1; A.peer_n = peer_n;
    if (peer_tab) return launch_gather<T, 3>(A, st);
    return launch_gather<T, 2>(A, st);
}

static int score_gather_train_impl(int dtype, int rank, int64_t B, int64_t nt,
                                   const void* q, int64_t q_stride_b, int64_t q_stride_j,
                                   const void* table, const int64_t* tail_idx,
                                   const int64_t* head_idx, int64_t head_stride_b, int64_t head_stride_j,
                                   const void* bh, const void* bt, const double* hyper,
                                   void* loss_part, void* grad_scores, void* grad_q, void* grad_rows, void* pair_coef, void* g_bh,
                                   const void* const* peer_tab, const void* const* peer_bt, int64_t rows_per_owner, int peer_n, void* stream) {
    if (B == 0 || nt == 0) return CHK_OK;
    if (B < 0 || nt < 0 || rank < 2 || !q || !table || !tail_idx || !hyper || !loss_part || !grad_scores || !grad_q || (!grad_rows && !pair_coef) ||
        ((bh == nullptr) != (bt == nullptr)) || (bh && !head_idx)) {
        chk_set_error("chk_score_gather_train: bad argument"); return CHK_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CHK_F32) return score_gather_train_t<float>(rank, B, nt, q, q_stride_b, q_stride_j, table, tail_idx, head_idx, head_stride_b, head_stride_j, bh, bt, hyper, loss_part, grad_scores, grad_q, grad_rows, pair_coef, g_bh, peer_tab, peer_bt, rows_per_owner, peer_n, st);
    if (dtype == CHK_F64) return score_gather_train_t<double>(rank, B, nt, q, q_stride_b, q_stride_j, table, tail_idx, head_idx, head_stride_b, head_stride_j, bh, bt, hyper, loss_part, grad_scores, grad_q, grad_rows, pair_coef, g_bh, peer_tab, peer_bt, rows_per_owner, peer_n, st);
    chk_set_error("unknown dtype %d", dtype); return CHK_EINVAL;
}

extern "C" int chk_score_gather_train(int dtype, int rank, int64_t B, int64_t nt,
                                      const void* q, int64_t q_stride_b, int64_t q_stride_j,
                                      const void* table, const int64_t* tail_idx,
                                      const int64_t* head_idx, int64_t head_stride_b, int64_t head_stride_j,
                                      const void* bh, const void* bt, const double* hyper,
                                      void* loss_part, void* grad_scores, void* grad_q, void* grad_rows, void* pair_coef, void* g_bh,
                                      void* stream) {
    return score_gather_train_impl(dtype, rank, B, nt, q, q_stride_b, q_stride_j, table, tail_idx, head_idx, head_stride_b, head_stride_j,
                                   bh, bt, hyper, loss_part, grad_scores, grad_q, grad_rows, pair_coef, g_bh, nullptr, nullptr, 0, 0, stream);
}

// Owner-sharded variant (data parallel, tables in peer-accessible memory): tail rows and bt values are read from the copy of
// the rank that owns the row, peer_tables[row / rows_per_owner] (all copies have the full-table layout, only the owner's rows
// are current); everything else as chk_score_gather_train.
extern "C" int chk_score_gather_train_peer(int dtype, int rank, int64_t B, int64_t nt,
                                           const void* q, int64_t q_stride_b, int64_t q_stride_j,
                                           const void* const* peer_tables, const void* const* peer_bt, int64_t rows_per_owner, int world,
                                           const int64_t* tail_idx,
                                           const int64_t* head_idx, int64_t head_stride_b, int64_t head_stride_j,
                                           const void* bh, const double* hyper,
                                           void* loss_part, void* grad_scores, void* grad_q, void* grad_rows, void* pair_coef, void* g_bh,
                                           void* stream) {
    if (!peer_tables || rows_per_owner < 1 || rows_per_owner > 0x7fffffff || world < 1 || world > 32 || ((bh == nullptr) != (peer_bt == nullptr))) { chk_set_error("chk_score_gather_train_peer: bad argument"); return CHK_EINVAL; }
    // `table` / `bt` only take part in the argument checks of the shared implementation
    return score_gather_train_impl(dtype, rank, B, nt, q, q_stride_b, q_stride_j, peer_tables, tail_idx, head_idx, head_stride_b, head_stride_j,
                                   bh, peer_bt, hyper, loss_part, grad_scores, grad_q, grad_rows, pair_coef, g_bh, peer_tables, peer_bt,
                                   rows_per_owner, world, stream);
}

// out[i, :] = peer_tables[ids[i] / rows_per_owner][ids[i] * width + :]  (head rows / head biases of an owner-sharded table)
namespace {
template <typename T>
__global__ void __launch_bounds__(256) peer_gather_rows_kernel(const T* const* __restrict__ peers, int64_t rows_per_owner, const int64_t* __restrict__ ids,
                                                               int64_t n, int64_t width, T* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    if (width == 1) {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
            const int64_t id = ids[i];
            out[i] = peers[id / rows_per_owner][id];
        }
        return;
    }
    for (int64_t i = warp; i < n; i += nwarps) {
        const int64_t id = ids[i];
        const T* src = peers[id / rows_per_owner] + id * width;
        for (int64_t c = lane; c < width; c += 32) out[i * width + c] = src[c];
    }
}
}  // namespace

extern "C" int chk_peer_gather_rows(int dtype, const void* const* peer_tables, int64_t rows_per_owner, const int64_t* ids, int64_t n,
                                    int64_t width, void* out, void* stream) {
    if (n == 0 || width == 0) return CHK_OK;
    if (n < 0 || width < 0 || rows_per_owner < 1 || !peer_tables || !ids || !out) { chk_set_error("chk_peer_gather_rows: bad argument"); return CHK_EINVAL; }
    const int64_t threads = width == 1 ? n : n * 32;
    int64_t blocks = (threads + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CHK_F32) peer_gather_rows_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((const float* const*)peer_tables, rows_per_owner, ids, n, width, (float*)out);
    else if (dtype == CHK_F64) peer_gather_rows_kernel<double><<<(unsigned)blocks, 256, 0, st>>>((const double* const*)peer_tables, rows_per_owner, ids, n, width, (double*)out);
    else { chk_set_error("unknown dtype %d", dtype); return CHK_EINVAL; }
    CHK_CUDA_LAUNCH_CHECK("peer_gather_rows_kernel");
    return CHK_OK;
}
