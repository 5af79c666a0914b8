// Training-step kernels around K1 / K3: everything of one optimisation step that is not the
// query transform or the pair scoring, without torch glue ops and without floating-point atomics.
//
//   chk_train_prep      the sampler of KGOptimizer.get_neg_samples (reference optimizers/kg_optimizer.py:92-99: uniform
//                       tail ids != true tail; with double_neg also a corrupted head per negative, :78-91) as a
//                       counter-based generator (Philox4x32-10) on the device, and the id arrays the kernels consume.
//   chk_group_build     groups the slots of a step (one slot = one gradient-row contribution) by the table row they
//                       name: per-row count, segment base, slot order.  Ids only, so it runs beside the forward pass.
//   chk_reduce_apply    what embedding_dense_backward + torch.optim.Adagrad.step do (run.py:205,
//                       optimizers/kg_optimizer.py:265-270), fused and row-sparse: every touched row's contributions are
//                       summed by ONE warp in ascending slot order (bit-reproducible; duplicates are segment-reduced, not
//                       added with atomics) and the Adagrad update is applied to the row in place — or the
//                       row sum is written into a dense gradient (other optimizers, dense all_reduce in data parallel).
//                       With world > 1 the slots of every rank (all_gathered contributions) are reduced in (rank, slot)
//                       order: the receive side of the sparse embedding-gradient exchange, identical on every replica.
//   chk_dense_apply     torch.optim.Adagrad / Adam (defaults) over whole tables from a dense gradient, which it clears.
//   chk_rowsum_groups   out[b,:] = sum_j in[b,j,:] in ascending j (double_neg: per-pair relation-row gradients).
#include "chk_common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------- sampler
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}
// uniform integer in [0, n) from 64 random bits (multiply-high: bias < n / 2^64)
__device__ __forceinline__ int64_t uniform_below(uint32_t a, uint32_t b, int64_t n) {
    return (int64_t)__umul64hi(((unsigned long long)a << 32) | b, (unsigned long long)n);
}

struct PrepArgs {
    const int64_t* batch; int64_t B, neg, n_entities; int double_neg;
    const int64_t* inj_tails; const int64_t* inj_heads;      // [B, neg] injected negatives (tests / overridden sampler) or NULL
    unsigned long long seed; const int* step_id; unsigned stream;
    int64_t* heads; int64_t* rels; int64_t* tails;
};

__global__ void __launch_bounds__(256) train_prep_kernel(PrepArgs A) {
    const int64_t nt = A.neg + 1, total = A.B * nt;
    const unsigned step = (unsigned)*A.step_id;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / nt, j = i - b * nt;
        const int64_t h = A.batch[3 * b], r = A.batch[3 * b + 1], t = A.batch[3 * b + 2];
        int64_t tj = t, hj = h;
        if (j > 0) {
            const int64_t k = b * A.neg + (j - 1);
            if (A.inj_tails) {
                tj = A.inj_tails[k];
                if (A.double_neg) hj = A.inj_heads ? A.inj_heads[k] : h;
            } else {
                const uint4 rnd = philox4x32_10(make_uint4((uint32_t)k, (uint32_t)(k >> 32), step, A.stream),
                                                make_uint2((uint32_t)A.seed, (uint32_t)(A.seed >> 32)));
                tj = uniform_below(rnd.x, rnd.y, A.n_entities - 1);
                tj += tj >= t;                                                   // uniform over the entities != true tail
                if (A.double_neg) { hj = uniform_below(rnd.z, rnd.w, A.n_entities - 1); hj += hj >= h; }
            }
        }
        A.tails[i] = tj;
        if (A.double_neg) { A.heads[i] = hj; A.rels[i] = r; }
        else if (j == 0) { A.heads[b] = h; A.rels[b] = r; }
    }
}

// ---------------------------------------------------------------------------------------------------- grouping
// work (int32): [0] nseg  [1] cursor  [2..3] pad | count[n_keys] | base[n_keys] | pos[total] | order[total] | seg[total]
struct GroupView {
    int* hdr; int* count; int* base; int* pos; int* order; int* seg;
};
__host__ __device__ inline GroupView group_view(void* work, int64_t n_keys, int64_t total) {
    int* w = (int*)work;
    GroupView v;
    v.hdr = w; v.count = w + 4; v.base = v.count + n_keys; v.pos = v.base + n_keys; v.order = v.pos + total; v.seg = v.order + total;
    return v;
}

__global__ void __launch_bounds__(256) group_count_kernel(const int64_t* __restrict__ ids, int64_t total, GroupView v) {
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x)
        v.pos[g] = atomicAdd(v.count + ids[g], 1);
}
__global__ void __launch_bounds__(256) group_alloc_kernel(const int64_t* __restrict__ ids, int64_t total, GroupView v) {
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        if (v.pos[g] != 0) continue;                                  // the slot that arrived first allocates its row's segment
        const int id = (int)ids[g];
        v.base[id] = atomicAdd(v.hdr + 1, v.count[id]);
        v.seg[atomicAdd(v.hdr, 1)] = id;
    }
}
__global__ void __launch_bounds__(256) group_order_kernel(const int64_t* __restrict__ ids, int64_t total, GroupView v) {
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x)
        v.order[v.base[ids[g]] + v.pos[g]] = (int)g;
}

// ---------------------------------------------------------------------------------------------------- reduce + apply
constexpr int RWARPS = 8;                 // warps per CTA
constexpr int SORT_CAP = 2048;            // slots of one segment a warp can sort in shared memory

template <typename T> struct RCol {
    T* param; T* s0; T* dense; int width;
    const T* src[2]; int lo[2], hi[2]; int64_t rstride[2];
};
template <typename T> struct RGroup {
    GroupView v; const int64_t* ids; int slots_per_rank; int total; int n_cols; int single;   // single: every slot names row 0 (no grouping)
    RCol<T> col[CHK_RED_MAX_COLS];
};
template <typename T> struct RArgsStep {
    RGroup<T> g[CHK_RED_MAX_GROUPS]; int n_groups; int opt; const double* hyper;
};

__device__ __forceinline__ int warp_bitonic_sort(int v, int lane) {          // ascending across the 32 lanes
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const int o = __shfl_xor_sync(CHK_FULL, v, j);
            const bool up = ((lane & k) == 0) == ((lane & j) == 0);
            v = up ? min(v, o) : max(v, o);
        }
    return v;
}

// sort buf[0..n) ascending (n <= SORT_CAP, buf padded with INT_MAX up to the next power of two) by one warp
__device__ __forceinline__ void warp_smem_sort(int* buf, int n, int lane) {
    int np = 64; while (np < n) np <<= 1;
    for (int i = n + lane; i < np; i += 32) buf[i] = 0x7fffffff;
    __syncwarp();
    for (int k = 2; k <= np; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < (np >> 1); t += 32) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));          // index with bit j clear
                const int a = buf[i], b = buf[i | j];
                const bool up = (i & k) == 0;
                if ((a > b) == up) { buf[i] = b; buf[i | j] = a; }
            }
            __syncwarp();
        }
}

template <typename T>
__device__ __forceinline__ void adagrad_apply(T& p, T g, T& a, T lr, T eps) {
    a = Sc<T>::fma_(g, g, a);
    p -= lr * g / (Sc<T>::sqrt_(a) + eps);
}

template <typename T> struct V2;
template <> struct V2<float> { using type = float2; };
template <> struct V2<double> { using type = double2; };

// address of the contribution row of global slot s for source i of a column (rank-major slot numbering)
template <typename T>
__device__ __forceinline__ const T* src_row(const RCol<T>& c, int i, int s, int spr) {
    const int k = s / spr, ls = s - k * spr;
    if (ls < c.lo[i] || ls >= c.hi[i]) return nullptr;
    return c.src[i] + (int64_t)k * c.rstride[i] + (int64_t)(ls - c.lo[i]) * c.width;
}

template <typename T>
__global__ void __launch_bounds__(RWARPS * 32) reduce_apply_kernel(const RArgsStep<T> A) {
    using V = typename V2<T>::type;
    extern __shared__ int sort_smem[];                                 // [RWARPS][SORT_CAP]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int* sbuf = sort_smem + warp * SORT_CAP;
    const RGroup<T>& G = A.g[blockIdx.y];
    const T lr = (T)A.hyper[0], eps = (T)A.hyper[1];
    const int nseg = G.single ? 1 : G.v.hdr[0];
    const int spr = G.slots_per_rank;
    for (int sg = blockIdx.x * RWARPS + warp; sg < nseg; sg += gridDim.x * RWARPS) {
        int id = 0, len = G.total, base = 0;
        if (!G.single) { id = G.v.seg[sg]; len = G.v.count[id]; base = G.v.base[id]; }
        // ---- the segment's slots in ascending order: registers (<= 32), shared memory (<= SORT_CAP), else selection
        int mine = 0x7fffffff;
        const int* sorted = nullptr;                                   // non-null: sorted list in memory
        if (G.single) {
            sorted = nullptr;                                          // identity order, slot = position
        } else if (len <= 32) {
            if (lane < len) mine = G.v.order[base + lane];
            mine = warp_bitonic_sort(mine, lane);
        } else if (len <= SORT_CAP) {
            for (int i = lane; i < len; i += 32) sbuf[i] = G.v.order[base + i];
            warp_smem_sort(sbuf, len, lane);
            sorted = sbuf;
        } else {
            // pathological segment (one row named by > SORT_CAP slots): in-place selection sort in global memory by the warp
            int* o = G.v.order + base;
            for (int k = 0; k < len - 1; ++k) {
                int best = 0x7fffffff, at = -1;
                for (int i = k + lane; i < len; i += 32) { const int x = o[i]; if (x < best) { best = x; at = i; } }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    const int ob = __shfl_xor_sync(CHK_FULL, best, d), oa = __shfl_xor_sync(CHK_FULL, at, d);
                    if (ob < best) { best = ob; at = oa; }
                }
                if (lane == 0 && at != k) { o[at] = o[k]; o[k] = best; }
                __syncwarp();
            }
            sorted = o;
        }
        auto slot_at = [&](int k) -> int {
            if (G.single) return k;
            if (sorted) return sorted[k];
            return __shfl_sync(CHK_FULL, mine, k);
        };
        // ---- every column of the group: sum the contribution rows in slot order, then update / write the row
        for (int ci = 0; ci < G.n_cols; ++ci) {
            const RCol<T>& c = G.col[ci];
            const int64_t rowoff = (int64_t)id * c.width;
            if (c.width == 1) {
                // scalar column: lane-strided loads in slot order, fixed butterfly (deterministic for a given sorted order)
                T acc = T(0);
                for (int k0 = 0; k0 < len; k0 += 32) {
                    const int k = k0 + lane;
                    int s = (G.single || sorted) ? (k < len ? (G.single ? k : sorted[k]) : -1) : (k < len ? mine : -1);
                    T v = T(0);
                    if (s >= 0) {
#pragma unroll
                        for (int i = 0; i < 2; ++i) if (c.src[i]) { const T* p = src_row<T>(c, i, s, spr); if (p) v += *p; }
                    }
                    acc += warp_sum<T>(v);
                }
                if (lane == 0) {
                    if (c.dense) c.dense[rowoff] = acc;
                    else { T p = c.param[rowoff], a = c.s0[rowoff]; adagrad_apply<T>(p, acc, a, lr, eps); c.param[rowoff] = p; c.s0[rowoff] = a; }
                }
                continue;
            }
            const int w2 = c.width >> 1;                               // even widths: two elements per lane
            for (int c0 = 0; c0 < w2; c0 += 32) {
                const int cc = c0 + lane;
                const bool on = cc < w2;
                V acc; acc.x = T(0); acc.y = T(0);
                for (int k0 = 0; k0 < len; k0 += 4) {                  // four rows in flight, added in slot order
                    V v[4][2];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int k = k0 + u;
                        const int s = k < len ? slot_at(k) : -1;
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            v[u][i].x = T(0); v[u][i].y = T(0);
                            if (s >= 0 && c.src[i] && on) {
                                const T* p = src_row<T>(c, i, s, spr);
                                if (p) v[u][i] = reinterpret_cast<const V*>(p)[cc];
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int i = 0; i < 2; ++i) { acc.x += v[u][i].x; acc.y += v[u][i].y; }
                }
                if (on) {
                    if (c.dense) reinterpret_cast<V*>(c.dense + rowoff)[cc] = acc;
                    else {
                        V p = reinterpret_cast<V*>(c.param + rowoff)[cc], a = reinterpret_cast<V*>(c.s0 + rowoff)[cc];
                        adagrad_apply<T>(p.x, acc.x, a.x, lr, eps); adagrad_apply<T>(p.y, acc.y, a.y, lr, eps);
                        reinterpret_cast<V*>(c.param + rowoff)[cc] = p; reinterpret_cast<V*>(c.s0 + rowoff)[cc] = a;
                    }
                }
            }
        }
        __syncwarp();
        if (!G.single && lane == 0) G.v.count[id] = 0;                 // the count array is all-zero again for the next step
    }
}

// reset the grouping headers (nseg, cursor) after the reduce of a step
struct HdrList { int* h[CHK_RED_MAX_GROUPS]; int n; };
template <typename T>
__global__ void __launch_bounds__(256) step_finish_kernel(HdrList H, const T* __restrict__ loss_part, int64_t n_loss, T* __restrict__ loss_accum,
                                                          int* __restrict__ step_id) {
    __shared__ T red[8];
    if (threadIdx.x < H.n) { H.h[threadIdx.x][0] = 0; H.h[threadIdx.x][1] = 0; }
    if (loss_part) {                                                   // fixed-order sum of the per-row loss partials
        T acc = T(0);
        for (int64_t i = threadIdx.x; i < n_loss; i += blockDim.x) acc += loss_part[i];
        acc = warp_sum<T>(acc);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            T s = T(0);
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
            *loss_accum += s;
        }
    }
    if (threadIdx.x == 0 && step_id) *step_id += 1;
}

// ---------------------------------------------------------------------------------------------------- dense apply
struct DTab { void* param; void* grad; void* s0; void* s1; int64_t n; };
struct DList { DTab t[CHK_MAX_TABLES]; };

// torch.optim.Adagrad (lr_decay = 0, weight_decay = 0) / torch.optim.Adam (betas, eps; no weight decay, no amsgrad) over whole
// tables from a dense gradient; the gradient is cleared.  Adam follows torch's arithmetic: exp_avg.lerp_(g, 1-b1);
// exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2); denom = sqrt(exp_avg_sq)/sqrt(1-b2^t) + eps; p.addcdiv_(exp_avg, denom, -lr/(1-b1^t)).
template <typename T, int OPT>
__global__ void __launch_bounds__(256) dense_apply_kernel(DList L, const double* __restrict__ hyper, const int* __restrict__ step_id) {
    const DTab d = L.t[blockIdx.y];
    T* p = (T*)d.param; T* g = (T*)d.grad; T* s0 = (T*)d.s0; T* s1 = (T*)d.s1;
    const T lr = (T)hyper[0], eps = (T)hyper[1];
    T b2 = T(0), w1 = T(0), w2 = T(0), step_size = T(0), bc2s = T(1);
    if (OPT == CHK_OPT_ADAM) {
        const double beta1 = hyper[4], beta2 = hyper[5];
        const double t = (double)(*step_id);                            // 1-based step number
        b2 = (T)beta2; w1 = (T)(1.0 - beta1); w2 = (T)(1.0 - beta2);
        step_size = (T)(hyper[0] / (1.0 - pow(beta1, t)));
        bc2s = (T)sqrt(1.0 - pow(beta2, t));
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < d.n; i += (int64_t)gridDim.x * blockDim.x) {
        const T gv = g[i];
        if (OPT == CHK_OPT_ADAGRAD) {
            if (gv != T(0)) { T pv = p[i], av = s0[i]; adagrad_apply<T>(pv, gv, av, lr, eps); p[i] = pv; s0[i] = av; g[i] = T(0); }
        } else {
            T m = s0[i], v = s1[i];
            m = m + w1 * (gv - m);                                      // lerp_(g, 1 - beta1), weight < 0.5 form
            v = Sc<T>::fma_(w2 * gv, gv, v * b2);                       // mul_(beta2).addcmul_(g, g, value = 1 - beta2)
            const T denom = Sc<T>::sqrt_(v) / bc2s + eps;
            p[i] = p[i] - step_size * (m / denom);
            s0[i] = m; s1[i] = v;
            if (gv != T(0)) g[i] = T(0);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) rowsum_groups_kernel(const T* __restrict__ in, int64_t B, int64_t nj, int64_t width, T* __restrict__ out) {
    const int64_t total = B * width;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / width, c = i - b * width;
        const T* p = in + b * nj * width + c;
        T acc = T(0);
        for (int64_t j = 0; j < nj; ++j) acc += p[j * width];
        out[i] = acc;
    }
}

// N3 / F2 of the positive call's factors (reference optimizers/regularizers.py:21-58; models/base.py:175-198): value
// w * sum_f sum |f|^p / B into the row's loss partial, gradient w*p*|f|^(p-2) f / B into the row's contribution rows.
template <typename T>
__global__ void __launch_bounds__(128) reg_factors_kernel(int power, T weight, const double* __restrict__ hyper, int64_t B,
                                                          const T* __restrict__ ent, int64_t ew, const T* __restrict__ rel, int64_t rw,
                                                          const int64_t* __restrict__ heads, int64_t hs, const int64_t* __restrict__ rels,
                                                          const int64_t* __restrict__ tails, int64_t ts,
                                                          T* __restrict__ g_ent, int64_t ges, T* __restrict__ g_rel, int64_t grs,
                                                          T* __restrict__ g_tail, int64_t gts, T* __restrict__ loss_part) {
    __shared__ T red[4];
    const int64_t b = blockIdx.x;
    const int64_t n_valid = (int64_t)hyper[3];
    if (b >= n_valid) return;                                           // padding row of a ragged batch
    const T scale = (T)hyper[6];                                        // 1 / (rows of the GLOBAL batch)
    auto one = [&](const T* f, T* g, int64_t w) -> T {
        T acc = T(0);
        for (int64_t c = threadIdx.x; c < w; c += blockDim.x) {
            const T x = f[c], ax = Sc<T>::abs_(x);
            if (power == 3) { acc += ax * ax * ax; g[c] += T(3) * ax * x * weight * scale; }
            else { acc += x * x; g[c] += T(2) * x * weight * scale; }
        }
        return acc;
    };
    T acc = one(ent + heads[b * hs] * ew, g_ent + b * ges, ew);
    acc += one(rel + rels[b * hs] * rw, g_rel + b * grs, rw);
    acc += one(ent + tails[b * ts] * ew, g_tail + b * gts, ew);
    acc = warp_sum<T>(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) loss_part[b] += (red[0] + red[1] + red[2] + red[3]) * weight * scale;
}

int grid_for(int64_t items, int per_block, int cap) {
    int64_t b = (items + per_block - 1) / per_block;
    if (b < 1) b = 1;
    if (b > cap) b = cap;
    return (int)b;
}

}  // namespace

extern "C" int chk_train_prep(const int64_t* batch, int64_t B, int64_t neg, int64_t n_entities, int double_neg,
                              const int64_t* injected_tails, const int64_t* injected_heads, uint64_t seed,
                              const int32_t* step_id, uint32_t stream_id, int64_t* heads, int64_t* rels, int64_t* tails, void* stream) {
    if (B == 0) return CHK_OK;
    if (B < 0 || neg < 0 || n_entities < 2 || !batch || !step_id || !heads || !rels || !tails) { chk_set_error("chk_train_prep: bad argument"); return CHK_EINVAL; }
    PrepArgs A{batch, B, neg, n_entities, double_neg, injected_tails, injected_heads, seed, step_id, stream_id, heads, rels, tails};
    train_prep_kernel<<<grid_for(B * (neg + 1), 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(A);
    CHK_CUDA_LAUNCH_CHECK("train_prep_kernel");
    return CHK_OK;
}

extern "C" int64_t chk_group_workspace_bytes(int64_t n_keys, int64_t total_slots) {
    if (n_keys < 1 || total_slots < 0 || n_keys > 0x7fffffff || total_slots > 0x7fffffff) return -1;
    return (int64_t)sizeof(int) * (4 + 2 * n_keys + 3 * total_slots);
}

extern "C" int chk_group_build(const int64_t* ids, int64_t total_slots, int64_t n_keys, void* work, void* stream) {
    if (total_slots == 0) return CHK_OK;
    if (!ids || !work || chk_group_workspace_bytes(n_keys, total_slots) < 0) { chk_set_error("chk_group_build: bad argument"); return CHK_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    GroupView v = group_view(work, n_keys, total_slots);
    const int grid = grid_for(total_slots, 256, 148 * 8);
    group_count_kernel<<<grid, 256, 0, st>>>(ids, total_slots, v);
    group_alloc_kernel<<<grid, 256, 0, st>>>(ids, total_slots, v);
    group_order_kernel<<<grid, 256, 0, st>>>(ids, total_slots, v);
    CHK_CUDA_LAUNCH_CHECK("group kernels");
    return CHK_OK;
}

template <typename T>
static int reduce_apply_t(int opt, const chk_red_group* groups, int n_groups, const double* hyper, cudaStream_t st) {
    RArgsStep<T> A{};
    A.n_groups = n_groups; A.opt = opt; A.hyper = hyper;
    int64_t max_seg = 1;
    for (int gi = 0; gi < n_groups; ++gi) {
        const chk_red_group& g = groups[gi];
        RGroup<T>& R = A.g[gi];
        const int64_t total = g.slots_per_rank * (int64_t)g.world;
        if (g.n_cols < 1 || g.n_cols > CHK_RED_MAX_COLS || g.slots_per_rank < 1 || g.world < 1 || total > 0x7fffffff ||
            (!g.single_row && (!g.ids || !g.work || g.n_keys < 1))) { chk_set_error("chk_reduce_apply: bad group %d", gi); return CHK_EINVAL; }
        R.single = g.single_row; R.ids = g.ids; R.slots_per_rank = (int)g.slots_per_rank; R.total = (int)total; R.n_cols = g.n_cols;
        if (!g.single_row) R.v = group_view(g.work, g.n_keys, total);
        for (int ci = 0; ci < g.n_cols; ++ci) {
            const chk_red_col& c = g.cols[ci];
            if (!c.param || c.width < 1 || (c.width > 1 && (c.width & 1)) || (!c.dense_grad && (opt != CHK_OPT_ADAGRAD || !c.state0)) || !c.src[0]) {
                chk_set_error("chk_reduce_apply: bad column %d of group %d (in-place update needs CHK_OPT_ADAGRAD + state0; width 1 or even)", ci, gi);
                return CHK_EINVAL;
            }
            RCol<T>& C = R.col[ci];
            C.param = (T*)c.param; C.s0 = (T*)c.state0; C.dense = (T*)c.dense_grad; C.width = (int)c.width;
            for (int i = 0; i < 2; ++i) { C.src[i] = (const T*)c.src[i]; C.lo[i] = (int)c.lo[i]; C.hi[i] = (int)c.hi[i]; C.rstride[i] = c.rank_stride[i]; }
        }
        if (!g.single_row && total > max_seg) max_seg = total;
    }
    static bool attr_set = false;
    const size_t smem = (size_t)RWARPS * SORT_CAP * sizeof(int);
    if (!attr_set) {
        cudaFuncSetAttribute(reduce_apply_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(reduce_apply_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set = true;
    }
    dim3 grid((unsigned)grid_for(max_seg, RWARPS, 148 * 6), (unsigned)n_groups);
    reduce_apply_kernel<T><<<grid, RWARPS * 32, smem, st>>>(A);
    CHK_CUDA_LAUNCH_CHECK("reduce_apply_kernel");
    return CHK_OK;
}

extern "C" int chk_reduce_apply(int dtype, int opt, const chk_red_group* groups, int n_groups, const double* hyper, void* stream) {
    if (n_groups < 1 || n_groups > CHK_RED_MAX_GROUPS || !groups || !hyper) { chk_set_error("chk_reduce_apply: bad argument"); return CHK_EINVAL; }
    if (dtype == CHK_F32) return reduce_apply_t<float>(opt, groups, n_groups, hyper, (cudaStream_t)stream);
    if (dtype == CHK_F64) return reduce_apply_t<double>(opt, groups, n_groups, hyper, (cudaStream_t)stream);
    chk_set_error("unknown dtype %d", dtype); return CHK_EINVAL;
}

extern "C" int chk_step_finish(int dtype, void* const* group_works, int n_groups, const void* loss_part, int64_t n_loss, void* loss_accum,
                               int32_t* step_id, void* stream) {
    if (n_groups < 0 || n_groups > CHK_RED_MAX_GROUPS || (loss_part && !loss_accum)) { chk_set_error("chk_step_finish: bad argument"); return CHK_EINVAL; }
    HdrList H{}; H.n = n_groups;
    for (int i = 0; i < n_groups; ++i) { if (!group_works[i]) { chk_set_error("chk_step_finish: null workspace"); return CHK_EINVAL; } H.h[i] = (int*)group_works[i]; }
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CHK_F32) step_finish_kernel<float><<<1, 256, 0, st>>>(H, (const float*)loss_part, n_loss, (float*)loss_accum, step_id);
    else if (dtype == CHK_F64) step_finish_kernel<double><<<1, 256, 0, st>>>(H, (const double*)loss_part, n_loss, (double*)loss_accum, step_id);
    else { chk_set_error("unknown dtype %d", dtype); return CHK_EINVAL; }
    CHK_CUDA_LAUNCH_CHECK("step_finish_kernel");
    return CHK_OK;
}

extern "C" int chk_dense_apply(int dtype, int opt, const chk_dense_tab* tabs, int n_tables, const double* hyper, const int32_t* step_id, void* stream) {
    if (n_tables < 1 || n_tables > CHK_MAX_TABLES || !tabs || !hyper || (opt != CHK_OPT_ADAGRAD && opt != CHK_OPT_ADAM) || (opt == CHK_OPT_ADAM && !step_id)) {
        chk_set_error("chk_dense_apply: bad argument"); return CHK_EINVAL;
    }
    DList L{}; int64_t mx = 0;
    for (int i = 0; i < n_tables; ++i) {
        const chk_dense_tab& t = tabs[i];
        if (!t.param || !t.grad || !t.state0 || (opt == CHK_OPT_ADAM && !t.state1) || t.n < 0) { chk_set_error("chk_dense_apply: bad table %d", i); return CHK_EINVAL; }
        L.t[i] = DTab{t.param, t.grad, t.state0, t.state1, t.n};
        if (t.n > mx) mx = t.n;
    }
    if (mx == 0) return CHK_OK;
    dim3 grid((unsigned)grid_for(mx, 256 * 4, 148 * 8), (unsigned)n_tables);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CHK_F32 && opt == CHK_OPT_ADAGRAD) dense_apply_kernel<float, CHK_OPT_ADAGRAD><<<grid, 256, 0, st>>>(L, hyper, step_id);
    else if (dtype == CHK_F32) dense_apply_kernel<float, CHK_OPT_ADAM><<<grid, 256, 0, st>>>(L, hyper, step_id);
    else if (dtype == CHK_F64 && opt == CHK_OPT_ADAGRAD) dense_apply_kernel<double, CHK_OPT_ADAGRAD><<<grid, 256, 0, st>>>(L, hyper, step_id);
    else if (dtype == CHK_F64) dense_apply_kernel<double, CHK_OPT_ADAM><<<grid, 256, 0, st>>>(L, hyper, step_id);
    else { chk_set_error("unknown dtype %d", dtype); return CHK_EINVAL; }
    CHK_CUDA_LAUNCH_CHECK("dense_apply_kernel");
    return CHK_OK;
}

extern "C" int chk_rowsum_groups(int dtype, const void* in, int64_t B, int64_t nj, int64_t width, void* out, void* stream) {
    if (B == 0 || width == 0) return CHK_OK;
    if (B < 0 || nj < 1 || width < 0 || !in || !out) { chk_set_error("chk_rowsum_groups: bad argument"); return CHK_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = grid_for(B * width, 256, 148 * 8);
    if (dtype == CHK_F32) rowsum_groups_kernel<float><<<grid, 256, 0, st>>>((const float*)in, B, nj, width, (float*)out);
    else if (dtype == CHK_F64) rowsum_groups_kernel<double><<<grid, 256, 0, st>>>((const double*)in, B, nj, width, (double*)out);
    else { chk_set_error("unknown dtype %d", dtype); return CHK_EINVAL; }
    CHK_CUDA_LAUNCH_CHECK("rowsum_groups_kernel");
    return CHK_OK;
}

extern "C" int chk_reg_factors(int dtype, int power, double weight, const double* hyper, int64_t B,
                               const void* entity, int64_t ent_width, const void* rel, int64_t rel_width,
                               const int64_t* heads, int64_t head_stride, const int64_t* rels, const int64_t* tails, int64_t tail_stride,
                               void* g_ent_rows, int64_t g_ent_stride, void* g_rel_rows, int64_t g_rel_stride,
                               void* g_tail_rows, int64_t g_tail_stride, void* loss_part, void* stream) {
    if (B == 0) return CHK_OK;
    if (B < 0 || (power != 2 && power != 3) || !hyper || !entity || !rel || !heads || !rels || !tails || !g_ent_rows || !g_rel_rows || !g_tail_rows || !loss_part) {
        chk_set_error("chk_reg_factors: bad argument"); return CHK_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CHK_F32) reg_factors_kernel<float><<<(unsigned)B, 128, 0, st>>>(power, (float)weight, hyper, B, (const float*)entity, ent_width, (const float*)rel, rel_width,
                                                                                 heads, head_stride, rels, tails, tail_stride, (float*)g_ent_rows, g_ent_stride,
                                                                                 (float*)g_rel_rows, g_rel_stride, (float*)g_tail_rows, g_tail_stride, (float*)loss_part);
    else if (dtype == CHK_F64) reg_factors_kernel<double><<<(unsigned)B, 128, 0, st>>>(power, weight, hyper, B, (const double*)entity, ent_width, (const double*)rel, rel_width,
                                                                                      heads, head_stride, rels, tails, tail_stride, (double*)g_ent_rows, g_ent_stride,
                                                                                      (double*)g_rel_rows, g_rel_stride, (double*)g_tail_rows, g_tail_stride, (double*)loss_part);
    else { chk_set_error("unknown dtype %d", dtype); return CHK_EINVAL; }
    CHK_CUDA_LAUNCH_CHECK("reg_factors_kernel");
    return CHK_OK;
}
