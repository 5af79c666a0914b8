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
#include <cstdlib>
#include "chk_common.cuh"
#ifndef CHK_COEF_ONEPASS
#define CHK_COEF_ONEPASS 1     // fp32 wide rows with pair coefficients: the whole row in ONE pass at 2 CTAs per SM and 128 registers (r2: 140 -> 126 us
#endif                         // at the 4M-entity config; two passes of 5 + 4 chunks at 80 registers spilled and broadcast the descriptors twice).
// Tried and dropped in r2 (measured, no gain): prefetching the next segment's parameter / state rows to shared memory with
// cp.async (129 us), computing the next segment's descriptors one iteration ahead (177 us: spills).  The kernel moves 428 MB of
// random 2 KB rows read-modify-write at 3.4 TB/s (ncu: profiles/r2_final_train_big4m_raw.csv); a streaming copy reaches 6.5.

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
// work (int32): [0] short segments  [1] cursor  [2] long segments  [3] pad | count[n_keys] | base[n_keys] | pos[total] | order[total] |
//               seg[total] (short-segment rows from the front, long-segment rows from the back) | slen[total] | sbase[total]
//               (length and start in `order` of the segment listed at the same position of seg)
struct GroupView {
    int* hdr; int* count; int* base; int* pos; int* order; int* seg; int* slen; int* sbase;
};
__host__ __device__ inline GroupView group_view(void* work, int64_t n_keys, int64_t total) {
    int* w = (int*)work;
    GroupView v;
    v.hdr = w; v.count = w + 4; v.base = v.count + n_keys; v.pos = v.base + n_keys; v.order = v.pos + total; v.seg = v.order + total;
    v.slen = v.seg + total; v.sbase = v.slen + total;
    return v;
}

// Slots whose row lies outside [own_lo, own_hi) are not grouped (pos = -1): the rank does not own the row (owner-sharded tables).
__global__ void __launch_bounds__(256) group_count_kernel(const int64_t* __restrict__ ids, int64_t total, GroupView v, int64_t own_lo, int64_t own_hi) {
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t id = ids[g];
        v.pos[g] = (id >= own_lo && id < own_hi) ? atomicAdd(v.count + id, 1) : -1;
    }
}
// The slot that arrived first at a row (pos == 0) allocates the row's segment in `order` and appends the row to the list of
// short segments (<= SHORT_MAX slots: one warp reduces them from registers) or, from the end of the same array, to the list of
// long segments (a whole CTA each).  The two global counters are bumped once per warp (ballot + prefix sum), not per leader.
constexpr int SHORT_MAX = 32;
__global__ void __launch_bounds__(256) group_alloc_kernel(const int64_t* __restrict__ ids, int64_t total, GroupView v) {
    const int lane = threadIdx.x & 31;
    const int64_t span = (total + 31) & ~int64_t(31);
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < span; g += (int64_t)gridDim.x * blockDim.x) {
        const bool lead = g < total && v.pos[g] == 0;
        int id = 0, cnt = 0;
        if (lead) { id = (int)ids[g]; cnt = v.count[id]; }
        const bool is_long = lead && cnt > SHORT_MAX;
        const unsigned m_lead = __ballot_sync(CHK_FULL, lead), m_long = __ballot_sync(CHK_FULL, is_long);
        if (m_lead == 0) continue;
        int pre = cnt;                                                   // inclusive prefix sum of the leaders' counts
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(CHK_FULL, pre, o); if (lane >= o) pre += t; }
        const int warp_total = __shfl_sync(CHK_FULL, pre, 31);
        const unsigned m_short = m_lead & ~m_long;
        int base0 = 0, s0 = 0, l0 = 0;
        if (lane == 0) {
            base0 = atomicAdd(v.hdr + 1, warp_total);
            if (m_short) s0 = atomicAdd(v.hdr, __popc(m_short));
            if (m_long) l0 = atomicAdd(v.hdr + 2, __popc(m_long));
        }
        base0 = __shfl_sync(CHK_FULL, base0, 0); s0 = __shfl_sync(CHK_FULL, s0, 0); l0 = __shfl_sync(CHK_FULL, l0, 0);
        if (lead) {
            const int b0 = base0 + pre - cnt;
            v.base[id] = b0;
            const unsigned below = (1u << lane) - 1u;
            const int at = is_long ? (int)total - 1 - (l0 + __popc(m_long & below)) : s0 + __popc(m_short & below);
            v.seg[at] = id; v.slen[at] = cnt; v.sbase[at] = b0;
        }
    }
}
__global__ void __launch_bounds__(256) group_order_kernel(const int64_t* __restrict__ ids, int64_t total, GroupView v) {
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t id = ids[g];
        const int p = v.pos[g];
        if (p < 0) continue;                     // not this rank's row
        v.order[v.base[id] + p] = (int)g;
        if (p == 0) v.count[id] = 0;             // nobody reads the count any more (the segment list carries the lengths): ready for the next step
    }
}

// ---------------------------------------------------------------------------------------------------- reduce + apply
constexpr int RWARPS = 8;                 // warps per CTA
constexpr int SORT_CAP = 4096;            // slots of a long segment the CTA sorts in shared memory (longer: in place, global)

template <typename T> struct RCol {
    T* param; T* s0; T* dense; int width;
    const T* src[2]; int lo[2], hi[2]; int64_t rstride[2];
    const T* coef; int pair_nt; int64_t cstride;      // computed source (see chk_red_col): src[1] = query rows, coef = (c1, c2, c3, pad) per pair
};
template <typename T> struct RGroup {
    GroupView v; const int64_t* ids; int slots_per_rank; int total; int n_cols; int single;   // single: every slot names row 0 (no grouping)
    RCol<T> col[CHK_RED_MAX_COLS];
};
template <typename T> struct RArgsStep {
    RGroup<T> g[CHK_RED_MAX_GROUPS]; int n_groups; int opt; const double* hyper;
    // optional end-of-step duties of the LAST block to finish (saves the chk_step_finish launch)
    int finish; const T* loss_part; int64_t n_loss; T* loss_accum; int* step_id; int* ticket;
};

template <int N>
__device__ __forceinline__ int warp_bitonic_sort_n(int v, int lane) {        // ascending across the first N lanes (N = 2^m <= 32)
#pragma unroll
    for (int k = 2; k <= N; k <<= 1)
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const int o = __shfl_xor_sync(CHK_FULL, v, j);
            const bool up = ((lane & k) == 0) == ((lane & j) == 0);
            v = up ? min(v, o) : max(v, o);
        }
    return v;
}
// lanes >= len hold INT_MAX; only as many stages as the (warp-uniform) length needs
__device__ __forceinline__ int warp_sort_len(int v, int lane, int len) {
    if (len <= 1) return v;
    if (len <= 2) return warp_bitonic_sort_n<2>(v, lane);
    if (len <= 4) return warp_bitonic_sort_n<4>(v, lane);
    if (len <= 8) return warp_bitonic_sort_n<8>(v, lane);
    if (len <= 16) return warp_bitonic_sort_n<16>(v, lane);
    return warp_bitonic_sort_n<32>(v, lane);
}
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

// Ascending sort of buf[0..n) (shared or global memory) by the whole CTA: bitonic network in its ascending-only form (the
// first step of every merge compares mirrored positions), so positions >= n act as +infinity without being stored and any
// n works in place.
__device__ __forceinline__ void cta_sort(int* buf, int n) {
    int np = 2; while (np < n) np <<= 1;
    for (int k = 2; k <= np; k <<= 1) {
        for (int t = threadIdx.x; t < (np >> 1); t += blockDim.x) {                 // mirror step
            const int h = k >> 1;
            const int i = ((t / h) * k) + (t % h), l = i ^ (k - 1);
            if (l < n) { const int a = buf[i], b = buf[l]; if (a > b) { buf[i] = b; buf[l] = a; } }
        }
        __syncthreads();
        for (int j = k >> 2; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < (np >> 1); t += blockDim.x) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j;
                if (l < n) { const int a = buf[i], b = buf[l]; if (a > b) { buf[i] = b; buf[l] = a; } }
            }
            __syncthreads();
        }
    }
}

// torch.optim.Adagrad element update: sum += g*g; p -= lr * g / (sqrt(sum) + eps).  fp64: IEEE sqrt and division.  fp32: the
// SFU approximations (sqrt.approx / rcp.approx, <= 2 ulp each; sqrt.approx(0) = 0 and the denominator is >= eps > 0): the IEEE
// sequences cost ~25 instructions per element and made the row update instruction-bound (ncu r2: 16 elements per lane at rank
// 257); the deviation is a few ulp of an update that is itself ~1e-2 of the parameter.
template <typename T>
__device__ __forceinline__ void adagrad_apply(T& p, T g, T& a, T lr, T eps) {
    a = Sc<T>::fma_(g, g, a);
    p -= lr * g / (Sc<T>::sqrt_(a) + eps);
}
template <>
__device__ __forceinline__ void adagrad_apply<float>(float& p, float g, float& a, float lr, float eps) {
    a = __fmaf_rn(g, g, a);
    float sq, rc;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(sq) : "f"(a));
    asm("rcp.approx.f32 %0, %1;" : "=f"(rc) : "f"(sq + eps));
    p = __fmaf_rn(-(lr * g), rc, p);
}

template <typename T> struct V2;
template <> struct V2<float> { using type = float2; };
template <> struct V2<double> { using type = double2; };

// address of the contribution row of global slot s for source i of a column (rank-major slot numbering)
template <typename T>
__device__ __forceinline__ const T* src_row(const RCol<T>& c, int i, int s, int spr) {
    int k = 0, ls = s;
    if (s >= spr) { k = s / spr; ls = s - k * spr; }
    if (ls < c.lo[i] || ls >= c.hi[i] || !c.src[i]) return nullptr;
    return c.src[i] + (int64_t)k * c.rstride[i] + (int64_t)(ls - c.lo[i]) * c.width;
}

// One warp sums NCH 64-element chunks (two elements per lane each), starting at chunk ch0, of one column over the slots
// slot_at(0..len) IN THAT ORDER, then applies Adagrad to / writes the dense gradient of those elements of row `id`.
// The parameter / state elements are loaded before the summation so their latency overlaps the contribution rows'.
template <typename T, int NCH, typename SlotAt>
__device__ __forceinline__ void col_chunks(const RCol<T>& c, int id, int len, int spr, int ch0, T lr, T eps, int lane, SlotAt slot_at) {
    using V = typename V2<T>::type;
    const int w2 = c.width >> 1;
    const int64_t rowoff = (int64_t)id * c.width;
    V acc[NCH], pv[NCH], av[NCH];
    bool on[NCH];
#pragma unroll
    for (int h = 0; h < NCH; ++h) {
        acc[h].x = T(0); acc[h].y = T(0);
        on[h] = (ch0 + h) * 32 + lane < w2;
        if (!c.dense && on[h]) {
            pv[h] = reinterpret_cast<const V*>(c.param + rowoff)[(ch0 + h) * 32 + lane];
            av[h] = reinterpret_cast<const V*>(c.s0 + rowoff)[(ch0 + h) * 32 + lane];
        }
    }
    constexpr int U = NCH <= 2 ? 4 : 2;                             // rows in flight
    for (int k0 = 0; k0 < len; k0 += U) {
        V v[U][NCH];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int s = k0 + u < len ? slot_at(k0 + u) : -1;
            const T* p = nullptr;                                       // a slot belongs to at most one source range
            if (s >= 0) {
                p = src_row<T>(c, 0, s, spr);
                if (!p && c.src[1]) p = src_row<T>(c, 1, s, spr);
            }
#pragma unroll
            for (int h = 0; h < NCH; ++h) {
                v[u][h].x = T(0); v[u][h].y = T(0);
                if (p && on[h]) v[u][h] = reinterpret_cast<const V*>(p)[(ch0 + h) * 32 + lane];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int h = 0; h < NCH; ++h) { acc[h].x += v[u][h].x; acc[h].y += v[u][h].y; }
    }
#pragma unroll
    for (int h = 0; h < NCH; ++h) {
        if (!on[h]) continue;
        const int cc = (ch0 + h) * 32 + lane;
        if (c.dense) reinterpret_cast<V*>(c.dense + rowoff)[cc] = acc[h];
        else {
            adagrad_apply<T>(pv[h].x, acc[h].x, av[h].x, lr, eps); adagrad_apply<T>(pv[h].y, acc[h].y, av[h].y, lr, eps);
            reinterpret_cast<V*>(c.param + rowoff)[cc] = pv[h]; reinterpret_cast<V*>(c.s0 + rowoff)[cc] = av[h];
        }
    }
}

// scalar column (bh, bt, c): lane-strided loads in slot order, fixed butterfly (deterministic for a given slot order)
template <typename T, typename SlotAtLane>
__device__ __forceinline__ void col_scalar(const RCol<T>& c, int id, int len, int spr, T lr, T eps, int lane, SlotAtLane slot_of_pos) {
    T acc = T(0);
    for (int k0 = 0; k0 < len; k0 += 32) {
        const int k = k0 + lane;
        const int s = k < len ? slot_of_pos(k) : -1;
        T v = T(0);
        if (s >= 0) {
#pragma unroll
            for (int i = 0; i < 2; ++i) if (c.src[i]) { const T* p = src_row<T>(c, i, s, spr); if (p) v += *p; }
        }
        acc += warp_sum<T>(v);
    }
    if (lane == 0) {
        if (c.dense) c.dense[id] = acc;
        else { T p = c.param[id], a = c.s0[id]; adagrad_apply<T>(p, acc, a, lr, eps); c.param[id] = p; c.s0[id] = a; }
    }
}

// MC = 64-element chunks a warp accumulates per work item.  MC = 2: narrow rows (rank <= 33), 64 registers, 4 CTAs per SM.
// MC = 9: wide rows (a whole rank-257 row per warp: parameter, state and contribution chunks all in flight), 128 registers.

// Short-segment column pass with LANE-PARALLEL row addressing: lane k holds the address of the contribution row of the k-th
// slot (rp, nullptr when the slot does not feed this column), so a row costs two shuffles + NCH loads + the adds instead of
// re-deriving its address per chunk.  Sums chunks ch0 .. ch0+NCH-1 (64 elements each) over k = 0..len-1 in that order.
template <typename T>
__device__ __forceinline__ const T* shfl_ptr(const T* p, int k) {
    const unsigned long long v = (unsigned long long)p;
    const unsigned lo = __shfl_sync(CHK_FULL, (unsigned)v, k), hi = __shfl_sync(CHK_FULL, (unsigned)(v >> 32), k);
    return (const T*)(((unsigned long long)hi << 32) | lo);
}
template <typename T, int NCH>
__device__ __forceinline__ void col_chunks_ptr(const RCol<T>& c, int id, int len, int ch0, T lr, T eps, int lane, const T* rp) {
    using V = typename V2<T>::type;
    const int w2 = c.width >> 1;
    const int64_t rowoff = (int64_t)id * c.width;
    V acc[NCH], pv[NCH], av[NCH];
    bool on[NCH];
#pragma unroll
    for (int h = 0; h < NCH; ++h) {
        acc[h].x = T(0); acc[h].y = T(0);
        on[h] = (ch0 + h) * 32 + lane < w2;
        if (!c.dense && on[h]) {
            pv[h] = reinterpret_cast<const V*>(c.param + rowoff)[(ch0 + h) * 32 + lane];
            av[h] = reinterpret_cast<const V*>(c.s0 + rowoff)[(ch0 + h) * 32 + lane];
        }
    }
    const int col0 = ch0 * 32 + lane;
    constexpr int U = NCH <= 2 ? 4 : (NCH <= 4 ? 2 : 1);               // rows in flight (a wide row already has NCH loads in flight)
    int k = 0;
    for (; k + U <= len; k += U) {                                      // U rows in flight, added in slot order
        V v[U][NCH];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const V* p = reinterpret_cast<const V*>(shfl_ptr<T>(rp, k + u));
#pragma unroll
            for (int h = 0; h < NCH; ++h) { v[u][h].x = T(0); v[u][h].y = T(0); if (p && on[h]) v[u][h] = p[col0 + h * 32]; }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int h = 0; h < NCH; ++h) { acc[h].x += v[u][h].x; acc[h].y += v[u][h].y; }
    }
    for (; k < len; ++k) {
        const V* p = reinterpret_cast<const V*>(shfl_ptr<T>(rp, k));
#pragma unroll
        for (int h = 0; h < NCH; ++h) if (p && on[h]) { const V v = p[col0 + h * 32]; acc[h].x += v.x; acc[h].y += v.y; }
    }
#pragma unroll
    for (int h = 0; h < NCH; ++h) {
        if (!on[h]) continue;
        const int cc = col0 + h * 32;
        if (c.dense) reinterpret_cast<V*>(c.dense + rowoff)[cc] = acc[h];
        else {
            adagrad_apply<T>(pv[h].x, acc[h].x, av[h].x, lr, eps); adagrad_apply<T>(pv[h].y, acc[h].y, av[h].y, lr, eps);
            reinterpret_cast<V*>(c.param + rowoff)[cc] = pv[h]; reinterpret_cast<V*>(c.s0 + rowoff)[cc] = av[h];
        }
    }
}

// partial sum of ONE 64-element chunk of a column over the slots slot_at(k), k in [k_lo, k_hi), in that order
template <typename T, typename SlotAt>
__device__ __forceinline__ typename V2<T>::type chunk_partial(const RCol<T>& c, int spr, int ch, int k_lo, int k_hi, int lane, SlotAt slot_at) {
    using V = typename V2<T>::type;
    const bool on = ch * 32 + lane < (c.width >> 1);
    V acc; acc.x = T(0); acc.y = T(0);
    constexpr int U = 8;
    for (int k0 = k_lo; k0 < k_hi; k0 += U) {
        V v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            v[u].x = T(0); v[u].y = T(0);
            if (k0 + u < k_hi) {
                const int s = slot_at(k0 + u);
                const T* p = src_row<T>(c, 0, s, spr);
                if (!p && c.src[1]) p = src_row<T>(c, 1, s, spr);
                if (p && on) v[u] = reinterpret_cast<const V*>(p)[ch * 32 + lane];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) { acc.x += v[u].x; acc.y += v[u].y; }
    }
    return acc;
}

// ---- fast path for the entity-keyed group (compile-time row width 2*W2 elements) ---------------------------------------
// CNT 64-element chunks starting at chunk CH0 of the wide column, over the slots k = 0..len-1 in order; lane k holds the address
// of the k-th contribution row (rp).  Everything about the layout is a compile-time constant or a hoisted register.
template <typename T, int W2, int CH0, int CNT>
__device__ __forceinline__ void fast_pass(T* __restrict__ param, T* __restrict__ st, T* __restrict__ dense, int64_t rowoff, int len,
                                          int lane, const T* rp, T lr, T eps) {
    using V = typename V2<T>::type;
    V acc[CNT], pv[CNT], av[CNT];
#pragma unroll
    for (int h = 0; h < CNT; ++h) {
        acc[h].x = T(0); acc[h].y = T(0);
        const bool on = (CH0 + h) * 32 + 32 <= W2 || lane < W2 - (CH0 + h) * 32;
        if (!dense && on) {
            pv[h] = reinterpret_cast<const V*>(param + rowoff)[(CH0 + h) * 32 + lane];
            av[h] = reinterpret_cast<const V*>(st + rowoff)[(CH0 + h) * 32 + lane];
        }
    }
    constexpr int U = CNT <= 2 ? 8 : (CNT <= 3 ? 4 : 1);              // rows in flight; the last batch is predicated, not serialised
    for (int k = 0; k < len; k += U) {
        V v[U][CNT];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const V* p = reinterpret_cast<const V*>(shfl_ptr<T>(rp, (k + u) & 31)) + CH0 * 32 + lane;
            const bool live = k + u < len;
#pragma unroll
            for (int h = 0; h < CNT; ++h) {
                const bool on = (CH0 + h) * 32 + 32 <= W2 || lane < W2 - (CH0 + h) * 32;
                v[u][h].x = T(0); v[u][h].y = T(0);
                if (live && on) v[u][h] = p[h * 32];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int h = 0; h < CNT; ++h) { acc[h].x += v[u][h].x; acc[h].y += v[u][h].y; }
    }
#pragma unroll
    for (int h = 0; h < CNT; ++h) {
        const bool on = (CH0 + h) * 32 + 32 <= W2 || lane < W2 - (CH0 + h) * 32;
        if (!on) continue;
        const int cc = (CH0 + h) * 32 + lane;
        if (dense) reinterpret_cast<V*>(dense + rowoff)[cc] = acc[h];
        else {
            adagrad_apply<T>(pv[h].x, acc[h].x, av[h].x, lr, eps); adagrad_apply<T>(pv[h].y, acc[h].y, av[h].y, lr, eps);
            reinterpret_cast<V*>(param + rowoff)[cc] = pv[h]; reinterpret_cast<V*>(st + rowoff)[cc] = av[h];
        }
    }
}

// Short segments of group 0 when its first column is 2*W2 wide with every slot feeding it (sources 0 / 1 partition the slots)
// and the other columns are scalar: one warp per segment, the whole row in one item.
template <typename T, int W2>
__device__ __forceinline__ void fast_entity_phase1(const RGroup<T>& G, T lr, T eps, int lane, int gwarp, int nwarps) {
    constexpr int NCH = (W2 + 31) / 32;
    const RCol<T>& c0 = G.col[0];
    const T* const src0 = c0.src[0]; const T* const src1 = c0.src[1];
    const int hi0 = c0.hi[0], lo1 = c0.lo[1];
    const int64_t rs0 = c0.rstride[0], rs1 = c0.rstride[1];
    T* const param = c0.param; T* const st = c0.s0; T* const dense = c0.dense;
    const int spr = G.slots_per_rank, nsc = G.n_cols - 1;
    const int nshort = G.v.hdr[0];
    auto meta = [&](int sg, int& id, int& len, int& base) {
        id = 0; len = 0; base = 0;
        if (sg < nshort) { id = G.v.seg[sg]; len = G.v.slen[sg]; base = G.v.sbase[sg]; }
    };
    int id0, len0, base0, id1, len1, base1, mine0 = 0x7fffffff, mine1;
    meta(gwarp, id0, len0, base0);
    meta(gwarp + nwarps, id1, len1, base1);
    if (lane < len0) mine0 = G.v.order[base0 + lane];
    for (int sg = gwarp; sg < nshort; sg += nwarps) {
        int id2, len2, base2;
        meta(sg + 2 * nwarps, id2, len2, base2);
        mine1 = 0x7fffffff;
        if (lane < len1) mine1 = G.v.order[base1 + lane];
        const int id = id0, len = len0;
        const int mine = warp_sort_len(mine0, lane, len);
        int rk = 0, ls = mine;                                          // rank-major slot numbering (data parallel)
        if (lane < len && mine >= spr) { rk = mine / spr; ls = mine - rk * spr; }
        const T* rp = nullptr;
        if (lane < len) rp = ls < hi0 ? src0 + rk * rs0 + (int64_t)ls * (2 * W2) : src1 + rk * rs1 + (int64_t)(ls - lo1) * (2 * W2);
        T sp[2], sa[2], sv[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            sp[j] = T(0); sa[j] = T(0); sv[j] = T(0);
            if (j < nsc) {
                const RCol<T>& c = G.col[1 + j];
                if (lane == 0 && !c.dense) { sp[j] = c.param[id]; sa[j] = c.s0[id]; }
                if (lane < len && ls >= c.lo[0] && ls < c.hi[0]) sv[j] = c.src[0][rk * c.rstride[0] + (ls - c.lo[0])];
            }
        }
        const int64_t rowoff = (int64_t)id * (2 * W2);
        if constexpr (NCH <= 5) fast_pass<T, W2, 0, (NCH <= 5 ? NCH : 1)>(param, st, dense, rowoff, len, lane, rp, lr, eps);
        else {
            fast_pass<T, W2, 0, 5>(param, st, dense, rowoff, len, lane, rp, lr, eps);
            fast_pass<T, W2, 5, (NCH > 5 ? NCH - 5 : 1)>(param, st, dense, rowoff, len, lane, rp, lr, eps);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (j < nsc) {
                const RCol<T>& c = G.col[1 + j];
                const T g = warp_sum<T>(sv[j]);                         // fixed butterfly over the sorted positions: deterministic
                if (lane == 0) {
                    if (c.dense) c.dense[id] = g;
                    else { T pp = sp[j], aa = sa[j]; adagrad_apply<T>(pp, g, aa, lr, eps); c.param[id] = pp; c.s0[id] = aa; }
                }
            }
        }
        id0 = id1; len0 = len1; base0 = base1; mine0 = mine1;
        id1 = id2; len1 = len2; base1 = base2;
    }
}


// ---- computed-source path of the entity-keyed group ----------------------------------------------------------------
// The tail-row gradient of a (query, tail) pair is rebuilt from the pair's three scalars, the query row z and the tail row w
// the warp is about to update, with the operations K3 would have used to store it (bit-identical to the stored-row path):
//     g_re[k] = fma(c1, z_re[k], fma(c2, z_im[k], -c3 * w_re[k])),   g_im[k] = fma(c1, z_im[k], fma(-c2, z_re[k], -c3 * w_im[k]))
// so a pair costs 16 bytes (+ an L2-resident query row) instead of a 2R-wide row written by K3, read here and — data parallel —
// all_gathered.  Lane l owns the complex coefficients k = 32 c + l of chunk c: elements k and R + k of the row.
template <typename T> struct CoefSrc {
    const T* src0; const T* qsrc; const T* coef; int64_t rs0, rsq, cs; int hi0, lo1, nt, spr;
};
template <typename T> struct V4;
template <> struct V4<float> { using type = float4; };
template <> struct V4<double> { using type = double4; };

// descriptor of slot s: stored row (head-entity gradient of the K1 adjoint) or pair (query row + coefficients)
template <typename T, int R>
__device__ __forceinline__ void coef_desc(const CoefSrc<T>& C, int s, bool live, const T*& dp, T& d1, T& d2, T& d3, bool& pair) {
    dp = nullptr; d1 = T(0); d2 = T(0); d3 = T(0); pair = false;
    if (!live) return;
    int rk = 0, ls = s;
    if (s >= C.spr) { rk = s / C.spr; ls = s - rk * C.spr; }
    if (ls < C.hi0) { dp = C.src0 + rk * C.rs0 + (int64_t)ls * (2 * R); return; }
    const int pi = ls - C.lo1;
    const int qi = C.nt ? pi / C.nt : pi;
    dp = C.qsrc + rk * C.rsq + (int64_t)qi * (2 * R);
    const T* cf = C.coef + rk * C.cs + (int64_t)pi * 4;
    if constexpr (sizeof(T) == 4) { const float4 v = *reinterpret_cast<const float4*>(cf); d1 = v.x; d2 = v.y; d3 = v.z; }
    else { const double2 a = reinterpret_cast<const double2*>(cf)[0]; d1 = a.x; d2 = a.y; d3 = cf[2]; }
    pair = true;
}
template <typename T>
__device__ __forceinline__ T shfl_val(T v, int k) {
    if constexpr (sizeof(T) == 4) return __shfl_sync(CHK_FULL, v, k);
    else {
        const unsigned long long b = (unsigned long long)__double_as_longlong(v);
        const unsigned lo = __shfl_sync(CHK_FULL, (unsigned)b, k), hi = __shfl_sync(CHK_FULL, (unsigned)(b >> 32), k);
        return __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
    }
}
template <int R, int C0, int H>
__device__ __forceinline__ bool coef_on(int lane) { return (C0 + H) * 32 + 32 <= R || lane < R - (C0 + H) * 32; }

// adds the contributions of the (<= 32) slots described lane-wise by (dp, d1, d2, d3, pairmask) to (ar, ai), in lane order,
// for complex chunks C0 .. C0+CNT-1; (wr, wi) = the row's current values of those coefficients
template <typename T, int R, int C0, int CNT>
__device__ __forceinline__ void coef_accumulate(T (&ar)[CNT], T (&ai)[CNT], const T (&wr)[CNT], const T (&wi)[CNT], int len, int lane,
                                                const T* dp, T d1, T d2, T d3, unsigned pairmask) {
    constexpr int U = CNT <= 2 ? 4 : (CNT <= 3 ? 2 : 1);                 // slots in flight
    for (int k = 0; k < len; k += U) {
        T zr[U][CNT], zi[U][CNT], c1[U], c2[U], c3[U];
        bool pr[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int kk = (k + u) & 31;
            const T* p = shfl_ptr<T>(dp, kk);
            c1[u] = shfl_val<T>(d1, kk); c2[u] = shfl_val<T>(d2, kk); c3[u] = shfl_val<T>(d3, kk);
            pr[u] = (pairmask >> kk) & 1u;
            const bool live = k + u < len;
#pragma unroll
            for (int h = 0; h < CNT; ++h) {
                const bool on = live && ((C0 + h) * 32 + 32 <= R || lane < R - (C0 + h) * 32);
                zr[u][h] = T(0); zi[u][h] = T(0);
                if (on) { zr[u][h] = p[(C0 + h) * 32 + lane]; zi[u][h] = p[R + (C0 + h) * 32 + lane]; }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int h = 0; h < CNT; ++h) {
                const T gr = Sc<T>::fma_(c1[u], zr[u][h], Sc<T>::fma_(c2[u], zi[u][h], -c3[u] * wr[h]));
                const T gi = Sc<T>::fma_(c1[u], zi[u][h], Sc<T>::fma_(-c2[u], zr[u][h], -c3[u] * wi[h]));
                ar[h] += pr[u] ? gr : zr[u][h];
                ai[h] += pr[u] ? gi : zi[u][h];
            }
    }
}
template <typename T, int R, int C0, int CNT>
__device__ __forceinline__ void coef_load_row(const T* __restrict__ base, int64_t rowoff, int lane, T (&xr)[CNT], T (&xi)[CNT]) {
#pragma unroll
    for (int h = 0; h < CNT; ++h) {
        const bool on = (C0 + h) * 32 + 32 <= R || lane < R - (C0 + h) * 32;
        xr[h] = T(0); xi[h] = T(0);
        if (on) { xr[h] = base[rowoff + (C0 + h) * 32 + lane]; xi[h] = base[rowoff + R + (C0 + h) * 32 + lane]; }
    }
}
template <typename T, int R, int C0, int CNT>
__device__ __forceinline__ void coef_finish(T* __restrict__ param, T* __restrict__ st, T* __restrict__ dense, int64_t rowoff, int lane,
                                            T (&ar)[CNT], T (&ai)[CNT], T (&wr)[CNT], T (&wi)[CNT], T (&sr)[CNT], T (&si)[CNT], T lr, T eps) {
#pragma unroll
    for (int h = 0; h < CNT; ++h) {
        const bool on = (C0 + h) * 32 + 32 <= R || lane < R - (C0 + h) * 32;
        if (!on) continue;
        const int64_t e = rowoff + (C0 + h) * 32 + lane;
        if (dense) { dense[e] = ar[h]; dense[e + R] = ai[h]; }
        else {
            adagrad_apply<T>(wr[h], ar[h], sr[h], lr, eps); adagrad_apply<T>(wi[h], ai[h], si[h], lr, eps);
            param[e] = wr[h]; param[e + R] = wi[h]; st[e] = sr[h]; st[e + R] = si[h];
        }
    }
}
// one warp, one short segment, chunks C0 .. C0+CNT-1
template <typename T, int R, int C0, int CNT>
__device__ __forceinline__ void coef_pass(T* __restrict__ param, T* __restrict__ st, T* __restrict__ dense, int64_t rowoff, int len, int lane,
                                          const T* dp, T d1, T d2, T d3, unsigned pairmask, T lr, T eps) {
    T ar[CNT], ai[CNT], wr[CNT], wi[CNT], sr[CNT], si[CNT];
    coef_load_row<T, R, C0, CNT>(param, rowoff, lane, wr, wi);
    if (!dense) coef_load_row<T, R, C0, CNT>(st, rowoff, lane, sr, si);
#pragma unroll
    for (int h = 0; h < CNT; ++h) { ar[h] = T(0); ai[h] = T(0); }
    coef_accumulate<T, R, C0, CNT>(ar, ai, wr, wi, len, lane, dp, d1, d2, d3, pairmask);
    coef_finish<T, R, C0, CNT>(param, st, dense, rowoff, lane, ar, ai, wr, wi, sr, si, lr, eps);
}

template <typename T, int R>
__device__ __forceinline__ CoefSrc<T> coef_src_of(const RGroup<T>& G) {
    const RCol<T>& c0 = G.col[0];
    CoefSrc<T> C;
    C.src0 = c0.src[0]; C.qsrc = c0.src[1]; C.coef = c0.coef; C.rs0 = c0.rstride[0]; C.rsq = c0.rstride[1]; C.cs = c0.cstride;
    C.hi0 = c0.hi[0]; C.lo1 = c0.lo[1]; C.nt = c0.pair_nt; C.spr = G.slots_per_rank;
    return C;
}

// short segments (<= 32 slots) of the entity group, one warp per segment (same walk as fast_entity_phase1)
template <typename T, int R>
__device__ __forceinline__ void coef_entity_phase1(const RGroup<T>& G, T lr, T eps, int lane, int gwarp, int nwarps) {
    constexpr int NCH = (R + 31) / 32;
    const RCol<T>& c0 = G.col[0];
    const CoefSrc<T> C = coef_src_of<T, R>(G);
    T* const param = c0.param; T* const st = c0.s0; T* const dense = c0.dense;
    const int spr = G.slots_per_rank, nsc = G.n_cols - 1;
    const int nshort = G.v.hdr[0];
    auto meta = [&](int sg, int& id, int& len, int& base) {
        id = 0; len = 0; base = 0;
        if (sg < nshort) { id = G.v.seg[sg]; len = G.v.slen[sg]; base = G.v.sbase[sg]; }
    };
    int id0, len0, base0, id1, len1, base1, mine0 = 0x7fffffff, mine1;
    meta(gwarp, id0, len0, base0);
    meta(gwarp + nwarps, id1, len1, base1);
    if (lane < len0) mine0 = G.v.order[base0 + lane];
    for (int sg = gwarp; sg < nshort; sg += nwarps) {
        int id2, len2, base2;
        meta(sg + 2 * nwarps, id2, len2, base2);
        mine1 = 0x7fffffff;
        if (lane < len1) mine1 = G.v.order[base1 + lane];
        const int id = id0, len = len0;
        const int mine = warp_sort_len(mine0, lane, len);
        const T* dp; T d1, d2, d3; bool pair;
        coef_desc<T, R>(C, mine, lane < len, dp, d1, d2, d3, pair);
        const unsigned pairmask = __ballot_sync(CHK_FULL, pair);
        int rk = 0, ls = mine;
        if (lane < len && mine >= spr) { rk = mine / spr; ls = mine - rk * spr; }
        T sp[2], sa[2], sv[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            sp[j] = T(0); sa[j] = T(0); sv[j] = T(0);
            if (j < nsc) {
                const RCol<T>& c = G.col[1 + j];
                if (lane == 0 && !c.dense) { sp[j] = c.param[id]; sa[j] = c.s0[id]; }
                if (lane < len && ls >= c.lo[0] && ls < c.hi[0]) sv[j] = c.src[0][rk * c.rstride[0] + (ls - c.lo[0])];
            }
        }
        const int64_t rowoff = (int64_t)id * (2 * R);
        if constexpr (NCH <= 5 || (CHK_COEF_ONEPASS && sizeof(T) == 4)) coef_pass<T, R, 0, NCH>(param, st, dense, rowoff, len, lane, dp, d1, d2, d3, pairmask, lr, eps);
        else {
            coef_pass<T, R, 0, 5>(param, st, dense, rowoff, len, lane, dp, d1, d2, d3, pairmask, lr, eps);
            coef_pass<T, R, 5, (NCH > 5 ? NCH - 5 : 1)>(param, st, dense, rowoff, len, lane, dp, d1, d2, d3, pairmask, lr, eps);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (j < nsc) {
                const RCol<T>& c = G.col[1 + j];
                const T g = warp_sum<T>(sv[j]);
                if (lane == 0) {
                    if (c.dense) c.dense[id] = g;
                    else { T pp = sp[j], aa = sa[j]; adagrad_apply<T>(pp, g, aa, lr, eps); c.param[id] = pp; c.s0[id] = aa; }
                }
            }
        }
        id0 = id1; len0 = len1; base0 = base1; mine0 = mine1;
        id1 = id2; len1 = len2; base1 = base2;
    }
}

// long segment (> 32 slots) of the entity group, one CTA: every warp sums a contiguous eighth of the sorted slots (32 at a
// time through the lane-wise descriptors), warp 0 adds the eight partials in warp order and applies the update — the same
// order of additions as the stored-row path.  `cpart` = [RWARPS][2 * CMAX][32] partials.
template <typename T, int R, int C0, int CNT>
__device__ __forceinline__ void coef_long_pass(const CoefSrc<T>& C, T* __restrict__ param, T* __restrict__ st, T* __restrict__ dense, int64_t rowoff,
                                               const int* sorted, int k_lo, int k_hi, int lane, int warp, T* cpart, T lr, T eps) {
    T ar[CNT], ai[CNT], wr[CNT], wi[CNT], sr[CNT], si[CNT];
    coef_load_row<T, R, C0, CNT>(param, rowoff, lane, wr, wi);
#pragma unroll
    for (int h = 0; h < CNT; ++h) { ar[h] = T(0); ai[h] = T(0); }
    for (int k0 = k_lo; k0 < k_hi; k0 += 32) {
        const int n = min(32, k_hi - k0);
        const T* dp; T d1, d2, d3; bool pair;
        coef_desc<T, R>(C, lane < n ? sorted[k0 + lane] : 0, lane < n, dp, d1, d2, d3, pair);
        const unsigned pairmask = __ballot_sync(CHK_FULL, pair);
        coef_accumulate<T, R, C0, CNT>(ar, ai, wr, wi, n, lane, dp, d1, d2, d3, pairmask);
    }
    __syncthreads();                                                       // cpart of the previous pass / segment is free
#pragma unroll
    for (int h = 0; h < CNT; ++h) { cpart[(warp * 2 * CNT + 2 * h) * 32 + lane] = ar[h]; cpart[(warp * 2 * CNT + 2 * h + 1) * 32 + lane] = ai[h]; }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int h = 0; h < CNT; ++h) {
            T a = cpart[(2 * h) * 32 + lane], b = cpart[(2 * h + 1) * 32 + lane];
#pragma unroll
            for (int w = 1; w < RWARPS; ++w) { a += cpart[(w * 2 * CNT + 2 * h) * 32 + lane]; b += cpart[(w * 2 * CNT + 2 * h + 1) * 32 + lane]; }
            ar[h] = a; ai[h] = b;
        }
        if (!dense) coef_load_row<T, R, C0, CNT>(st, rowoff, lane, sr, si);
        coef_finish<T, R, C0, CNT>(param, st, dense, rowoff, lane, ar, ai, wr, wi, sr, si, lr, eps);
    }
}
template <typename T, int R>
__device__ __forceinline__ void coef_long_segment(const RGroup<T>& G, int id, int len, const int* sorted, int lane, int warp, T* cpart, T lr, T eps) {
    constexpr int NCH = (R + 31) / 32;
    const RCol<T>& c0 = G.col[0];
    const CoefSrc<T> C = coef_src_of<T, R>(G);
    const int per = (len + RWARPS - 1) / RWARPS;
    const int k_lo = min(warp * per, len), k_hi = min(k_lo + per, len);
    const int64_t rowoff = (int64_t)id * (2 * R);
    if constexpr (NCH <= 5) coef_long_pass<T, R, 0, (NCH <= 5 ? NCH : 1)>(C, c0.param, c0.s0, c0.dense, rowoff, sorted, k_lo, k_hi, lane, warp, cpart, lr, eps);
    else {
        coef_long_pass<T, R, 0, 5>(C, c0.param, c0.s0, c0.dense, rowoff, sorted, k_lo, k_hi, lane, warp, cpart, lr, eps);
        coef_long_pass<T, R, 5, (NCH > 5 ? NCH - 5 : 1)>(C, c0.param, c0.s0, c0.dense, rowoff, sorted, k_lo, k_hi, lane, warp, cpart, lr, eps);
    }
}

template <typename T, int W2, bool COEF>
__global__ void __launch_bounds__(RWARPS * 32, (W2 > 0 && W2 <= 65) || W2 == 0 ? 4 : ((CHK_COEF_ONEPASS && COEF && sizeof(T) == 4) ? 2 : 3)) reduce_apply_kernel(const RArgsStep<T> A) {
    using V = typename V2<T>::type;
    constexpr int MAXCH = 2;
    __shared__ int sbuf[SORT_CAP];
    __shared__ V part[RWARPS][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const T lr = (T)A.hyper[0], eps = (T)A.hyper[1];
    for (int gi = 0; gi < A.n_groups; ++gi) {
        const RGroup<T>& G = A.g[gi];
        const int spr = G.slots_per_rank;
        // CTA roles: the last `nlc` CTAs of the grid take this group's long segments (and nothing else of the group), so the
        // CTA-wide passes run BESIDE the short segments instead of after them; tiny grids do both in turn.
        const int nlong = G.single ? 1 : G.v.hdr[2];
        const int nlc = min(nlong, (int)gridDim.x >> 2);
        const int first_long = (int)gridDim.x - nlc;
        const bool long_cta = nlc > 0 && (int)blockIdx.x >= first_long;
        const int gwarp = long_cta ? 0x3fffffff : blockIdx.x * RWARPS + warp, nwarps = first_long * RWARPS;
        // ---- phase 1: short segments (<= 32 slots).  Work item = (segment, chunk group): one warp sorts the segment's slot
        //      list in registers and reduces MAXCH 64-element chunks of every column (wide rows are split over several warps;
        //      the scalar columns ride with chunk group 0).  The item's metadata (row id, length, start, slot list) is loaded
        //      one and two items ahead, so the only exposed latency per item is the one of the rows themselves.
        int ncg = 1;                                                   // chunk groups per segment = widest column's
        for (int ci = 0; ci < G.n_cols; ++ci) {
            const int nch = ((G.col[ci].width >> 1) + 31) >> 5;
            ncg = max(ncg, (nch + MAXCH - 1) / MAXCH);
        }
        int nshort = G.single ? 0 : G.v.hdr[0];
        if constexpr (W2 > 0) {                                        // compile-time-width fast path for the entity-keyed group
            if (gi == 0) {
                if constexpr (COEF) coef_entity_phase1<T, W2>(G, lr, eps, lane, gwarp, nwarps);
                else fast_entity_phase1<T, W2>(G, lr, eps, lane, gwarp, nwarps);
                nshort = 0;
            }
        }
        const int nitems = nshort * ncg;
        auto meta = [&](int it, int& id, int& len, int& base) {
            id = 0; len = 0; base = 0;
            if (it < nitems) { const int sg = it / ncg; id = G.v.seg[sg]; len = G.v.slen[sg]; base = G.v.sbase[sg]; }
        };
        int id0, len0, base0, id1, len1, base1, mine0 = 0x7fffffff, mine1 = 0x7fffffff;
        meta(gwarp, id0, len0, base0);
        meta(gwarp + nwarps, id1, len1, base1);
        if (lane < len0) mine0 = G.v.order[base0 + lane];
        for (int it = gwarp; it < nitems; it += nwarps) {
            int id2, len2, base2;
            meta(it + 2 * nwarps, id2, len2, base2);                   // two ahead: metadata
            mine1 = 0x7fffffff;
            if (lane < len1) mine1 = G.v.order[base1 + lane];          // one ahead: its slot list
            const int cg = it % ncg, id = id0, len = len0;
            const int mine = warp_sort_len(mine0, lane, len);          // lane k: the k-th smallest slot of the segment
            // scalar columns first (chunk group 0): their parameter / state loads are issued before the row traffic
            T sp[CHK_RED_MAX_COLS], sa[CHK_RED_MAX_COLS], sv[CHK_RED_MAX_COLS];
            if (cg == 0) {
#pragma unroll
                for (int ci = 0; ci < CHK_RED_MAX_COLS; ++ci) {
                    sp[ci] = T(0); sa[ci] = T(0); sv[ci] = T(0);
                    if (ci < G.n_cols && G.col[ci].width == 1) {
                        const RCol<T>& c = G.col[ci];
                        if (lane == 0 && !c.dense) { sp[ci] = c.param[id]; sa[ci] = c.s0[id]; }
                        if (lane < len) {
                            const T* p = src_row<T>(c, 0, mine, spr);
                            if (!p && c.src[1]) p = src_row<T>(c, 1, mine, spr);
                            if (p) sv[ci] = *p;
                        }
                    }
                }
            }
            for (int ci = 0; ci < G.n_cols; ++ci) {
                const RCol<T>& c = G.col[ci];
                if (c.width == 1) continue;
                const int nch = ((c.width >> 1) + 31) >> 5;
                int ch = cg * MAXCH;
                const int ch_end = min(nch, ch + MAXCH);
                if (ch >= nch) continue;
                const T* rp = nullptr;                                 // lane k: row of the k-th slot in this column's sources
                if (lane < len) { rp = src_row<T>(c, 0, mine, spr); if (!rp && c.src[1]) rp = src_row<T>(c, 1, mine, spr); }
                while (ch < ch_end) {
                    const int left = ch_end - ch;
                    if (left >= 2) { col_chunks_ptr<T, 2>(c, id, len, ch, lr, eps, lane, rp); ch += 2; }
                    else { col_chunks_ptr<T, 1>(c, id, len, ch, lr, eps, lane, rp); ch += 1; }
                }
            }
            if (cg == 0) {
#pragma unroll
                for (int ci = 0; ci < CHK_RED_MAX_COLS; ++ci) {
                    if (ci < G.n_cols && G.col[ci].width == 1) {
                        const RCol<T>& c = G.col[ci];
                        const T g = warp_sum<T>(sv[ci]);               // fixed butterfly over the sorted positions: deterministic
                        if (lane == 0) {
                            if (c.dense) c.dense[id] = g;
                            else { T pp = sp[ci], aa = sa[ci]; adagrad_apply<T>(pp, g, aa, lr, eps); c.param[id] = pp; c.s0[id] = aa; }
                        }
                    }
                }
            }
            id0 = id1; len0 = len1; base0 = base1; mine0 = mine1;
            id1 = id2; len1 = len2; base1 = base2;
        }
        // ---- phase 2: long segments (and the single-row group), one CTA each: CTA-wide sort of the slot list; per 64-element
        //      chunk every warp sums a contiguous eighth of the slots and warp 0 adds the eight partials in warp order
        //      (fixed order: deterministic) and applies the update
        const int li0 = nlc > 0 ? (long_cta ? (int)blockIdx.x - first_long : nlong) : (int)blockIdx.x;
        const int li_step = nlc > 0 ? nlc : (int)gridDim.x;
        for (int li = li0; li < nlong; li += li_step) {
            int id = 0, len = G.total, base = 0;
            const int* sorted = nullptr;                               // nullptr: identity order (single-row group)
            __syncthreads();                                           // sbuf / part of the previous segment are free
            if (!G.single) {
                const int at = G.total - 1 - li;
                id = G.v.seg[at]; len = G.v.slen[at]; base = G.v.sbase[at];
                int* o = G.v.order + base;
                if (len <= SORT_CAP) {
                    for (int i = threadIdx.x; i < len; i += blockDim.x) sbuf[i] = o[i];
                    __syncthreads();
                    cta_sort(sbuf, len);
                    sorted = sbuf;
                } else {
                    __syncthreads();
                    cta_sort(o, len);                                  // pathological: > SORT_CAP slots name one row
                    sorted = o;
                }
            }
            auto slot_at = [&](int k) -> int { return sorted ? sorted[k] : k; };
            const int per = (len + RWARPS - 1) / RWARPS;
            const int k_lo = min(warp * per, len), k_hi = min(k_lo + per, len);
            for (int ci = 0; ci < G.n_cols; ++ci) {
                const RCol<T>& c = G.col[ci];
                if (c.width == 1) { if (warp == 0) col_scalar<T>(c, id, len, spr, lr, eps, lane, slot_at); continue; }
                if constexpr (COEF && W2 > 0) {
                    if (gi == 0 && ci == 0) {                              // pairs are rebuilt from their coefficients
                        __shared__ T cpart[RWARPS * 2 * 5 * 32];
                        coef_long_segment<T, W2>(G, id, len, sorted, lane, warp, cpart, lr, eps);
                        continue;
                    }
                }
                const int w2 = c.width >> 1, nch = (w2 + 31) >> 5;
                const int64_t rowoff = (int64_t)id * c.width;
                for (int ch = 0; ch < nch; ++ch) {
                    part[warp][lane] = chunk_partial<T>(c, spr, ch, k_lo, k_hi, lane, slot_at);
                    __syncthreads();
                    if (warp == 0 && ch * 32 + lane < w2) {
                        V acc = part[0][lane];
#pragma unroll
                        for (int w = 1; w < RWARPS; ++w) { acc.x += part[w][lane].x; acc.y += part[w][lane].y; }
                        const int cc = ch * 32 + lane;
                        if (c.dense) reinterpret_cast<V*>(c.dense + rowoff)[cc] = acc;
                        else {
                            V pv = reinterpret_cast<V*>(c.param + rowoff)[cc], av = reinterpret_cast<V*>(c.s0 + rowoff)[cc];
                            adagrad_apply<T>(pv.x, acc.x, av.x, lr, eps); adagrad_apply<T>(pv.y, acc.y, av.y, lr, eps);
                            reinterpret_cast<V*>(c.param + rowoff)[cc] = pv; reinterpret_cast<V*>(c.s0 + rowoff)[cc] = av;
                        }
                    }
                    __syncthreads();
                }
            }
        }
    }
    if (!A.finish) return;
    // ---- end of step: the block that finishes last resets the grouping headers, adds the loss partials in a fixed order
    //      and bumps the step counter (every other block is past its last read of them)
    __shared__ int last_flag;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last_flag = atomicAdd(A.ticket, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!last_flag) return;
    if (threadIdx.x < A.n_groups && !A.g[threadIdx.x].single) {
        int* h = A.g[threadIdx.x].v.hdr;
        h[0] = 0; h[1] = 0; h[2] = 0;
    }
    if (threadIdx.x == 0) *A.ticket = 0;
    if (A.loss_part) {
        T acc = T(0);
        for (int64_t i = threadIdx.x; i < A.n_loss; i += blockDim.x) acc += A.loss_part[i];
        acc = warp_sum<T>(acc);
        T* red = reinterpret_cast<T*>(sbuf);
        __syncthreads();
        if (lane == 0) red[warp] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            T t = T(0);
            for (int w = 0; w < RWARPS; ++w) t += red[w];
            *A.loss_accum += t;
        }
    }
    if (threadIdx.x == 0 && A.step_id) *A.step_id += 1;
}

// reset the grouping headers (nseg, cursor) after the reduce of a step
struct HdrList { int* h[CHK_RED_MAX_GROUPS]; int n; };
template <typename T>
__global__ void __launch_bounds__(256) step_finish_kernel(HdrList H, const T* __restrict__ loss_part, int64_t n_loss, T* __restrict__ loss_accum,
                                                          int* __restrict__ step_id) {
    __shared__ T red[8];
    if (threadIdx.x < H.n) { H.h[threadIdx.x][0] = 0; H.h[threadIdx.x][1] = 0; H.h[threadIdx.x][2] = 0; }
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

// out[b, c] = sum_j in[b, j, c].  One CTA per (b, 32-column chunk): warp w adds the rows j = w, w + 8, ... (each a coalesced
// 128-byte read, four in flight), the eight partial sums are combined in warp order — a fixed order, so the result is
// bit-reproducible.  (r2: the first version ran one thread per (b, c) over all j serially: 20 us per table at B = 500,
// nt = 101 — three of them were a fifth of the double_neg step.)
template <typename T>
__global__ void __launch_bounds__(256) rowsum_groups_kernel(const T* __restrict__ in, int64_t B, int64_t nj, int64_t width, T* __restrict__ out) {
    __shared__ T part[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t chunks = (width + 31) / 32;
    for (int64_t item = blockIdx.x; item < B * chunks; item += gridDim.x) {
        const int64_t b = item / chunks, c = (item - b * chunks) * 32 + lane;
        const bool on = c < width;
        const T* p = in + b * nj * width + c;
        T acc = T(0);
        int64_t j = warp;
        for (; j + 24 < nj; j += 32) {
            T v0 = T(0), v1 = T(0), v2 = T(0), v3 = T(0);
            if (on) { v0 = p[j * width]; v1 = p[(j + 8) * width]; v2 = p[(j + 16) * width]; v3 = p[(j + 24) * width]; }
            acc += v0; acc += v1; acc += v2; acc += v3;
        }
        for (; j < nj; j += 8) if (on) acc += p[j * width];
        __syncthreads();                                               // part of the previous item has been read
        part[warp][lane] = acc;
        __syncthreads();
        if (warp == 0 && on) {
            T s = part[0][lane];
#pragma unroll
            for (int w = 1; w < 8; ++w) s += part[w][lane];
            out[b * width + c] = s;
        }
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
    return (int64_t)sizeof(int) * (4 + 2 * n_keys + 5 * total_slots);
}

extern "C" int chk_group_build(const int64_t* ids, int64_t total_slots, int64_t n_keys, int64_t own_lo, int64_t own_hi, void* work, void* stream) {
    if (total_slots == 0) return CHK_OK;
    if (!ids || !work || chk_group_workspace_bytes(n_keys, total_slots) < 0 || own_lo < 0 || own_hi > n_keys) { chk_set_error("chk_group_build: bad argument"); return CHK_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    GroupView v = group_view(work, n_keys, total_slots);
    const int grid = grid_for(total_slots, 256, 148 * 8);
    group_count_kernel<<<grid, 256, 0, st>>>(ids, total_slots, v, own_lo, own_hi);
    group_alloc_kernel<<<grid, 256, 0, st>>>(ids, total_slots, v);
    group_order_kernel<<<grid, 256, 0, st>>>(ids, total_slots, v);
    CHK_CUDA_LAUNCH_CHECK("group kernels");
    return CHK_OK;
}

template <typename T>
static int reduce_apply_t(int opt, const chk_red_group* groups, int n_groups, const double* hyper, int finish, const void* loss_part,
                          int64_t n_loss, void* loss_accum, int32_t* step_id, cudaStream_t st) {
    RArgsStep<T> A{};
    A.n_groups = n_groups; A.opt = opt; A.hyper = hyper;
    A.finish = finish; A.loss_part = (const T*)loss_part; A.n_loss = n_loss; A.loss_accum = (T*)loss_accum; A.step_id = step_id;
    if (finish) {
        for (int gi = 0; gi < n_groups && !A.ticket; ++gi) if (!groups[gi].single_row && groups[gi].work) A.ticket = (int*)groups[gi].work + 3;
        if (!A.ticket || (loss_part && !loss_accum)) { chk_set_error("chk_reduce_apply: finish needs a grouped group (ticket) and loss_accum with loss_part"); return CHK_EINVAL; }
    }
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
            C.coef = (const T*)c.pair_coef; C.pair_nt = (int)c.pair_nt; C.cstride = c.coef_rank_stride;
            if (c.pair_coef && (gi != 0 || ci != 0 || c.pair_nt < 0)) { chk_set_error("chk_reduce_apply: pair_coef is for the first column of group 0"); return CHK_EINVAL; }
        }
        if (!g.single_row && g.n_keys * 0 + total > max_seg) max_seg = total;
    }
    // one resident wave; every CTA walks all groups
    // one resident wave; every CTA walks all groups.  Group 0 takes the compile-time-width fast path when it has the entity-group
    // shape: first column 2*W2 wide fed by every slot (two sources partitioning the slots), the other columns scalar.
    static const char* force = getenv("CHK_REDUCE_VARIANT");          // measurement override: "generic"
    int w2 = 0;
    {
        const chk_red_group& g0 = groups[0];
        const chk_red_col& c0 = g0.cols[0];
        bool ok = !g0.single_row && c0.width >= 2 && c0.src[0] && c0.src[1] && c0.lo[0] == 0 && c0.hi[0] == c0.lo[1] && c0.hi[1] == g0.slots_per_rank &&
                  g0.n_cols <= 3 && !(force && force[0] == 'g');
        for (int ci = 1; ci < g0.n_cols && ok; ++ci) ok = g0.cols[ci].width == 1 && g0.cols[ci].src[0] && !g0.cols[ci].src[1];
        if (ok) w2 = (int)(c0.width >> 1);
    }
    const bool coef = groups[0].cols[0].pair_coef != nullptr;
    if (coef && !(w2 == 9 || w2 == 17 || w2 == 33 || w2 == 65 || w2 == 129 || w2 == 257)) {
        chk_set_error("chk_reduce_apply: pair_coef needs the entity-group shape and a rank in {9,17,33,65,129,257}"); return CHK_EINVAL;
    }
    const int grid4 = grid_for(max_seg, RWARPS, 148 * 4);
    const int grid3 = grid_for(max_seg, RWARPS, 148 * ((CHK_COEF_ONEPASS && coef && sizeof(T) == 4) ? 2 : 3));   // one resident wave
#define CHK_RED_LAUNCH(W, G)                                                                                   \
    case W: if (coef) reduce_apply_kernel<T, W, true><<<G, RWARPS * 32, 0, st>>>(A);                          \
            else reduce_apply_kernel<T, W, false><<<G, RWARPS * 32, 0, st>>>(A);                              \
            break;
    switch (w2) {
        CHK_RED_LAUNCH(9, grid4) CHK_RED_LAUNCH(17, grid4) CHK_RED_LAUNCH(33, grid4) CHK_RED_LAUNCH(65, grid4)
        CHK_RED_LAUNCH(129, grid3) CHK_RED_LAUNCH(257, grid3)
        default: reduce_apply_kernel<T, 0, false><<<grid4, RWARPS * 32, 0, st>>>(A); break;
    }
#undef CHK_RED_LAUNCH
    CHK_CUDA_LAUNCH_CHECK("reduce_apply_kernel");
    return CHK_OK;
}

extern "C" int chk_reduce_apply(int dtype, int opt, const chk_red_group* groups, int n_groups, const double* hyper,
                                int finish_step, const void* loss_part, int64_t n_loss, void* loss_accum, int32_t* step_id, void* stream) {
    if (n_groups < 1 || n_groups > CHK_RED_MAX_GROUPS || !groups || !hyper) { chk_set_error("chk_reduce_apply: bad argument"); return CHK_EINVAL; }
    if (dtype == CHK_F32) return reduce_apply_t<float>(opt, groups, n_groups, hyper, finish_step, loss_part, n_loss, loss_accum, step_id, (cudaStream_t)stream);
    if (dtype == CHK_F64) return reduce_apply_t<double>(opt, groups, n_groups, hyper, finish_step, loss_part, n_loss, loss_accum, step_id, (cudaStream_t)stream);
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
    const int grid = grid_for(B * ((width + 31) / 32), 1, 148 * 8);
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
