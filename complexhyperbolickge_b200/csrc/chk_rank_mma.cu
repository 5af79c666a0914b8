// K2 (tensor-core tier) — filtered rank counts with the Hermitian contraction on tcgen05 (sm_100a only).
//
// Same contract as the exact tier (chk_rank.cu; reference models/base.py:243-271 + Distance.forward,
// utils/complexhyperbolic.py:212-237) and the SAME integer counts, by filter-and-refine:
//
//   1. The complex contraction  <z_i, w_e> = sum_k z_ik conj(w_ek)  is a real GEMM
//          D[e, 2i]   = sum_k w_e[k] * qre_i[k]      qre_i = [ Re z_i |  Im z_i ]
//          D[e, 2i+1] = sum_k w_e[k] * qim_i[k]      qim_i = [ Im z_i | -Re z_i ]      (K = 2r)
//      run as tcgen05.mma kind::f16 on bf16 operands with fp32 accumulation in TMEM.  fp32 inputs are split
//      a = hi + lo (hi = bf16(a), lo = bf16(a - hi)) and three products hi*hi + hi*lo + lo*hi are
//      accumulated into the same TMEM tile ("bf16x3": per-product error <= 3.004 * 2^-18 |a||b|).
//      UMMA M = 128 entity rows (TMEM lanes), N = 256 query rows (= 128 queries, TMEM columns).
//   2. Operands arrive through the TMA engine as 1-D bulk copies (cp.async.bulk + mbarrier complete_tx)
//      of PRE-TILED blocks: the entity table has a bf16 hi/lo shadow (built once per evaluation pass,
//      chk_entity_shadow_build) whose blocks are byte images of the shared-memory operand tile in the
//      canonical no-swizzle K-major UMMA layout (8-row x 16-byte core matrices), so one 16 KB copy per
//      stage feeds the entity side and one 32 KB copy the query side.  4-stage full/empty mbarrier ring,
//      one elected producer thread, one elected MMA-issuer thread, accumulators double-buffered in TMEM
//      (2 x 256 columns) so the epilogue of tile t overlaps the MMAs of tile t+1.
//   3. Epilogue (8 warps, tcgen05.ld 32x32b): per pair an APPROXIMATE score s~ (MUFU rsqrt / lg2) and a
//      proven bound Delta >= |s~ - s_exact| on its distance to the exact tier's fp32 score (split error,
//      accumulation error of both tiers by Cauchy-Schwarz with the row norms, evaluation roundoff, and
//      |ds/dx| = 2 acosh(x)/sqrt(x^2-1)).  Pairs with s~ - Delta >= target are counted, pairs with
//      s~ + Delta < target are dropped, the (rare) rest goes to a list that recheck_kernel re-scores with
//      the canonical exact chain (exact_pair) — so counts are identical to CHK_RANK_FMA.  Pairs that are
//      provably in the clamp regime (x <= 1+eps, every pair at init_size=1e-3) are decided exactly in the
//      epilogue: their exact score is (bh+bt) + s(1+eps), bit-for-bit what the exact tier computes.
//
// Persistent kernel: one CTA per SM, static round-robin over entity tiles, all query tiles per entity
// tile back-to-back (the entity blocks are re-read from L2, from HBM only once per call).
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include "chk_common.cuh"

namespace {

constexpr int TILE_E = 128;                 // entity rows per tile  (UMMA M)
constexpr int TILE_QR = 256;                // query rows per tile   (UMMA N) = 2 rows per query
constexpr int TILE_Q = TILE_QR / 2;         // queries per tile
constexpr int KC = 32;                      // K elements per pipeline stage (2 UMMA K-steps of 16)
constexpr int A_PART = TILE_E * KC * 2;     // bytes of the hi (or lo) half of an entity block
constexpr int A_BLOCK = 2 * A_PART;         // 16 KB
constexpr int B_PART = TILE_QR * KC * 2;
constexpr int B_BLOCK = 2 * B_PART;         // 32 KB
constexpr int MAX_B = 1024;                 // queries per launch (8 query tiles)
constexpr int THREADS = 384;                // warp 0 producer, 1 MMA issuer, 2 TMEM alloc, 3 spare, 4..11 epilogue
constexpr int EPI_WARPS = 8;
constexpr int TMEM_COLS = 512;
constexpr int HDR_BYTES = 256;
constexpr unsigned SPIN_LIMIT = 1u << 28;   // bounded mbarrier spin: a protocol bug traps instead of hanging the GPU

constexpr int SMEM_QC = MAX_B * 16;
constexpr int SMEM_QNY = MAX_B * 8;
constexpr int SMEM_CNT = MAX_B * 4;
constexpr int SMEM_BAR = 256;

// The MMA contracts over complex coefficients 0..r-2 only: K = 2(r-1) = n is a power of two (a multiple of KC for
// every supported rank), operand rows are [Re_0..Re_{r-2} | Im_0..Im_{r-2}].  The last coefficient (k = r-1, the
// Nyquist bin) is added in the epilogue with four fp32 FMAs, so no MMA work is spent on zero padding.
__host__ __device__ inline int kpad_of(int rank) { return (2 * (rank - 1) + KC - 1) / KC * KC; }
// column k' of an operand row -> column of the [Re | Im] source row (or -1: zero padding)
__host__ __device__ inline int src_col(int rank, int kp) {
    const int h = rank - 1;
    return kp < h ? kp : (kp < 2 * h ? rank + (kp - h) : -1);
}

// byte offset of element (row, k) inside one part of a block with R rows (no-swizzle K-major canonical layout):
// core matrix = 8 rows x 8 bf16 (128 contiguous bytes); K-adjacent core matrices R*16 bytes apart (LBO),
// row-group-adjacent ones 128 bytes apart (SBO).
__host__ __device__ inline int tile_off(int R, int row, int k) {
    return (k >> 3) * (R * 16) + (row >> 3) * 128 + (row & 7) * 16 + (k & 7) * 2;
}

// ---------------------------------------------------------------------------------------------- PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > SPIN_LIMIT) { printf("chk_rank_mma: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// Single-thread instructions issued from CONVERGED warps: every lane runs the surrounding (warp-uniform) control flow
// and the instruction itself is predicated on elect.sync inside the asm block.  Keeping the roles' loops uniform lets
// ptxas hold descriptors / addresses in uniform registers instead of serialising a divergent `if (lane == 0)` region.
__device__ __forceinline__ void bulk_g2s_elect(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred pe;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "@pe cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_elect(uint32_t bar, uint32_t bytes) {
    asm volatile(
        "{\n\t.reg .pred pe;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "@pe mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}"
        ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float rsqrt_approx(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// shared-memory matrix descriptor, no swizzle, K-major (cute::UMMA::SmemDescriptor bit layout):
// [0,14) addr>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=0 (SWIZZLE_NONE)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor kind::f16: D=f32 (bit 4), A=B=bf16 (bits 7,10), K-major both, N>>3 at 17, M>>4 at 24
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------- operand builders
__device__ __forceinline__ void split_bf16(float a, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(a);
    lo = __float2bfloat16_rn(a - __bfloat162float(hi));
}
__device__ __forceinline__ void split_bf16(double a, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __double2bfloat16(a);                                   // round-to-nearest
    lo = __double2bfloat16(a - (double)__bfloat162float(hi));    // residual formed in fp64: |a - hi - lo| <= 2^-18 |a|
}

// entity [n_rows, 2r] (fp32 or fp64) -> blocks[(et*nk + kc)] = { hi part | lo part },
// aux[e] = { ||w_e|| (rounded up), Re w_{r-1}, Im w_{r-1}, 1/hn_e } and bt32[e] = (float) bt_e  (all fp32: epilogue inputs)
template <typename TIn>
__global__ void __launch_bounds__(256) entity_shadow_kernel(const TIn* __restrict__ entity, const TIn* __restrict__ hn,
                                                            const TIn* __restrict__ bt, int64_t n_rows, int r, int nk,
                                                            uint8_t* __restrict__ blocks, float4* __restrict__ aux,
                                                            float* __restrict__ bt32) {
    __shared__ TIn sT[TILE_E][KC + 1];
    __shared__ float sSq[TILE_E][4];
    const int64_t et = blockIdx.x;
    const int tid = threadIdx.x;
    const int K2 = 2 * r;
    float sq[2] = {0.f, 0.f};
    for (int kc = 0; kc < nk; ++kc) {
        __syncthreads();
        for (int idx = tid; idx < TILE_E * KC; idx += 256) {
            int row = idx / KC, kk = idx - row * KC;
            int64_t e = et * TILE_E + row;
            int k = src_col(r, kc * KC + kk);
            sT[row][kk] = (e < n_rows && k >= 0) ? entity[e * K2 + k] : TIn(0);
        }
        __syncthreads();
        uint8_t* blk = blocks + ((size_t)et * nk + kc) * A_BLOCK;
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            int item = tid + it * 256;                 // (row, kcore): consecutive threads -> consecutive rows
            int row = item & (TILE_E - 1), kcore = item >> 7;
            __align__(16) __nv_bfloat16 h[8], l[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const TIn a = sT[row][kcore * 8 + j];
                split_bf16(a, h[j], l[j]);
                sq[it] = fmaf((float)a, (float)a, sq[it]);
            }
            int off = tile_off(TILE_E, row, kcore * 8);
            *reinterpret_cast<uint4*>(blk + off) = *reinterpret_cast<const uint4*>(h);
            *reinterpret_cast<uint4*>(blk + A_PART + off) = *reinterpret_cast<const uint4*>(l);
        }
    }
#pragma unroll
    for (int it = 0; it < 2; ++it) { int item = tid + it * 256; sSq[item & (TILE_E - 1)][item >> 7] = sq[it]; }
    __syncthreads();
    if (tid < TILE_E) {
        float s = (sSq[tid][0] + sSq[tid][1]) + (sSq[tid][2] + sSq[tid][3]);
        const int64_t e = et * TILE_E + tid;
        const float wrn = e < n_rows ? (float)entity[e * K2 + r - 1] : 0.f, win = e < n_rows ? (float)entity[e * K2 + 2 * r - 1] : 0.f;
        s = fmaf(wrn, wrn, fmaf(win, win, s));
        aux[e] = make_float4(sqrtf(s) * (1.0f + 1e-5f), wrn, win, e < n_rows ? (float)(TIn(1) / hn[e]) : -1.0f);
        bt32[e] = e < n_rows ? (bt ? (float)bt[e] : 0.f) : -1e30f;      // padding rows can never reach a target
    }
}

// q fp32 [b, 2r] -> query blocks[(qt*nk + kc)] (rows 2i: [Re|Im], rows 2i+1: [Im|-Re]); grid (nk, n_qt), 256 threads
// pair_layout: the block is two 128-row halves {hi | lo}{hi | lo} (one per CTA of a pair) instead of {hi | lo} of 256 rows
template <typename TIn>
__global__ void __launch_bounds__(256) query_blocks_kernel(const TIn* __restrict__ q, int b, int r, int nk, int pair_layout,
                                                           uint8_t* __restrict__ blocks) {
    const int kc = blockIdx.x, qt = blockIdx.y;
    const int K2 = 2 * r;
    uint8_t* blk = blocks + ((size_t)qt * nk + kc) * B_BLOCK;
    for (int item = threadIdx.x; item < TILE_QR * (KC / 8); item += 256) {
        int row = item & (TILE_QR - 1), kcore = item >> 8;
        int i = qt * TILE_Q + (row >> 1);
        bool im_row = row & 1;
        __align__(16) __nv_bfloat16 h[8], l[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int kp = kc * KC + kcore * 8 + j, hh = r - 1;
            TIn a = TIn(0);
            if (i < b && kp < 2 * hh) {
                const TIn* z = q + (size_t)i * K2;        // re-row: [Re z | Im z], im-row: [Im z | -Re z]
                a = !im_row ? (kp < hh ? z[kp] : z[r + kp - hh]) : (kp < hh ? z[r + kp] : -z[kp - hh]);
            }
            split_bf16(a, h[j], l[j]);
        }
        int off, lo_off;
        if (pair_layout) { off = (row >> 7) * (B_BLOCK / 2) + tile_off(TILE_QR / 2, row & 127, kcore * 8); lo_off = B_PART / 2; }
        else { off = tile_off(TILE_QR, row, kcore * 8); lo_off = B_PART; }
        *reinterpret_cast<uint4*>(blk + off) = *reinterpret_cast<const uint4*>(h);
        *reinterpret_cast<uint4*>(blk + lo_off + off) = *reinterpret_cast<const uint4*>(l);
    }
}

// per-query constants {2/zn, bh, target, ||z||} and the batch maxima of ||z|| and |bh| (hdr[1], hdr[2]).  One warp per query.
template <typename TIn>
__global__ void __launch_bounds__(256) query_consts_kernel(const TIn* __restrict__ q, const TIn* __restrict__ qn,
                                                           const TIn* __restrict__ bh_vals, const TIn* __restrict__ target,
                                                           int b, int r, float4* __restrict__ qc, float2* __restrict__ qny,
                                                           unsigned* __restrict__ hdr) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= b) return;
    const TIn* z = q + (size_t)i * 2 * r;
    float s = 0.f;
    for (int k = lane; k < 2 * r; k += 32) s = fmaf((float)z[k], (float)z[k], s);
    s = warp_sum<float>(s);
    if (lane == 0) {
        const float nz = sqrtf(s) * (1.0f + 1e-5f), bh = bh_vals ? (float)bh_vals[i] : 0.f;
        qc[i] = make_float4((float)(TIn(2) / qn[i]), bh, (float)target[i], nz);
        qny[i] = make_float2((float)z[r - 1], (float)z[2 * r - 1]);        // Nyquist coefficient of the query
        atomicMax(hdr + 1, __float_as_uint(nz));            // non-negative floats order like their bit patterns
        atomicMax(hdr + 2, __float_as_uint(fabsf(bh)));
    }
}

// ---------------------------------------------------------------------------------------------- main kernel
struct MmaArgs {
    const uint8_t* a_blocks; const float4* aux;        // entity shadow: operand blocks, {||w||, Re w_{r-1}, Im w_{r-1}}
    const uint8_t* b_blocks; const float4* qc; const float2* qny;   // per-call query operands
    const float* bt32;                                 // entity shadow: fp32 copy of bt (0 without bias, -1e30 on padding rows)
    float xclamp;                                      // 1 + eps of the MODEL dtype (4e-3 fp32, 1e-5 fp64), as fp32
    int exact_clamp;                                   // fp32 models: clamp-regime pairs are decided bit-exactly in the epilogue
    int64_t n_rows; int b, nk, n_et, n_qt;
    float eps_dot;                                     // |re~ - re_exact| <= eps_dot * ||z|| ||w|| (both tiers' errors)
    unsigned* hdr; uint2* list; unsigned list_cap;     // hdr[0] = list length, hdr[1..2] = max||z||, max|bh|, hdr[4] = overflow (sticky)
    unsigned long long* counts;
    int dump_raw;                                      // DEBUG bring-up: store (re, im) instead of (score, band)
    float* dbg_scores; float* dbg_band;                // DEBUG: [b, n_rows] approximate score and band
};

// Geometry of the two variants.
//   PAIR = false: one CTA per tile, tcgen05.mma.cta_group::1, M = 128 entities x N = 256 query rows; per stage the
//                 CTA stages its entity block and the WHOLE query block (48 KB at KC = 32).
//   PAIR = true : a 2-CTA cluster per tile pair, tcgen05.mma.cta_group::2 issued by the leader CTA, M = 256 entities
//                 (128 TMEM lanes in each CTA) x N = 256 query rows; each CTA stages its own entity block and HALF
//                 of the query block (32 KB per stage) — the tensor core reads the other half from the peer's shared
//                 memory, which halves the query-side shared-memory fill and read traffic per CTA.
template <bool PAIR> struct Geo {
    static constexpr int B_ROWS = PAIR ? TILE_QR / 2 : TILE_QR;      // query rows staged per CTA
    static constexpr int B_BYTES = 2 * B_ROWS * KC * 2;              // hi + lo
    static constexpr int STAGE = A_BLOCK + B_BYTES;
    static constexpr int NSTAGE = PAIR ? 6 : 4;
    static constexpr int NBAR = 3 * NSTAGE + 4;                      // full, empty, peer_full, tfull[2], tempty[2]
    static constexpr int SMEM = NSTAGE * STAGE + SMEM_QC + SMEM_QNY + SMEM_CNT + SMEM_BAR;
    static constexpr int UMMA_M = PAIR ? 2 * TILE_E : TILE_E;
};

template <bool PAIR> __device__ __forceinline__ void tc_mma_issue(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (PAIR) {
        asm volatile(
            "{\n\t.reg .pred p, pe;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "elect.sync _|pe, 0xffffffff;\n\t"
            "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p, pe;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "elect.sync _|pe, 0xffffffff;\n\t"
            "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    }
}
// MMA-completion arrive; PAIR: on the barrier at this offset in BOTH CTAs of the pair
template <bool PAIR> __device__ __forceinline__ void tc_commit_to(uint32_t bar) {
    if constexpr (PAIR) {
        asm volatile(
            "{\n\t.reg .pred pe;\n\t"
            "elect.sync _|pe, 0xffffffff;\n\t"
            "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
            ::"r"(bar), "h"((uint16_t)3) : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred pe;\n\t"
            "elect.sync _|pe, 0xffffffff;\n\t"
            "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
            ::"r"(bar) : "memory");
    }
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nid_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at local address `bar` of CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster_elect(uint32_t bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t.reg .pred pe;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "@pe mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(bar), "r"(rank) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(bar), "r"(rank) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    unsigned spins = 0;
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (++spins > SPIN_LIMIT) { printf("chk_rank_mma: cluster mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
    }
}

template <bool DEBUG, bool PAIR>
__global__ void __launch_bounds__(THREADS, 1) rank_mma_kernel(const MmaArgs A) {
    using G = Geo<PAIR>;
    constexpr int NS = G::NSTAGE;
    extern __shared__ __align__(1024) uint8_t smem[];
    float4* sQc = reinterpret_cast<float4*>(smem + NS * G::STAGE);
    float2* sQny = reinterpret_cast<float2*>(smem + NS * G::STAGE + SMEM_QC);
    int* sCnt = reinterpret_cast<int*>(smem + NS * G::STAGE + SMEM_QC + SMEM_QNY);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NS * G::STAGE + SMEM_QC + SMEM_QNY + SMEM_CNT);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + G::NBAR);
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + NS), bar_pfull = smem_u32(bars + 2 * NS);
    const uint32_t bar_tfull = smem_u32(bars + 3 * NS), bar_tempty = smem_u32(bars + 3 * NS + 2);
    const uint32_t stage0 = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = PAIR ? cluster_ctarank() : 0u;              // 0 = leader (issues the MMAs)
    const int unit = PAIR ? (int)cluster_id_x() : (int)blockIdx.x;     // persistent work unit (CTA or CTA pair)
    const int n_units = PAIR ? (int)cluster_nid_x() : (int)gridDim.x;
    const int n_super = PAIR ? (A.n_et + 1) / 2 : A.n_et;              // entity tiles (or tile pairs) in the shard

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); mbar_init(bar_pfull + 8 * s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, PAIR ? 2 * EPI_WARPS : EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        if constexpr (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    for (int i = threadIdx.x; i < MAX_B; i += THREADS) {
        // slots past the batch: a huge FINITE target (the band has a 2^-22 |target| term; inf would make it inf and list the pair)
        sQc[i] = i < A.b ? A.qc[i] : make_float4(-2.f, 0.f, 3.0e38f, 0.f);
        sQny[i] = i < A.b ? A.qny[i] : make_float2(0.f, 0.f);
        sCnt[i] = 0;
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();            // both CTAs' barriers and TMEM exist before any cross-CTA signal
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_items = ((n_super - unit + n_units - 1) / n_units) * A.n_qt;

    if (warp == 0) {
        // ===== producer: the warp runs the loop converged, one elected lane issues the bulk copies =====
        {
            int stage = 0; uint32_t phase = 0;
            for (int it = 0; it < n_items; ++it) {
                const int sup = unit + (it / A.n_qt) * n_units, qt = it % A.n_qt;
                const int et = PAIR ? 2 * sup + (int)crank : sup;      // the shadow is padded to an even tile count
                const uint8_t* asrc = A.a_blocks + (size_t)et * A.nk * A_BLOCK;
                const uint8_t* bsrc = A.b_blocks + (size_t)qt * A.nk * B_BLOCK + (PAIR ? crank * G::B_BYTES : 0);
                for (int kc = 0; kc < A.nk; ++kc) {
                    mbar_wait_cluster(bar_empty + 8 * stage, phase ^ 1);
                    mbar_arrive_expect_tx_elect(bar_full + 8 * stage, G::STAGE);
                    const uint32_t dst = stage0 + stage * G::STAGE;
                    bulk_g2s_elect(dst, asrc + (size_t)kc * A_BLOCK, A_BLOCK, bar_full + 8 * stage);
                    bulk_g2s_elect(dst + A_BLOCK, bsrc + (size_t)kc * B_BLOCK, G::B_BYTES, bar_full + 8 * stage);
                    if (++stage == NS) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        {
            if (crank == 0) {
                // ===== MMA issuer: warp 1 of the leader CTA, converged; one elected lane issues =====
                constexpr uint32_t idesc = umma_idesc(G::UMMA_M, TILE_QR);
                constexpr uint32_t a_lbo = TILE_E * 16, b_lbo = G::B_ROWS * 16, sbo = 128;   // K-step / row-group strides
                constexpr uint32_t b_part = G::B_ROWS * KC * 2;
                int stage = 0; uint32_t phase = 0;
                for (int it = 0; it < n_items; ++it) {
                    const int acc = it & 1; const uint32_t acc_phase = (it >> 1) & 1;
                    mbar_wait_cluster(bar_tempty + 8 * acc, acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * TILE_QR;
                    for (int kc = 0; kc < A.nk; ++kc) {
                        mbar_wait(bar_full + 8 * stage, phase);
                        if constexpr (PAIR) mbar_wait_cluster(bar_pfull + 8 * stage, phase);   // the peer's operands landed too
                        tc_fence_after();
                        const uint32_t sa = stage0 + stage * G::STAGE, sb = sa + A_BLOCK;
#pragma unroll
                        for (int j = 0; j < KC / 16; ++j) {
                            const uint64_t a_hi = umma_desc(sa + j * (TILE_E * 32), a_lbo, sbo);
                            const uint64_t a_lo = umma_desc(sa + A_PART + j * (TILE_E * 32), a_lbo, sbo);
                            const uint64_t b_hi = umma_desc(sb + j * (G::B_ROWS * 32), b_lbo, sbo);
                            const uint64_t b_lo = umma_desc(sb + b_part + j * (G::B_ROWS * 32), b_lbo, sbo);
                            tc_mma_issue<PAIR>(d_tmem, a_hi, b_hi, idesc, (kc | j) != 0);
                            tc_mma_issue<PAIR>(d_tmem, a_hi, b_lo, idesc, 1);
                            tc_mma_issue<PAIR>(d_tmem, a_lo, b_hi, idesc, 1);
                        }
                        tc_commit_to<PAIR>(bar_empty + 8 * stage);        // frees the smem stage (in both CTAs) when these MMAs retire
                        if (++stage == NS) { stage = 0; phase ^= 1; }
                    }
                    tc_commit_to<PAIR>(bar_tfull + 8 * acc);              // accumulator tile complete (both CTAs' epilogues)
                }
            } else {
                // ===== peer CTA: relay "my operands have landed" to the leader's peer_full barriers =====
                int stage = 0; uint32_t phase = 0;
                for (int it = 0; it < n_items; ++it)
                    for (int kc = 0; kc < A.nk; ++kc) {
                        mbar_wait(bar_full + 8 * stage, phase);
                        mbar_arrive_cluster_elect(bar_pfull + 8 * stage, 0);
                        if (++stage == NS) { stage = 0; phase ^= 1; }
                    }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> approximate score + band -> counts / re-check list =====
        // Per pair (re, im from TMEM; query constants broadcast from smem; entity constants in registers):
        //   mod2 = (re-1)^2 + im^2,  x = g mod2 - 1  (g = 2/(zn wn)),  dx >= |x~ - x_exact|,
        //   l = lg2(xc + sqrt(xc^2-1)) (xc = max(x, 1+eps)),  s~ = bias - ln2^2 l^2,
        //   band = (2 rho + dx) dx + slop,  rho = acosh/sqrt(x^2-1) = ln2 l rsqrt(xc^2-1)   (|d rho/dx| <= 1/3).
        const int lq = warp & 3;                       // TMEM lane quarter this warp may read
        const int chalf = (warp - 4) >> 2;             // which 128 columns (64 queries) of the tile
        const float xclamp = A.xclamp;
        const float s_clamp = score_from_x<float>(1.0f + Sc<float>::ball_eps, false, 0.f, 0.f);   // exact fp32 tier's -acosh(1+eps)^2
        const float eps = A.eps_dot;
        const float nz_max = __uint_as_float(A.hdr[1]), bh_max = __uint_as_float(A.hdr[2]);
        constexpr float kx = 24.f * 5.9604645e-8f;     // relative roundoff of (x+1) in both tiers (incl. fp32 copies of fp64 inputs)
        constexpr float LN2 = 0.69314718f, LN2SQ = 0.48045301f;
        constexpr float k1 = (3.8146973e-6f + 4.7683716e-7f) * LN2SQ;          // (2^-18 + 2^-21) d^2: evaluation roundoff of d^2
        const uint32_t tbase = tmem_base + ((uint32_t)(lq * 32) << 16) + chalf * 128;
        for (int it = 0; it < n_items; ++it) {
            const int sup = unit + (it / A.n_qt) * n_units, qt = it % A.n_qt;
            const int et = PAIR ? 2 * sup + (int)crank : sup;
            const int acc = it & 1; const uint32_t acc_phase = (it >> 1) & 1;
            const int64_t e = (int64_t)et * TILE_E + lq * 32 + lane;
            const bool e_ok = e < A.n_rows;
            // the shadow is padded to whole tiles: rows past the end of the shard carry bt = -1e30 (their score can never
            // reach a target and stays finite), zero operands and 1/hn = -1
            const float4 ax = A.aux[e];
            const float bte = A.bt32[e];
            const float nwe = ax.x, wrn = ax.y, win = ax.z, iwn = ax.w;   // ||w||, Nyquist coefficient, 1/hn
            (void)e_ok;
            // dm2 = [sqrt2 eps P (1 + 4.3 eps P) + kx] (1 + mod2) >= |mod2~ - mod2| + kx mod2, P = ||z|| ||w|| <= nz_max ||w||
            const float cae = 1.4142136f * eps * nwe * (1.0f + 4.3f * eps * nz_max * nwe);
            // additive part of the band: 2^-21 (lg2/rsqrt approximation at small d) + 2^-22 (xc^2-1 cancellation)
            // + 2^-22 (|bh| + |bt|) (rounding of the bias adds)
            const float slop0 = 4.7683716e-7f + 2.3841858e-7f + 2.3841858e-7f * (fabsf(bte) + bh_max);
            mbar_wait_cluster(bar_tfull + 8 * acc, acc_phase);
            tc_fence_after();
            const uint32_t tacc = tbase + acc * TILE_QR;

            auto process = [&](const uint32_t (&v)[32], int g) {
                const int qbase = qt * TILE_Q + chalf * 64 + g * 16;
                unsigned m_sure = 0, m_hi = 0;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float4 c = sQc[qbase + j];              // {2/zn, bh, target (3e38 beyond b), ||z||}: broadcast
                    const float2 zn_ = sQny[qbase + j];           // coefficient r-1 is contracted here, not in the MMA
                    const float re = fmaf(zn_.x, wrn, fmaf(zn_.y, win, __uint_as_float(v[2 * j])));
                    const float im = fmaf(zn_.y, wrn, fmaf(-zn_.x, win, __uint_as_float(v[2 * j + 1])));
                    const float r1 = re - 1.0f;
                    const float mod2 = fmaf(r1, r1, im * im);
                    const float pe = fmaf(c.w, cae, kx);
                    const float dm2 = fmaf(pe, mod2, pe);
                    const float gq = c.x * iwn;                   // 2/(zn wn) > 0
                    const float x = fmaf(gq, mod2, -1.0f);
                    const float dx = gq * dm2;                    // bound on |x~ - x_exact| (before the clamp)
                    const bool clamped = A.exact_clamp && (x + dx) <= xclamp;     // the exact tier's x is clamped for sure
                    const float xc = fmaxf(x, xclamp);
                    const float t = fmaf(xc, xc, -1.0f);
                    const float rs = rsqrt_approx(t);
                    const float l = lg2_approx(fmaf(t, rs, xc));
                    const float a2 = fmaf(l * rs, 2.0f * LN2, dx);       // 2 rho + dx
                    const float l2 = l * l;
                    const float bias = __fadd_rn(c.y, bte);
                    float band = fmaf(a2, dx, fmaf(k1, l2, fmaf(2.3841858e-7f, fabsf(c.z), slop0)));   // + 2^-22 |target|
                    float s = fmaf(-LN2SQ, l2, bias);
                    if (clamped) { s = __fadd_rn(bias, s_clamp); band = 0.f; }
                    m_sure |= (s - band >= c.z) ? (1u << j) : 0u;             // certainly >= target
                    m_hi |= !(s + band < c.z) ? (1u << j) : 0u;               // possibly >= target (NaN -> re-check)
                    if (DEBUG) {
                        if (e < A.n_rows && qbase + j < A.b) {
                            A.dbg_scores[(size_t)(qbase + j) * A.n_rows + e] = A.dump_raw ? re : s;
                            A.dbg_band[(size_t)(qbase + j) * A.n_rows + e] = A.dump_raw ? im : band;
                        }
                    }
                }
                // per-query counts of this warp's 32 entities: spread 4 mask bits into 4 bytes, sum over the lanes (<= 32 per byte)
                unsigned ambmask = m_hi & ~m_sure;
                const unsigned r0 = __reduce_add_sync(CHK_FULL, ((m_sure & 0xFu) * 0x00204081u) & 0x01010101u);
                const unsigned r1_ = __reduce_add_sync(CHK_FULL, (((m_sure >> 4) & 0xFu) * 0x00204081u) & 0x01010101u);
                const unsigned r2 = __reduce_add_sync(CHK_FULL, (((m_sure >> 8) & 0xFu) * 0x00204081u) & 0x01010101u);
                const unsigned r3 = __reduce_add_sync(CHK_FULL, (((m_sure >> 12) & 0xFu) * 0x00204081u) & 0x01010101u);
                if (lane < 16) {
                    const unsigned rr = lane < 8 ? (lane < 4 ? r0 : r1_) : (lane < 12 ? r2 : r3);
                    const unsigned cnt = (rr >> (8 * (lane & 3))) & 0xffu;
                    if (cnt) atomicAdd(&sCnt[qbase + lane], (int)cnt);
                }
                if (__any_sync(CHK_FULL, ambmask != 0)) {
                    while (ambmask) {
                        const int j = __ffs(ambmask) - 1;
                        ambmask &= ambmask - 1;
                        const unsigned slot = atomicAdd(A.hdr, 1u);
                        if (slot < A.list_cap) A.list[slot] = make_uint2((unsigned)(qbase + j), (unsigned)e);
                        else atomicExch(A.hdr + 4, 1u);
                    }
                }
            };

            uint32_t va[32], vb[32];                   // TMEM loads software-pipelined one 16-query group ahead
            tc_ld32(tacc + 0, va);  tc_wait_ld();
            tc_ld32(tacc + 32, vb); process(va, 0); tc_wait_ld();
            tc_ld32(tacc + 64, va); process(vb, 1); tc_wait_ld();
            tc_ld32(tacc + 96, vb); process(va, 2); tc_wait_ld();
            tc_fence_before();                         // all TMEM reads of this tile are done: hand the buffer back
            __syncwarp();
            if (lane == 0) { if constexpr (PAIR) mbar_arrive_cluster(bar_tempty + 8 * acc, 0); else mbar_arrive(bar_tempty + 8 * acc); }
            process(vb, 3);
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();            // no CTA leaves while its peer may still signal it or read its smem
    for (int i = threadIdx.x; i < A.b; i += THREADS)
        if (sCnt[i]) atomicAdd(A.counts + i, (unsigned long long)sCnt[i]);
    if (warp == 2) {
        tc_fence_after();
        if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// exact re-check of the pairs the epilogue could not decide (canonical chain -> same bits as the exact tier);
// one lane per pair, rows staged with coalesced warp loads (warp_exact_pairs)
template <typename T>
__global__ void __launch_bounds__(32) recheck_kernel(RArgs<T> A, const unsigned* __restrict__ hdr,
                                                                                const uint2* __restrict__ list, unsigned cap) {
    constexpr int WARPS = 1;
    __shared__ PairTiles<T> S[WARPS];
    const unsigned n = hdr[0] < cap ? hdr[0] : cap;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (unsigned t0 = (blockIdx.x * WARPS + warp) * 32; t0 < n; t0 += gridDim.x * WARPS * 32) {
        const unsigned t = t0 + lane;
        const bool valid = t < n;
        const uint2 p = valid ? list[t] : make_uint2(0u, 0u);
        const T s = warp_exact_pairs<T>(A, p.x, p.y, valid, S[warp]);
        if (valid && s >= A.target[p.x]) atomicAdd(A.counts + p.x, 1ull);
    }
}

struct Workspace {
    unsigned* hdr; float4* qc; float2* qny; uint8_t* b_blocks; uint2* list; unsigned list_cap;
};

bool carve_workspace(int rank, void* ws, int64_t bytes, Workspace& W) {
    const int nk = kpad_of(rank) / KC;
    const int64_t fixed = HDR_BYTES + SMEM_QC + SMEM_QNY + (int64_t)(MAX_B / TILE_Q) * nk * B_BLOCK;
    if (bytes < fixed + 8 * 1024) return false;
    uint8_t* p = (uint8_t*)ws;
    W.hdr = (unsigned*)p;
    W.qc = (float4*)(p + HDR_BYTES);
    W.qny = (float2*)(p + HDR_BYTES + SMEM_QC);
    W.b_blocks = p + HDR_BYTES + SMEM_QC + SMEM_QNY;
    W.list = (uint2*)(p + fixed);
    int64_t cap = (bytes - fixed) / 8;
    W.list_cap = (unsigned)(cap > 0x7fffffff ? 0x7fffffff : cap);
    return true;
}

int g_num_sms = 0;
thread_local cudaEvent_t g_prof_start = nullptr, g_prof_stop = nullptr;   // measurement support (chk_rank_mma_profile_events)

}  // namespace

extern "C" int chk_rank_mma_profile_events(void* ev_start, void* ev_stop) {
    g_prof_start = (cudaEvent_t)ev_start; g_prof_stop = (cudaEvent_t)ev_stop;
    return CHK_OK;
}

extern "C" int64_t chk_entity_shadow_bytes(int rank, int64_t n_rows) {
    if (rank < 2 || n_rows <= 0) return 0;
    const int64_t n_et = ((n_rows + TILE_E - 1) / TILE_E + 1) / 2 * 2;       // padded to whole tile pairs
    return n_et * (kpad_of(rank) / KC) * (int64_t)A_BLOCK + n_et * TILE_E * (16 + 4);      // blocks + aux (float4) + bt32
}

extern "C" int chk_entity_shadow_build(int dtype, int rank, int64_t n_rows, const void* entity, const void* hn, const void* bt,
                                       void* shadow, void* stream) {
    if (n_rows == 0) return CHK_OK;
    if (rank < 2 || n_rows < 0 || !entity || !hn || !shadow) { chk_set_error("chk_entity_shadow_build: bad argument"); return CHK_EINVAL; }
    const int64_t n_et = ((n_rows + TILE_E - 1) / TILE_E + 1) / 2 * 2;       // padded to whole tile pairs (zero rows)
    if (n_et > 0x7fffffff) { chk_set_error("chk_entity_shadow_build: shard too large"); return CHK_EUNSUPPORTED; }
    const int nk = kpad_of(rank) / KC;
    uint8_t* blocks = (uint8_t*)shadow;
    float4* aux = (float4*)(blocks + n_et * nk * (int64_t)A_BLOCK);
    float* bt32 = (float*)(aux + n_et * TILE_E);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CHK_F32) entity_shadow_kernel<float><<<(unsigned)n_et, 256, 0, st>>>((const float*)entity, (const float*)hn, (const float*)bt, n_rows, rank, nk, blocks, aux, bt32);
    else if (dtype == CHK_F64) entity_shadow_kernel<double><<<(unsigned)n_et, 256, 0, st>>>((const double*)entity, (const double*)hn, (const double*)bt, n_rows, rank, nk, blocks, aux, bt32);
    else { chk_set_error("unknown dtype %d", dtype); return CHK_EINVAL; }
    CHK_CUDA_LAUNCH_CHECK("entity_shadow_kernel");
    return CHK_OK;
}

extern "C" int64_t chk_rank_mma_workspace_bytes(int rank, int64_t b) {
    if (rank < 2) return 0;
    const int nk = kpad_of(rank) / KC;
    const int64_t fixed = HDR_BYTES + SMEM_QC + SMEM_QNY + (int64_t)(MAX_B / TILE_Q) * nk * B_BLOCK;
    int64_t cap = b * 8192;                      // re-check list entries
    if (cap < (1 << 20)) cap = 1 << 20;
    if (cap > (8 << 20)) cap = 8 << 20;
    return fixed + cap * 8;
}

extern "C" int chk_rank_mma_reset(void* workspace, void* stream) {
    if (!workspace) { chk_set_error("chk_rank_mma_reset: null workspace"); return CHK_EINVAL; }
    if (cudaMemsetAsync(workspace, 0, HDR_BYTES, (cudaStream_t)stream) != cudaSuccess) { chk_set_error("cudaMemsetAsync failed"); return CHK_ECUDA; }
    return CHK_OK;
}

extern "C" int chk_rank_mma_status(const void* workspace, int64_t* last_list_len, int* overflowed, void* stream) {
    if (!workspace) { chk_set_error("chk_rank_mma_status: null workspace"); return CHK_EINVAL; }
    unsigned h[5] = {0, 0, 0, 0, 0};
    cudaError_t e = cudaMemcpyAsync(h, workspace, sizeof(h), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e != cudaSuccess) { chk_set_error("chk_rank_mma_status: %s", cudaGetErrorString(e)); return CHK_ECUDA; }
    if (last_list_len) *last_list_len = h[0];
    if (overflowed) *overflowed = (int)h[4];
    return CHK_OK;
}

template <typename T>
static int rank_mma_launch_t(int rank, int64_t b, const void* q, const void* qn, const void* bh_vals, const void* target,
                           const void* entity, const void* hn, const void* bt, int64_t n_rows, const void* shadow,
                           void* workspace, int64_t workspace_bytes, int64_t* counts, float* dbg_scores, float* dbg_band,
                           cudaStream_t st) {
    Workspace W;
    if (!carve_workspace(rank, workspace, workspace_bytes, W)) { chk_set_error("CHK_RANK_MMA: workspace too small (%lld bytes)", (long long)workspace_bytes); return CHK_EINVAL; }
    if (n_rows > 0xffffffffLL) { chk_set_error("CHK_RANK_MMA: shard too large"); return CHK_EUNSUPPORTED; }
    if (g_num_sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || g_num_sms <= 0) {
            g_num_sms = 0; chk_set_error("CHK_RANK_MMA: no CUDA device"); return CHK_ECUDA;
        }
        if (cudaFuncSetAttribute(rank_mma_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<false>::SMEM) != cudaSuccess ||
            cudaFuncSetAttribute(rank_mma_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<false>::SMEM) != cudaSuccess ||
            cudaFuncSetAttribute(rank_mma_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<true>::SMEM) != cudaSuccess ||
            cudaFuncSetAttribute(rank_mma_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Geo<true>::SMEM) != cudaSuccess) {
            g_num_sms = 0; chk_set_error("CHK_RANK_MMA: cannot reserve %d bytes of shared memory: %s", Geo<true>::SMEM, cudaGetErrorString(cudaGetLastError())); return CHK_ECUDA;
        }
    }
    // CHK_MMA_CTA_PAIR=1 selects the 2-CTA pair (cta_group::2) variant.  Default is one CTA per tile: on this part the
    // kernel is tensor-pipe / power bound either way and the single-CTA variant measured ~9 % faster.
    const char* pe = getenv("CHK_MMA_CTA_PAIR");
    const bool pair = pe && pe[0] == '1';
    const int nk = kpad_of(rank) / KC;
    const int64_t n_et = (n_rows + TILE_E - 1) / TILE_E;
    const int64_t n_et_pad = (n_et + 1) / 2 * 2;
    const uint8_t* a_blocks = (const uint8_t*)shadow;
    const float4* aux = (const float4*)(a_blocks + n_et_pad * nk * (int64_t)A_BLOCK);
    // Bound on |re~ - re_exact| (and im) relative to ||z|| ||w|| >= sum_k |z_k||w_k| (Cauchy-Schwarz):
    //   exact tier's canonical BLOCKED chain (chk_common.cuh Chain<float>): 2*16 fused steps per block + r/16 block adds,
    //   each <= 2^-24 relative (standard recursive-summation bound);
    //   bf16 split: z w - (zh wh + zh wl + zl wh) = zl wl + dz w + (z - dz) dw with |dz| <= 2^-18 |z|, |dw| <= 2^-18 |w|,
    //   |zl wl| <= 2^-18 (1 + 2^-8) |z||w|  ->  3.004 * 2^-18, taken as 1.16e-5;
    //   tensor-core accumulation (hardware model, stated in DESIGN.md): every tcgen05.mma K=16 step adds its 16
    //   exact products to the fp32 accumulator with at most 2 units of 2^-23 relative to the largest magnitude
    //   involved (<= sum_k |a_k b_k|); 3 * Kpad/16 steps per accumulator.
    // tests/test_gpu_mma.py checks the observed |s~ - s| against the resulting band (it uses < 5 % of it).
    //   the Nyquist coefficient's four fp32 FMAs in the epilogue: 4 * 2^-24.
    // fp64 models: the exact tier is an fp64 chain (2r steps of 2^-53: nothing), the inputs of the epilogue are fp32
    // copies (covered by kx and the |target| / |bias| terms of the slop).
    const double chain_steps = sizeof(T) == 8 ? 0.0 : 2.0 * Chain<float>::BLK + (rank + Chain<float>::BLK - 1) / Chain<float>::BLK;
    const double eps_dot = (chain_steps + 4.0) * 5.9604644775390625e-8 + 1.16e-5 +
                           2.0 * (3.0 * nk * KC / 16.0) * 1.1920928955078125e-7;
    for (int64_t b0 = 0; b0 < b; b0 += MAX_B) {
        const int bc = (int)((b - b0) < MAX_B ? (b - b0) : MAX_B);
        const int n_qt = (bc + TILE_Q - 1) / TILE_Q;
        const T* qp = (const T*)q + b0 * 2 * rank;
        const T* qnp = (const T*)qn + b0;
        const T* bhp = bh_vals ? (const T*)bh_vals + b0 : nullptr;
        const T* tp = (const T*)target + b0;
        if (cudaMemsetAsync(W.hdr, 0, 16, st) != cudaSuccess) { chk_set_error("cudaMemsetAsync failed"); return CHK_ECUDA; }
        query_consts_kernel<T><<<(bc + 7) / 8, 256, 0, st>>>(qp, qnp, bhp, tp, bc, rank, W.qc, W.qny, W.hdr);
        CHK_CUDA_LAUNCH_CHECK("query_consts_kernel");
        query_blocks_kernel<T><<<dim3(nk, n_qt), 256, 0, st>>>(qp, bc, rank, nk, pair ? 1 : 0, W.b_blocks);
        CHK_CUDA_LAUNCH_CHECK("query_blocks_kernel");
        MmaArgs A{};
        A.a_blocks = a_blocks; A.aux = aux; A.b_blocks = W.b_blocks; A.qc = W.qc; A.qny = W.qny;
        A.bt32 = (const float*)(aux + n_et_pad * TILE_E);
        A.xclamp = (float)(T(1) + Sc<T>::ball_eps); A.exact_clamp = sizeof(T) == 4 ? 1 : 0;
        A.n_rows = n_rows; A.b = bc; A.nk = nk; A.n_et = (int)n_et; A.n_qt = n_qt;
        A.eps_dot = (float)eps_dot; A.hdr = W.hdr; A.list = W.list; A.list_cap = W.list_cap;
        A.counts = (unsigned long long*)counts + b0;
        { const char* dr = getenv("CHK_MMA_DUMP_RAW"); A.dump_raw = (dr && dr[0] == '1') ? 1 : 0; }
        A.dbg_scores = dbg_scores ? dbg_scores + b0 * n_rows : nullptr;
        A.dbg_band = dbg_band ? dbg_band + b0 * n_rows : nullptr;
        if (g_prof_start) cudaEventRecord(g_prof_start, st);
        if (pair) {
            const int64_t n_pairs = n_et_pad / 2, max_pairs = g_num_sms / 2;
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3((unsigned)(2 * (n_pairs < max_pairs ? n_pairs : max_pairs)));
            cfg.blockDim = dim3(THREADS);
            cfg.dynamicSmemBytes = Geo<true>::SMEM;
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            cudaError_t le = dbg_scores ? cudaLaunchKernelEx(&cfg, rank_mma_kernel<true, true>, A)
                                        : cudaLaunchKernelEx(&cfg, rank_mma_kernel<false, true>, A);
            if (le != cudaSuccess) { chk_set_error("rank_mma_kernel (cluster launch): %s", cudaGetErrorString(le)); return CHK_ECUDA; }
        } else {
            const unsigned grid = (unsigned)(n_et < g_num_sms ? n_et : g_num_sms);
            if (dbg_scores) rank_mma_kernel<true, false><<<grid, THREADS, Geo<false>::SMEM, st>>>(A);
            else rank_mma_kernel<false, false><<<grid, THREADS, Geo<false>::SMEM, st>>>(A);
        }
        if (g_prof_stop) cudaEventRecord(g_prof_stop, st);
        CHK_CUDA_LAUNCH_CHECK("rank_mma_kernel");
        RArgs<T> R{};
        R.q = qp; R.qn = qnp; R.bh_vals = bhp; R.target = tp; R.entity = (const T*)entity; R.hn = (const T*)hn;
        R.bt = (const T*)bt; R.b = bc; R.n_rows = n_rows; R.r = rank; R.counts = (unsigned long long*)counts + b0;
        recheck_kernel<T><<<g_num_sms * (sizeof(T) == 4 ? 5 : 6), 32, 0, st>>>(R, W.hdr, W.list, W.list_cap);
        CHK_CUDA_LAUNCH_CHECK("recheck_kernel");
    }
    return CHK_OK;
}

static int rank_mma_launch(int dtype, int rank, int64_t b, const void* q, const void* qn, const void* bh_vals, const void* target,
                           const void* entity, const void* hn, const void* bt, int64_t n_rows, const void* shadow,
                           void* workspace, int64_t workspace_bytes, int64_t* counts, float* dbg_scores, float* dbg_band,
                           cudaStream_t st) {
    if (dtype == CHK_F32) return rank_mma_launch_t<float>(rank, b, q, qn, bh_vals, target, entity, hn, bt, n_rows, shadow, workspace, workspace_bytes, counts, dbg_scores, dbg_band, st);
    if (dtype == CHK_F64) return rank_mma_launch_t<double>(rank, b, q, qn, bh_vals, target, entity, hn, bt, n_rows, shadow, workspace, workspace_bytes, counts, dbg_scores, dbg_band, st);
    chk_set_error("unknown dtype %d", dtype); return CHK_EINVAL;
}

int chk_rank_counts_mma(int dtype, int rank, int64_t b, const void* q, const void* qn, const void* bh_vals,
                        const void* target, const void* entity, const void* hn, const void* bt,
                        int64_t n_rows, const void* shadow, void* workspace, int64_t workspace_bytes,
                        int64_t* counts, cudaStream_t st) {
    return rank_mma_launch(dtype, rank, b, q, qn, bh_vals, target, entity, hn, bt, n_rows, shadow, workspace, workspace_bytes,
                           counts, nullptr, nullptr, st);
}

extern "C" int chk_score_all_mma(int dtype, int rank, int64_t b, const void* q, const void* qn, const void* bh_vals,
                                 const void* target, const void* entity, const void* hn, const void* bt, int64_t n_rows,
                                 const void* shadow, void* workspace, int64_t workspace_bytes, int64_t* counts,
                                 void* scores, void* band, void* stream) {
    if (b == 0 || n_rows == 0) return CHK_OK;
    if (b < 0 || n_rows < 0 || rank < 2 || !q || !qn || !target || !entity || !hn || !shadow || !workspace || !counts || !scores || !band ||
        ((bh_vals == nullptr) != (bt == nullptr))) { chk_set_error("chk_score_all_mma: bad argument"); return CHK_EINVAL; }
    return rank_mma_launch(dtype, rank, b, q, qn, bh_vals, target, entity, hn, bt, n_rows, shadow, workspace, workspace_bytes,
                           counts, (float*)scores, (float*)band, (cudaStream_t)stream);
}
