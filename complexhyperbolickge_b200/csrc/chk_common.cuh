// Shared device helpers for the chk_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/chk_b200.h"

#define CHK_FULL 0xffffffffu

void chk_set_error(const char* fmt, ...);

#define CHK_CUDA_LAUNCH_CHECK(name)                                              \
    do {                                                                         \
        cudaError_t e__ = cudaGetLastError();                                    \
        if (e__ != cudaSuccess) {                                                \
            chk_set_error("%s: %s", name, cudaGetErrorString(e__));              \
            return CHK_ECUDA;                                                    \
        }                                                                        \
    } while (0)

// ---- scalar traits -------------------------------------------------------------------------
template <typename T> struct Sc;
template <> struct Sc<float> {
    static constexpr float ball_eps = 4e-3f;      // utils/complexhyperbolic.py:13
    static constexpr float min_norm = 1e-15f;     // :12
    static constexpr float proj_top = (float)(1.0 - 1e-5);   // python double 1-1e-5 cast to fp32, :83-84
    __device__ static __forceinline__ float sqrt_(float x) { return sqrtf(x); }
    __device__ static __forceinline__ float tanh_(float x) { return tanhf(x); }
    __device__ static __forceinline__ float exp_(float x) { return expf(x); }
    __device__ static __forceinline__ float log1p_(float x) { return log1pf(x); }
    __device__ static __forceinline__ float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }
    __device__ static __forceinline__ float mul_(float a, float b) { return __fmul_rn(a, b); }   // never contracted
    __device__ static __forceinline__ float add_(float a, float b) { return __fadd_rn(a, b); }
    __device__ static __forceinline__ float max_(float a, float b) { return fmaxf(a, b); }
    __device__ static __forceinline__ float min_(float a, float b) { return fminf(a, b); }
    __device__ static __forceinline__ float abs_(float a) { return fabsf(a); }
};
template <> struct Sc<double> {
    static constexpr double ball_eps = 1e-5;
    static constexpr double min_norm = 1e-15;
    static constexpr double proj_top = 1.0 - 1e-5;
    __device__ static __forceinline__ double sqrt_(double x) { return sqrt(x); }
    __device__ static __forceinline__ double tanh_(double x) { return tanh(x); }
    __device__ static __forceinline__ double exp_(double x) { return exp(x); }
    __device__ static __forceinline__ double log1p_(double x) { return log1p(x); }
    __device__ static __forceinline__ double fma_(double a, double b, double c) { return __fma_rn(a, b, c); }
    __device__ static __forceinline__ double mul_(double a, double b) { return __dmul_rn(a, b); }
    __device__ static __forceinline__ double add_(double a, double b) { return __dadd_rn(a, b); }
    __device__ static __forceinline__ double max_(double a, double b) { return fmax(a, b); }
    __device__ static __forceinline__ double min_(double a, double b) { return fmin(a, b); }
    __device__ static __forceinline__ double abs_(double a) { return fabs(a); }
};

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(CHK_FULL, v, o);
    return v;
}

template <typename T>
__device__ __forceinline__ void warp_sum3(T& a, T& b, T& c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T ta = __shfl_xor_sync(CHK_FULL, a, o);
        T tb = __shfl_xor_sync(CHK_FULL, b, o);
        T tc = __shfl_xor_sync(CHK_FULL, c, o);
        a += ta; b += tb; c += tc;
    }
}

// ---- the canonical pair score ----------------------------------------------------------------
// Everything that must agree bit-for-bit between kernels (target score, rank tiles, filter pass,
// exact re-check of the tensor-core tier) goes through these two functions.
//
// Hermitian norm of a row sum, clamped: utils/complexhyperbolic.py:229-230.
template <typename T>
__device__ __forceinline__ T clamp_hnorm(T sumsq) {
    T v = sumsq - T(1);
    v = Sc<T>::max_(v, T(-1));
    return Sc<T>::min_(v, -Sc<T>::ball_eps);
}

// re = Re sum z conj(w) (WITHOUT the -1), im likewise; zn, wn clamped norms.
// x = 2|zw|^2/zn/wn - 1 clamped at 1+eps (utils/complexhyperbolic.py:231-234), d = acosh(x).
template <typename T>
__device__ __forceinline__ T clamped_x(T re, T im, T zn, T wn) {
    T r1 = re - T(1);
    T mod2 = Sc<T>::fma_(r1, r1, im * im);
    T x = (T(2) * mod2) / zn / wn - T(1);
    return Sc<T>::max_(x, T(1) + Sc<T>::ball_eps);
}

template <typename T>
__device__ __forceinline__ T acosh_x(T x) {
    // acosh(x) = log1p((x-1) + sqrt((x-1)(x+1))): accurate for x -> 1+
    T xm = x - T(1);
    return Sc<T>::log1p_(xm + Sc<T>::sqrt_(xm * (x + T(1))));
}

// score = (bh + bt) + (-d^2)  in that order (models/base.py:171); has_bias==false -> -d^2.
// The square and the two adds are separately rounded (mul_/add_ are never contracted into an FMA), as
// in the reference's eager ops, and so that every kernel produces the same bits from the same (x, bh, bt).
template <typename T>
__device__ __forceinline__ T score_from_x(T x, bool has_bias, T bh, T bt) {
    T d = acosh_x(x);
    T s = -Sc<T>::mul_(d, d);
    return has_bias ? Sc<T>::add_(Sc<T>::add_(bh, bt), s) : s;
}

template <typename T>
__device__ __forceinline__ T pair_score(T re, T im, T zn, T wn, bool has_bias, T bh, T bt) {
    return score_from_x<T>(clamped_x(re, im, zn, wn), has_bias, bh, bt);
}

// Canonical dot chain over complex index k (zr,zi,wr,wi are the four real planes of the two rows).
template <typename T>
__device__ __forceinline__ void dot_step(T zr, T zi, T wr, T wi, T& re, T& im) {
    re = Sc<T>::fma_(zr, wr, re);
    re = Sc<T>::fma_(zi, wi, re);
    im = Sc<T>::fma_(zi, wr, im);
    im = Sc<T>::fma_(-zr, wi, im);
}

// ---- the canonical summation order ------------------------------------------------------------------
// fp64: one ascending-k FMA chain.  fp32: BLOCKED — the chain restarts every CHK_BLK = 16 complex coefficients and the
// block partials are added (separately rounded) in ascending order.  Every exact kernel (tile kernel, target,
// filter pass, re-check) uses the same order, so their pair scores are bit-identical; the blocked order brings the
// worst-case rounding bound of the fp32 dot from 2r units down to 2*16 + r/16 (514 -> 49 at rank 257), which is
// what the tensor-core tier's error band (chk_rank_mma.cu) is built from.
template <typename T> struct Chain { static constexpr int BLK = 0; };
template <> struct Chain<float> { static constexpr int BLK = 16; };

// ---- exact per-pair score (canonical chain, one thread per pair) -----------------------------------
template <typename T>
__device__ __forceinline__ T exact_pair(const T* __restrict__ z, const T* __restrict__ w, int r, T zn, T wn,
                                        bool has_bias, T bh, T bt) {
    T re = T(0), im = T(0);
    if (Chain<T>::BLK == 0) {
        for (int k = 0; k < r; ++k) dot_step<T>(z[k], z[r + k], w[k], w[r + k], re, im);
    } else {
        for (int k0 = 0; k0 < r; k0 += Chain<T>::BLK) {
            T pr = T(0), pi = T(0);
            const int k1 = min(k0 + Chain<T>::BLK, r);
            for (int k = k0; k < k1; ++k) dot_step<T>(z[k], z[r + k], w[k], w[r + k], pr, pi);
            re = Sc<T>::add_(re, pr); im = Sc<T>::add_(im, pi);
        }
    }
    return pair_score<T>(re, im, zn, wn, has_bias, bh, bt);
}

// Arguments shared by the K2 kernels (exact tier, filter pass, tensor-core tier re-check).
template <typename T> struct RArgs {
    const T* q; const T* qn; const T* bh_vals; const T* target;
    const T* entity; const T* hn; const T* bt;
    int64_t b, n_rows; int r;
    T* scores;                       // MODE 0: [b, n_rows]
    unsigned long long* counts;      // MODE 1: [b]
};

// ---- exact per-pair scores, one LANE per pair, rows staged by the whole warp --------------------------
// Same arithmetic and order as exact_pair() (each lane walks k = 0..r-1 of its own pair), but the 2r-wide
// rows are fetched with coalesced 128-byte warp loads, 32 complex coefficients at a time, through a padded
// shared-memory transpose.  Used wherever a list of scattered (query, entity) pairs must be scored exactly:
// the filter pass and the re-check of the tensor-core tier.
template <typename T> struct PairTiles { T zr[32][33], zi[32][33], wr[32][33], wi[32][33]; };
// fp32: WHOLE rows of 8 pairs are staged per pass (consecutive copies walk one row, so DRAM sees 2 KB runs instead of
// scattered 128-byte pieces); every canonical 16-coefficient block sits in a 20-float slot and rows are 688 floats
// apart (= 16 mod 32), which makes the 128-bit reads of a quarter-warp (2 pairs x 4 lanes) hit 32 distinct banks.
template <> struct __align__(16) PairTiles<float> {
    static constexpr int BS = 20;                    // floats per staged block
    static constexpr int PL = 17 * BS;               // one plane (Re or Im): 16 full blocks + the tail block (rank <= 257)
    static constexpr int RS = 688;                   // row stride (2 * PL = 680, padded)
    static constexpr int SLOTS = 8;                  // pairs per pass
    float z[SLOTS][RS], w[SLOTS][RS];
};

__device__ __forceinline__ void cp_async_4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// fp64 (and the generic fallback): 32-coefficient chunks, single-buffered.
template <typename T>
__device__ __forceinline__ T warp_exact_pairs(const RArgs<T>& A, unsigned i, unsigned e, bool valid, PairTiles<T>& S) {
    const int lane = threadIdx.x & 31;
    const int r = A.r;
    T re = T(0), im = T(0);
    for (int k0 = 0; k0 < r; k0 += 32) {
        const int kc = min(32, r - k0);
        __syncwarp();
        // all 32 x 4 row segments of this chunk go out as asynchronous copies before anything is waited for
#pragma unroll 8
        for (int p = 0; p < 32; ++p) {
            const unsigned ip = __shfl_sync(CHK_FULL, i, p), ep = __shfl_sync(CHK_FULL, e, p);
            const int vp = __shfl_sync(CHK_FULL, (int)valid, p);
            if (vp && lane < kc) {
                const T* z = A.q + (size_t)ip * 2 * r + k0 + lane;
                const T* w = A.entity + (size_t)ep * 2 * r + k0 + lane;
                cp_async_8(&S.zr[p][lane], z); cp_async_8(&S.zi[p][lane], z + r);
                cp_async_8(&S.wr[p][lane], w); cp_async_8(&S.wi[p][lane], w + r);
            }
        }
        cp_async_wait_all();
        __syncwarp();
        if (valid)
            for (int kk = 0; kk < kc; ++kk) dot_step<T>(S.zr[lane][kk], S.zi[lane][kk], S.wr[lane][kk], S.wi[lane][kk], re, im);
    }
    if (!valid) return T(0);
    const bool has_bias = A.bt != nullptr;
    return pair_score<T>(re, im, A.qn[i], A.hn[e], has_bias, has_bias ? A.bh_vals[i] : T(0), has_bias ? A.bt[e] : T(0));
}

// fp32: the canonical order is BLOCKED (Chain<float>::BLK = 16): the blocks of a pair are independent chains whose
// partial sums are added in ascending block order.  A pass handles 8 pairs; FOUR lanes share a pair, lane `sub` walks
// blocks sub, sub+4, ... and after every round of four blocks the partials are folded into the running sum in block
// order (shuffles), so the bits equal exact_pair()'s.  Results return to the lane that owns the pair.
// R > 0: the rank is a compile-time constant (every loop unrolls, the copies use immediate offsets); R = 0: runtime rank.
template <int R>
__device__ __forceinline__ void exact_pairs_f32(const RArgs<float>& A, unsigned i, unsigned e, bool valid, PairTiles<float>& S,
                                                float& my_re, float& my_im) {
    using PT = PairTiles<float>;
    constexpr int BLK = Chain<float>::BLK, BS = PT::BS, PL = PT::PL;
    static_assert(BLK == 16, "staging assumes 16-coefficient blocks");
    const int lane = threadIdx.x & 31;
    const int r = R ? R : A.r;
    const int nfull = r / BLK, tail = r - nfull * BLK;
    const int c_lane = (lane >> 4) * BS + (lane & 15);                  // staged position of coefficient k0 + lane (k0 % 32 == 0)
    for (int pass = 0; pass < 4; ++pass) {
        if (!__any_sync(CHK_FULL, valid && (lane >> 3) == pass)) continue;
        __syncwarp();                                                   // readers of the previous pass / call are done with S
#pragma unroll
        for (int sl = 0; sl < PT::SLOTS; ++sl) {
            const int src = pass * 8 + sl;
            const unsigned ip = __shfl_sync(CHK_FULL, i, src), ep = __shfl_sync(CHK_FULL, e, src);
            if (!__shfl_sync(CHK_FULL, (int)valid, src)) continue;
            const float* zrow = A.q + (size_t)ip * 2 * r + lane;
            const float* wrow = A.entity + (size_t)ep * 2 * r + lane;
            float* zd = &S.z[sl][c_lane];
            float* wd = &S.w[sl][c_lane];
#pragma unroll
            for (int k0 = 0; k0 < r; k0 += 32) {                        // Re plane, then Im plane: one contiguous run per row
                const int d = (k0 >> 5) * 2 * BS;
                if (k0 + 32 <= r || k0 + lane < r) { cp_async_4(wd + d, wrow + k0); cp_async_4(zd + d, zrow + k0); }
            }
#pragma unroll
            for (int k0 = 0; k0 < r; k0 += 32) {
                const int d = PL + (k0 >> 5) * 2 * BS;
                if (k0 + 32 <= r || k0 + lane < r) { cp_async_4(wd + d, wrow + r + k0); cp_async_4(zd + d, zrow + r + k0); }
            }
        }
        cp_async_wait_all();
        __syncwarp();
        const int sl = lane >> 2, sub = lane & 3, base = lane & ~3;
        const float* zb = S.z[sl];
        const float* wb = S.w[sl];
        float re = 0.f, im = 0.f;
#pragma unroll
        for (int b0 = 0; b0 < nfull; b0 += 4) {
            const int b = b0 + sub;
            float pr = 0.f, pi = 0.f;
            if (b < nfull) {
                const float* zq = zb + b * BS;
                const float* wq = wb + b * BS;
#pragma unroll
                for (int k4 = 0; k4 < BLK; k4 += 4) {
                    const float4 zr = *reinterpret_cast<const float4*>(zq + k4);
                    const float4 zi = *reinterpret_cast<const float4*>(zq + PL + k4);
                    const float4 wr = *reinterpret_cast<const float4*>(wq + k4);
                    const float4 wi = *reinterpret_cast<const float4*>(wq + PL + k4);
                    dot_step<float>(zr.x, zi.x, wr.x, wi.x, pr, pi);
                    dot_step<float>(zr.y, zi.y, wr.y, wi.y, pr, pi);
                    dot_step<float>(zr.z, zi.z, wr.z, wi.z, pr, pi);
                    dot_step<float>(zr.w, zi.w, wr.w, wi.w, pr, pi);
                }
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {                               // fold the four partials in ascending block order
                const float tr = __shfl_sync(CHK_FULL, pr, base + t), ti = __shfl_sync(CHK_FULL, pi, base + t);
                if (b0 + t < nfull) { re = Sc<float>::add_(re, tr); im = Sc<float>::add_(im, ti); }
            }
        }
        if (tail > 0) {                                                 // the last (short) block, by all four lanes alike
            float pr = 0.f, pi = 0.f;
            const float* zq = zb + nfull * BS;
            const float* wq = wb + nfull * BS;
            for (int kk = 0; kk < tail; ++kk) dot_step<float>(zq[kk], zq[PL + kk], wq[kk], wq[PL + kk], pr, pi);
            re = Sc<float>::add_(re, pr); im = Sc<float>::add_(im, pi);
        }
        const float vr = __shfl_sync(CHK_FULL, re, (lane & 7) * 4), vi = __shfl_sync(CHK_FULL, im, (lane & 7) * 4);
        if ((lane >> 3) == pass) { my_re = vr; my_im = vi; }
    }
}

template <>
__device__ __forceinline__ float warp_exact_pairs<float>(const RArgs<float>& A, unsigned i, unsigned e, bool valid, PairTiles<float>& S) {
    float my_re = 0.f, my_im = 0.f;
    switch (A.r) {                                   // the supported ranks are 2^m + 1
        case 33: exact_pairs_f32<33>(A, i, e, valid, S, my_re, my_im); break;
        case 65: exact_pairs_f32<65>(A, i, e, valid, S, my_re, my_im); break;
        case 129: exact_pairs_f32<129>(A, i, e, valid, S, my_re, my_im); break;
        case 257: exact_pairs_f32<257>(A, i, e, valid, S, my_re, my_im); break;
        default: exact_pairs_f32<0>(A, i, e, valid, S, my_re, my_im); break;
    }
    if (!valid) return 0.f;
    const bool has_bias = A.bt != nullptr;
    return pair_score<float>(my_re, my_im, A.qn[i], A.hn[e], has_bias, has_bias ? A.bh_vals[i] : 0.f, has_bias ? A.bt[e] : 0.f);
}
