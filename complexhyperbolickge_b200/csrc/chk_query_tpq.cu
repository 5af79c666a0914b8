// K1, thread-per-query variant (fp32, rank <= 33): forward of get_queries with ONE THREAD per query.
//
// Same contract and formulas as chk_query.cu (reference models/complexhyperbolic.py:79-171,
// utils/complexhyperbolic.py:36-106, utils/euclidean.py:26-75), different mapping.  The lane-group kernel spends
// ~330 warp-instructions per query at rank 33 (shuffle butterflies, shuffle reductions, scalar sections replicated
// over the 8 lanes of a query, runtime twiddles in registers) and is issue bound at 15 percent of the HBM roofline.
// Here a thread owns its whole query: the n/2-point complex FFTs run in registers with compile-time twiddles
// (constant-bank operands), every norm / dot is a thread-local reduction (no shuffles), the scalar sections are per
// thread and not replicated, and the packed pair IS the Givens pair as before.  The warp stages the 32 gathered entity
// rows through shared memory with coalesced loads (and the 32 output rows back with coalesced stores); relation rows
// are read with per-thread loads that collapse to broadcasts when the lanes of a warp share a relation (queries sorted
// or grouped by relation), and stay correct — just slower — when they do not.
#include "chk_common.cuh"

namespace {

// cos / sin (2 pi k / 64), k = 0..31: every twiddle of the transforms up to n = 64 is one of these
__device__ __constant__ float W64C[32] = {1.000000000e+00f, 9.951847267e-01f, 9.807852804e-01f, 9.569403357e-01f, 9.238795325e-01f, 8.819212643e-01f, 8.314696123e-01f, 7.730104534e-01f, 7.071067812e-01f, 6.343932842e-01f, 5.555702330e-01f, 4.713967368e-01f, 3.826834324e-01f, 2.902846773e-01f, 1.950903220e-01f, 9.801714033e-02f, 6.123233996e-17f, -9.801714033e-02f, -1.950903220e-01f, -2.902846773e-01f, -3.826834324e-01f, -4.713967368e-01f, -5.555702330e-01f, -6.343932842e-01f, -7.071067812e-01f, -7.730104534e-01f, -8.314696123e-01f, -8.819212643e-01f, -9.238795325e-01f, -9.569403357e-01f, -9.807852804e-01f, -9.951847267e-01f};
__device__ __constant__ float W64S[32] = {0.000000000e+00f, 9.801714033e-02f, 1.950903220e-01f, 2.902846773e-01f, 3.826834324e-01f, 4.713967368e-01f, 5.555702330e-01f, 6.343932842e-01f, 7.071067812e-01f, 7.730104534e-01f, 8.314696123e-01f, 8.819212643e-01f, 9.238795325e-01f, 9.569403357e-01f, 9.807852804e-01f, 9.951847267e-01f, 1.000000000e+00f, 9.951847267e-01f, 9.807852804e-01f, 9.569403357e-01f, 9.238795325e-01f, 8.819212643e-01f, 8.314696123e-01f, 7.730104534e-01f, 7.071067812e-01f, 6.343932842e-01f, 5.555702330e-01f, 4.713967368e-01f, 3.826834324e-01f, 2.902846773e-01f, 1.950903220e-01f, 9.801714033e-02f};

template <int LOGM> __device__ __forceinline__ constexpr int brev(int e) {
    int o = 0;
    for (int i = 0; i < LOGM; ++i) o |= ((e >> i) & 1) << (LOGM - 1 - i);
    return o;
}

// inverse (e^{+}) DIF, natural order in, bit-reversed order out, unnormalised; all indices compile-time
template <int LOGM> __device__ __forceinline__ void fft_dif_inv_tl(float (&a)[1 << LOGM], float (&b)[1 << LOGM]) {
    constexpr int M = 1 << LOGM;
#pragma unroll
    for (int half = M / 2; half >= 1; half >>= 1) {
#pragma unroll
        for (int blk = 0; blk < M; blk += 2 * half) {
#pragma unroll
            for (int j = 0; j < half; ++j) {
                const int i0 = blk + j, i1 = i0 + half, t = j * (32 / half);
                const float ar = a[i0], ai = b[i0], br = a[i1], bi = b[i1];
                a[i0] = ar + br; b[i0] = ai + bi;
                const float dr = ar - br, di = ai - bi;
                if (t == 0) { a[i1] = dr; b[i1] = di; }
                else { a[i1] = dr * W64C[t] - di * W64S[t]; b[i1] = dr * W64S[t] + di * W64C[t]; }
            }
        }
    }
}
// forward (e^{-}) DIT, bit-reversed order in, natural order out, unnormalised
template <int LOGM> __device__ __forceinline__ void fft_dit_fwd_tl(float (&a)[1 << LOGM], float (&b)[1 << LOGM]) {
    constexpr int M = 1 << LOGM;
#pragma unroll
    for (int half = 1; half < M; half <<= 1) {
#pragma unroll
        for (int blk = 0; blk < M; blk += 2 * half) {
#pragma unroll
            for (int j = 0; j < half; ++j) {
                const int i0 = blk + j, i1 = i0 + half, t = j * (32 / half);
                float br = a[i1], bi = b[i1];
                if (t != 0) { const float xr = br, xi = bi; br = xr * W64C[t] + xi * W64S[t]; bi = xi * W64C[t] - xr * W64S[t]; }
                const float ar = a[i0], ai = b[i0];
                a[i0] = ar + br; b[i0] = ai + bi;
                a[i1] = ar - br; b[i1] = ai - bi;
            }
        }
    }
}

// C2R (irfft, ortho): spectrum row [Re X_0..X_M | Im X_0..X_M] -> packed real pairs (x_{2j}, x_{2j+1}) held at
// register p = bitrev(j).  Same algebra as c2r_from_global in chk_query.cu with one lane per query.
template <int LOGM> __device__ __forceinline__ void c2r_tl(const float* __restrict__ row, float (&a)[1 << LOGM], float (&b)[1 << LOGM]) {
    constexpr int M = 1 << LOGM, R = M + 1, N = 2 * M;
    const float inv_sqrt_n = rsqrtf((float)N);
#pragma unroll
    for (int k = 0; k < M; ++k) {
        float ar = row[k], ai = row[R + k];
        float br = row[M - k], bi = -row[R + M - k];
        if (k == 0) { ai = 0.f; bi = 0.f; }
        const float sr = ar + br, si = ai + bi, dr = ar - br, di = ai - bi;
        const int t = k * (32 / M);                         // exp(2 pi i k / n), n = 2M
        a[k] = (sr - (dr * W64S[t] + di * W64C[t])) * inv_sqrt_n;
        b[k] = (si + (dr * W64C[t] - di * W64S[t])) * inv_sqrt_n;
    }
    fft_dif_inv_tl<LOGM>(a, b);
}
// R2C (rfft, ortho): packed pairs in the bit-reversed layout -> spectrum row
template <int LOGM> __device__ __forceinline__ void r2c_tl(float (&a)[1 << LOGM], float (&b)[1 << LOGM], float* __restrict__ row) {
    constexpr int M = 1 << LOGM, R = M + 1, N = 2 * M;
    fft_dit_fwd_tl<LOGM>(a, b);
    const float inv_sqrt_n = rsqrtf((float)N);
#pragma unroll
    for (int k = 0; k < M; ++k) {
        const int pk = (M - k) % M;
        const float zr = a[k], zi = b[k], pr = a[pk], pi = b[pk];
        const float er = zr + pr, ei = zi - pi, dr = zr - pr, di = zi + pi;
        const int t = k * (32 / M);
        const float scl = 0.5f * inv_sqrt_n;
        row[k] = (er + (di * W64C[t] - dr * W64S[t])) * scl;
        row[R + k] = (k == 0) ? 0.f : (ei - (dr * W64C[t] + di * W64S[t])) * scl;
    }
    row[M] = (a[0] - b[0]) * inv_sqrt_n;
    row[R + M] = 0.f;
}

struct TArgs {
    const float* entity; const float* rel; const float* rel_diag; const float* ctx; const float* c_table;
    const int64_t* head_idx; const int64_t* rel_idx; const int32_t* perm;      // perm: optional processing order (sorted by relation)
    int64_t nq; int multi_c;
    float* out_q; float* out_c;
};

__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }

// expmap0 + project as ONE scale factor for a vector of squared norm n2: gamma = f * u with
// f = tanh(clamp(sc |u|)) / (sc |u|), then pulled back onto the ball if needed (utils/complexhyperbolic.py:41-54,72-87)
__device__ __forceinline__ float expmap_factor(float n2, float sc) {
    const float nu = fmaxf(sqrtf(n2), Sc<float>::min_norm);
    const float aa = sc * nu;
    float f = tanhf(fminf(fmaxf(aa, -15.f), 15.f)) / aa;
    const float nrm = fmaxf(f * nu, Sc<float>::min_norm), maxnorm = Sc<float>::proj_top / sc;
    if (nrm > maxnorm) f *= maxnorm / nrm;
    return f;
}

constexpr int TPQ_WARPS = 4;

template <int LOGM, int KIND>
__global__ void __launch_bounds__(TPQ_WARPS * 32) query_tpq_kernel(TArgs A) {
    constexpr int M = 1 << LOGM, N = 2 * M, R = M + 1, ROWW = 2 * R, PITCH = ROWW + 1;
    constexpr int RDW = (KIND == CHK_ATT) ? 2 * N : N;
    extern __shared__ float smem_f[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* stage2 = smem_f + (size_t)warp * 2 * 32 * PITCH;         // double buffer: 2 x 32 rows x (2r + 1 pad)
    const int64_t n_batches = (A.nq + 31) / 32;
    const int64_t wstep = (int64_t)gridDim.x * TPQ_WARPS;
    // gather of the 32 entity rows of a warp batch: asynchronous 4-byte copies (coalesced 128 B per warp instruction),
    // all in flight at once, no registers; the next batch is fetched while the current one is computed
    auto prefetch = [&](int64_t wb, float* stage, int64_t& qi, int64_t& rl, bool& ok) {
        const int64_t pos = wb * 32 + lane;
        ok = wb < n_batches && pos < A.nq;
        qi = ok ? (A.perm ? (int64_t)A.perm[pos] : pos) : 0;
        const int64_t h = A.head_idx[qi];
        rl = A.rel_idx[qi];
        if (wb < n_batches) {
#pragma unroll 8
            for (int p = 0; p < 32; ++p) {
                const int64_t hp = __shfl_sync(CHK_FULL, h, p);
                const float* src = A.entity + hp * ROWW;
#pragma unroll
                for (int k0 = 0; k0 < ROWW; k0 += 32)
                    if (k0 + lane < ROWW) cp_async_4(stage + p * PITCH + k0 + lane, src + k0 + lane);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int64_t wb = (int64_t)blockIdx.x * TPQ_WARPS + warp;
    int64_t qi_n = 0, rl_n = 0; bool ok_n = false;
    int buf = 0;
    prefetch(wb, stage2, qi_n, rl_n, ok_n);
    for (; wb < n_batches; wb += wstep, buf ^= 1) {
        float* stage = stage2 + buf * 32 * PITCH;
        const int64_t qi = qi_n, rl = rl_n; const bool ok = ok_n;
        __syncwarp();                                               // the other buffer's output rows have been stored
        prefetch(wb + wstep, stage2 + (buf ^ 1) * 32 * PITCH, qi_n, rl_n, ok_n);
        asm volatile("cp.async.wait_group 1;" ::: "memory");        // this batch's rows have landed
        __syncwarp();
        float a[M], b[M];
        c2r_tl<LOGM>(stage + lane * PITCH, a, b);                   // u, packed pairs in the bit-reversed layout
        // ---- per-thread hyperbolic section
        const float c_raw = A.multi_c ? A.c_table[rl] : A.c_table[0];
        const float c = A.multi_c ? softplus_f(c_raw) : c_raw;
        const float sc = sqrtf(c);
        const float2* relrow = reinterpret_cast<const float2*>(A.rel + rl * (2 * N));      // rows are 8-byte aligned (n even)
        const float2* rdrow = reinterpret_cast<const float2*>(A.rel_diag + rl * RDW);
        if (KIND == CHK_ROT) {
            float nu2 = 0.f, n1 = 0.f, d1 = 0.f;
#pragma unroll
            for (int p = 0; p < M; ++p) {
                const float2 t = __ldg(relrow + brev<LOGM>(p));
                nu2 = fmaf(a[p], a[p], fmaf(b[p], b[p], nu2));
                n1 = fmaf(t.x, t.x, fmaf(t.y, t.y, n1));
                d1 = fmaf(a[p], t.x, fmaf(b[p], t.y, d1));
            }
            const float f = expmap_factor(nu2, sc), f1 = expmap_factor(n1, sc);
            float x2 = f * f * nu2, y2 = f1 * f1 * n1, xy = f * f1 * d1;
            float Aq = 1.f + 2.f * c * xy + c * y2, Bq = 1.f - c * x2;
            float den = fmaxf(1.f + 2.f * c * xy + c * c * x2 * y2, Sc<float>::min_norm);
            float ka = Aq / den * f, kb = Bq / den * f1;            // m1 = ka * u + kb * t1
            float m2 = 0.f;
#pragma unroll
            for (int p = 0; p < M; ++p) {
                const float2 t = __ldg(relrow + brev<LOGM>(p));
                a[p] = fmaf(ka, a[p], kb * t.x); b[p] = fmaf(ka, b[p], kb * t.y);
                m2 = fmaf(a[p], a[p], fmaf(b[p], b[p], m2));
            }
            // project, then Givens rotation by the normalised pairs of rel_diag, and the dots with the second translation
            const float nrm = fmaxf(sqrtf(m2), Sc<float>::min_norm), maxnorm = Sc<float>::proj_top / sc;
            const float pj = nrm > maxnorm ? maxnorm / nrm : 1.f;
            float n2 = 0.f, d2 = 0.f;
#pragma unroll
            for (int p = 0; p < M; ++p) {
                const float2 g = __ldg(rdrow + brev<LOGM>(p));
                const float2 t = __ldg(relrow + M + brev<LOGM>(p));
                const float inv = pj * rsqrtf(g.x * g.x + g.y * g.y);
                const float g0 = g.x * inv, g1 = g.y * inv, x0 = a[p], x1 = b[p];
                a[p] = g0 * x0 - g1 * x1; b[p] = g0 * x1 + g1 * x0;
                n2 = fmaf(t.x, t.x, fmaf(t.y, t.y, n2));
                d2 = fmaf(a[p], t.x, fmaf(b[p], t.y, d2));
            }
            const float f2 = expmap_factor(n2, sc);
            x2 = pj * pj * m2; y2 = f2 * f2 * n2; xy = f2 * d2;
            Aq = 1.f + 2.f * c * xy + c * y2; Bq = 1.f - c * x2;
            den = fmaxf(1.f + 2.f * c * xy + c * c * x2 * y2, Sc<float>::min_norm);
            ka = Aq / den; kb = Bq / den * f2;
#pragma unroll
            for (int p = 0; p < M; ++p) {
                const float2 t = __ldg(relrow + M + brev<LOGM>(p));
                a[p] = fmaf(ka, a[p], kb * t.x); b[p] = fmaf(ka, b[p], kb * t.y);
            }
        } else {
            // REF: pre = refl(rel_diag, u); ATT: pre = w_ref * refl(rel_diag[n:], u) + w_rot * rot(rel_diag[:n], u)
            float w_ref = 1.f, w_rot = 0.f;
            if (KIND == CHK_ATT) {
                const float2* cxrow = reinterpret_cast<const float2*>(A.ctx + rl * N);
                const float scale = rsqrtf((float)R);                              // 1/sqrt(rank), models/complexhyperbolic.py:138
                float l_ref = 0.f, l_rot = 0.f;
#pragma unroll
                for (int p = 0; p < M; ++p) {
                    const int j = brev<LOGM>(p);
                    const float2 gr = __ldg(rdrow + j), gf = __ldg(rdrow + M + j), cx = __ldg(cxrow + j);
                    const float ir = rsqrtf(gr.x * gr.x + gr.y * gr.y), jf = rsqrtf(gf.x * gf.x + gf.y * gf.y);
                    const float r0 = gr.x * ir, r1 = gr.y * ir, f0 = gf.x * jf, f1_ = gf.y * jf, x0 = a[p], x1 = b[p];
                    const float rot0 = r0 * x0 - r1 * x1, rot1 = r0 * x1 + r1 * x0;
                    const float ref0 = f0 * x0 + f1_ * x1, ref1 = f0 * (-x0) + f1_ * x0;
                    l_ref += (cx.x * ref0) * scale + (cx.y * ref1) * scale;
                    l_rot += (cx.x * rot0) * scale + (cx.y * rot1) * scale;
                }
                const float mx = fmaxf(l_ref, l_rot), e_ref = expf(l_ref - mx), e_rot = expf(l_rot - mx);
                const float inv = 1.f / (e_ref + e_rot);
                w_ref = e_ref * inv; w_rot = e_rot * inv;
            }
            float p2 = 0.f, n1 = 0.f, d1 = 0.f;
#pragma unroll
            for (int p = 0; p < M; ++p) {
                const int j = brev<LOGM>(p);
                const float2 gf = __ldg(rdrow + (KIND == CHK_ATT ? M : 0) + j);
                const float jf = rsqrtf(gf.x * gf.x + gf.y * gf.y);
                const float f0 = gf.x * jf, f1_ = gf.y * jf, x0 = a[p], x1 = b[p];
                float o0 = f0 * x0 + f1_ * x1, o1 = f0 * (-x0) + f1_ * x0;     // "reflection" as coded (utils/euclidean.py:60-75)
                if (KIND == CHK_ATT) {
                    const float2 gr = __ldg(rdrow + j);
                    const float ir = rsqrtf(gr.x * gr.x + gr.y * gr.y), r0 = gr.x * ir, r1 = gr.y * ir;
                    o0 = w_ref * o0 + w_rot * (r0 * x0 - r1 * x1);
                    o1 = w_ref * o1 + w_rot * (r0 * x1 + r1 * x0);
                }
                a[p] = o0; b[p] = o1;
                const float2 t = __ldg(relrow + j);
                p2 = fmaf(o0, o0, fmaf(o1, o1, p2));
                n1 = fmaf(t.x, t.x, fmaf(t.y, t.y, n1));
                d1 = fmaf(o0, t.x, fmaf(o1, t.y, d1));
            }
            const float f = expmap_factor(p2, sc), f1 = expmap_factor(n1, sc);
            const float x2 = f * f * p2, y2 = f1 * f1 * n1, xy = f * f1 * d1;
            const float Aq = 1.f + 2.f * c * xy + c * y2, Bq = 1.f - c * x2;
            const float den = fmaxf(1.f + 2.f * c * xy + c * c * x2 * y2, Sc<float>::min_norm);
            const float ka = Aq / den * f, kb = Bq / den * f1;
            float m2 = 0.f;
#pragma unroll
            for (int p = 0; p < M; ++p) {
                const float2 t = __ldg(relrow + brev<LOGM>(p));
                a[p] = fmaf(ka, a[p], kb * t.x); b[p] = fmaf(ka, b[p], kb * t.y);
                m2 = fmaf(a[p], a[p], fmaf(b[p], b[p], m2));
            }
            const float nrm = fmaxf(sqrtf(m2), Sc<float>::min_norm), maxnorm = Sc<float>::proj_top / sc;
            if (nrm > maxnorm) {
                const float pj = maxnorm / nrm;
#pragma unroll
                for (int p = 0; p < M; ++p) { a[p] *= pj; b[p] *= pj; }
            }
        }
        // ---- rfft into the staging row (its input row is consumed), coalesced stores of the 32 rows
        __syncwarp();
        r2c_tl<LOGM>(a, b, stage + lane * PITCH);
        if (ok) A.out_c[qi] = c;
        __syncwarp();
#pragma unroll 8
        for (int p = 0; p < 32; ++p) {
            const int64_t qp = __shfl_sync(CHK_FULL, qi, p);
            const int okp = __shfl_sync(CHK_FULL, (int)ok, p);
            if (okp) {
                float* dst = A.out_q + qp * ROWW;
#pragma unroll
                for (int k0 = 0; k0 < ROWW; k0 += 32)
                    if (k0 + lane < ROWW) dst[k0 + lane] = stage[p * PITCH + k0 + lane];
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <int LOGM, int KIND>
int launch_tpq(const TArgs& A, cudaStream_t st) {
    constexpr int PITCH = 2 * ((1 << LOGM) + 1) + 1;
    const size_t smem = (size_t)TPQ_WARPS * 2 * 32 * PITCH * sizeof(float);
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(query_tpq_kernel<LOGM, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_done = true;
    }
    int64_t blocks = (A.nq + TPQ_WARPS * 32 - 1) / (TPQ_WARPS * 32);
    if (blocks > 148 * 8) blocks = 148 * 8;
    query_tpq_kernel<LOGM, KIND><<<(unsigned)blocks, TPQ_WARPS * 32, smem, st>>>(A);
    CHK_CUDA_LAUNCH_CHECK("query_tpq_kernel");
    return CHK_OK;
}

template <int KIND>
int dispatch_tpq(int rank, const TArgs& A, cudaStream_t st) {
    switch (rank) {
        case 9: return launch_tpq<3, KIND>(A, st);
        case 17: return launch_tpq<4, KIND>(A, st);
        case 33: return launch_tpq<5, KIND>(A, st);
        default: chk_set_error("thread-per-query K1: rank %d unsupported", rank); return CHK_EUNSUPPORTED;
    }
}

// ---- counting sort of the query positions by relation id (the processing order of the kernel above) ----------------
__global__ void __launch_bounds__(256) group_hist_kernel(const int64_t* __restrict__ keys, int64_t n, int n_keys, int* __restrict__ counts) {
    extern __shared__ int sh[];
    for (int i = threadIdx.x; i < n_keys; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) atomicAdd(&sh[(int)keys[i]], 1);
    __syncthreads();
    for (int i = threadIdx.x; i < n_keys; i += blockDim.x) if (sh[i]) atomicAdd(&counts[i], sh[i]);
}
__global__ void __launch_bounds__(1024) group_scan_kernel(int* __restrict__ counts, int n_keys) {       // exclusive scan in place, one block
    __shared__ int part[1024];
    const int per = (n_keys + 1023) / 1024, lo = threadIdx.x * per, hi = min(lo + per, n_keys);
    int s = 0;
    for (int i = lo; i < hi; ++i) s += counts[i];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        int v = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    int run = part[threadIdx.x] - s;
    for (int i = lo; i < hi; ++i) { int c = counts[i]; counts[i] = run; run += c; }
}
__global__ void __launch_bounds__(256) group_scatter_kernel(const int64_t* __restrict__ keys, int64_t n, int n_keys, int* __restrict__ offsets,
                                                            int32_t* __restrict__ perm) {
    // block-local ranks in shared memory, then ONE global atomic per (block, key) to reserve the block's slots
    extern __shared__ int sh[];                      // [n_keys] local counts, then [n_keys] global bases
    int* base = sh + n_keys;
    const int64_t chunk = (int64_t)blockDim.x * 8;
    for (int64_t c0 = (int64_t)blockIdx.x * chunk; c0 < n; c0 += (int64_t)gridDim.x * chunk) {
        for (int i = threadIdx.x; i < n_keys; i += blockDim.x) sh[i] = 0;
        __syncthreads();
        int key[8], rank_in_blk[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int64_t i = c0 + threadIdx.x + (int64_t)u * blockDim.x;
            key[u] = i < n ? (int)keys[i] : -1;
            rank_in_blk[u] = key[u] >= 0 ? atomicAdd(&sh[key[u]], 1) : 0;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < n_keys; i += blockDim.x) base[i] = sh[i] ? atomicAdd(&offsets[i], sh[i]) : 0;
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int64_t i = c0 + threadIdx.x + (int64_t)u * blockDim.x;
            if (key[u] >= 0) perm[base[key[u]] + rank_in_blk[u]] = (int32_t)i;
        }
        __syncthreads();
    }
}

}  // namespace

// perm = positions 0..n-1 grouped by key (any order inside a group); counts: int32 [n_keys] scratch.  n < 2^31, n_keys <= 12288.
int chk_group_by_key_impl(const int64_t* keys, int64_t n, int n_keys, int32_t* perm, int32_t* counts, cudaStream_t st) {
    if (n_keys < 1 || n_keys > 12288 || n >= (1LL << 31)) { chk_set_error("chk_group_by_key: n_keys in [1,12288], n < 2^31"); return CHK_EUNSUPPORTED; }
    if (cudaMemsetAsync(counts, 0, sizeof(int32_t) * n_keys, st) != cudaSuccess) { chk_set_error("cudaMemsetAsync failed"); return CHK_ECUDA; }
    int64_t blocks = (n + 2047) / 2048; if (blocks > 148 * 4) blocks = 148 * 4; if (blocks < 1) blocks = 1;
    group_hist_kernel<<<(unsigned)blocks, 256, n_keys * sizeof(int), st>>>(keys, n, n_keys, counts);
    group_scan_kernel<<<1, 1024, 0, st>>>(counts, n_keys);
    group_scatter_kernel<<<(unsigned)blocks, 256, 2 * n_keys * sizeof(int), st>>>(keys, n, n_keys, counts, perm);
    CHK_CUDA_LAUNCH_CHECK("group_by_key kernels");
    return CHK_OK;
}

// fp32, rank in {9, 17, 33}; perm may be NULL.  Called by chk_query_fwd (chk_query.cu).
int chk_query_fwd_tpq(int kind, int rank, int64_t nq, int multi_c, const void* entity, const void* rel, const void* rel_diag,
                      const void* ctx, const void* c_table, const int64_t* head_idx, const int64_t* rel_idx,
                      const int32_t* perm, void* out_q, void* out_c, cudaStream_t st) {
    TArgs A{(const float*)entity, (const float*)rel, (const float*)rel_diag, (const float*)ctx, (const float*)c_table,
            head_idx, rel_idx, perm, nq, multi_c, (float*)out_q, (float*)out_c};
    switch (kind) {
        case CHK_ROT: return dispatch_tpq<CHK_ROT>(rank, A, st);
        case CHK_REF: return dispatch_tpq<CHK_REF>(rank, A, st);
        case CHK_ATT:
            if (!ctx) { chk_set_error("FFTAttH needs context_vec"); return CHK_EINVAL; }
            return dispatch_tpq<CHK_ATT>(rank, A, st);
        default: chk_set_error("unknown model kind %d", kind); return CHK_EINVAL;
    }
}
