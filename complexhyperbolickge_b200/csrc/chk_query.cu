// K1 — query transform (get_queries) forward and backward, one lane-group per query, all in registers.
//
// Replaces FFTRotH/FFTRefH/FFTAttH.get_queries (reference models/complexhyperbolic.py:79-101,107-127,
// 144-171) and the ops it calls (utils/complexhyperbolic.py:36-54,72-106; utils/euclidean.py:26-75),
// plus their autograd adjoints.
//
// Design (B200): a real length-n (i)FFT is a length-M = n/2 complex FFT of the packed pairs
// z_j = x_{2j} + i x_{2j+1} plus an O(n) split / merge pass.  A group of L lanes owns one query and
// holds P = M/L complex points per lane (element e = p*L + lane_in_group).  Stages with span >= L are
// in-register butterflies, the rest are __shfl_xor butterflies.  The C2R transform is a DIF FFT
// (natural -> bit-reversed), every element-wise / norm / Moebius / Givens step runs in the bit-reversed
// layout (the packed pair IS the Givens pair), and the R2C transform is a DIT FFT (bit-reversed ->
// natural).  No shared memory, no intermediate tensor in HBM.  Twiddles are computed once per lane with
// sincospi in double and live in registers for the whole persistent loop.
#include <cstdlib>
#include "chk_common.cuh"

namespace {

template <int LOGM, int LOGL> struct Geo {
    static constexpr int M = 1 << LOGM;      // complex points = n/2
    static constexpr int L = 1 << LOGL;      // lanes per query
    static constexpr int P = M / L;          // points per lane
    static constexpr int N = 2 * M;          // real length n = dim
    static constexpr int R = M + 1;          // rank
    static constexpr int QPW = 32 / L;       // queries per warp
    static_assert(P >= 1, "lanes per query must not exceed points");
};

template <typename T, int P> struct V {
    T a[P];   // x[2j]
    T b[P];   // x[2j+1]
};

template <typename T, int LOGM, int LOGL> struct Tw {
    using G = Geo<LOGM, LOGL>;
    T shc[LOGL > 0 ? LOGL : 1], shs[LOGL > 0 ? LOGL : 1];   // shuffle stage i (h = 1<<i)
    T rgc[G::P], rgs[G::P];                                  // register stage hp: entries [hp-1+q], q<hp
    T nc[G::P], ns[G::P];                                    // exp(2 pi i e / n), e = p*L + gl
    __device__ void init(int gl) {
        double s, c;
#pragma unroll
        for (int i = 0; i < LOGL; ++i) {
            int h = 1 << i;
            sincospi(2.0 * (double)(gl & (h - 1)) / (double)(2 * h), &s, &c);
            shc[i] = (T)c; shs[i] = (T)s;
        }
#pragma unroll
        for (int hp = 1; hp < G::P; hp <<= 1) {
#pragma unroll
            for (int q = 0; q < hp; ++q) {
                sincospi(2.0 * (double)(q * G::L + gl) / (double)(2 * hp * G::L), &s, &c);
                rgc[hp - 1 + q] = (T)c; rgs[hp - 1 + q] = (T)s;
            }
        }
#pragma unroll
        for (int p = 0; p < G::P; ++p) {
            sincospi(2.0 * (double)(p * G::L + gl) / (double)G::N, &s, &c);
            nc[p] = (T)c; ns[p] = (T)s;
        }
    }
};

// group-wide sums (lanes of one query)
template <typename T, int LOGL>
__device__ __forceinline__ T gsum(T v) {
#pragma unroll
    for (int o = (1 << LOGL) >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(CHK_FULL, v, o);
    return v;
}
template <typename T, int LOGL>
__device__ __forceinline__ void gsum3(T& a, T& b, T& c) {
#pragma unroll
    for (int o = (1 << LOGL) >> 1; o > 0; o >>= 1) {
        T ta = __shfl_xor_sync(CHK_FULL, a, o), tb = __shfl_xor_sync(CHK_FULL, b, o), tc = __shfl_xor_sync(CHK_FULL, c, o);
        a += ta; b += tb; c += tc;
    }
}
template <typename T, int P, int LOGL>
__device__ __forceinline__ T vdot(const V<T, P>& x, const V<T, P>& y) {
    T s = T(0);
#pragma unroll
    for (int p = 0; p < P; ++p) { s = Sc<T>::fma_(x.a[p], y.a[p], s); s = Sc<T>::fma_(x.b[p], y.b[p], s); }
    return gsum<T, LOGL>(s);
}

// ---- FFT cores ------------------------------------------------------------------------------
// inverse (e^{+}) DIF: natural order in, bit-reversed order out, unnormalised.
template <typename T, int LOGM, int LOGL>
__device__ __forceinline__ void fft_dif_inv(V<T, Geo<LOGM, LOGL>::P>& x, const Tw<T, LOGM, LOGL>& tw, int gl) {
    using G = Geo<LOGM, LOGL>;
#pragma unroll
    for (int hp = G::P >> 1; hp >= 1; hp >>= 1) {
#pragma unroll
        for (int p = 0; p < G::P; ++p) {
            if ((p & hp) == 0) {
                const int q = p & (hp - 1);
                const T c = tw.rgc[hp - 1 + q], s = tw.rgs[hp - 1 + q];
                T ar = x.a[p], ai = x.b[p], br = x.a[p + hp], bi = x.b[p + hp];
                x.a[p] = ar + br; x.b[p] = ai + bi;
                T dr = ar - br, di = ai - bi;
                x.a[p + hp] = dr * c - di * s;
                x.b[p + hp] = dr * s + di * c;
            }
        }
    }
#pragma unroll
    for (int i = LOGL - 1; i >= 0; --i) {
        const int h = 1 << i;
        const bool up = (gl & h) != 0;
        const T c = tw.shc[i], s = tw.shs[i];
#pragma unroll
        for (int p = 0; p < G::P; ++p) {
            T yr = __shfl_xor_sync(CHK_FULL, x.a[p], h), yi = __shfl_xor_sync(CHK_FULL, x.b[p], h);
            if (!up) { x.a[p] += yr; x.b[p] += yi; }
            else {
                T dr = yr - x.a[p], di = yi - x.b[p];
                x.a[p] = dr * c - di * s;
                x.b[p] = dr * s + di * c;
            }
        }
    }
}

// forward (e^{-}) DIT: bit-reversed order in, natural order out, unnormalised.
template <typename T, int LOGM, int LOGL>
__device__ __forceinline__ void fft_dit_fwd(V<T, Geo<LOGM, LOGL>::P>& x, const Tw<T, LOGM, LOGL>& tw, int gl) {
    using G = Geo<LOGM, LOGL>;
#pragma unroll
    for (int i = 0; i < LOGL; ++i) {
        const int h = 1 << i;
        const bool up = (gl & h) != 0;
        const T c = tw.shc[i], s = tw.shs[i];
#pragma unroll
        for (int p = 0; p < G::P; ++p) {
            T tr = x.a[p], ti = x.b[p];
            if (up) { tr = x.a[p] * c + x.b[p] * s; ti = x.b[p] * c - x.a[p] * s; }   // * conj(tw)
            T yr = __shfl_xor_sync(CHK_FULL, tr, h), yi = __shfl_xor_sync(CHK_FULL, ti, h);
            x.a[p] = up ? (yr - tr) : (tr + yr);
            x.b[p] = up ? (yi - ti) : (ti + yi);
        }
    }
#pragma unroll
    for (int hp = 1; hp < G::P; hp <<= 1) {
#pragma unroll
        for (int p = 0; p < G::P; ++p) {
            if ((p & hp) == 0) {
                const int q = p & (hp - 1);
                const T c = tw.rgc[hp - 1 + q], s = tw.rgs[hp - 1 + q];
                T br = x.a[p + hp] * c + x.b[p + hp] * s;
                T bi = x.b[p + hp] * c - x.a[p + hp] * s;
                T ar = x.a[p], ai = x.b[p];
                x.a[p] = ar + br; x.b[p] = ai + bi;
                x.a[p + hp] = ar - br; x.b[p + hp] = ai - bi;
            }
        }
    }
}

// C2R, ortho: spectrum row [Re X_0..X_M | Im X_0..X_M] in global memory -> packed real pairs in
// bit-reversed layout.  mid_scale multiplies X_1..X_{M-1} (1 for irfft, 1/2 for the adjoint of rfft).
template <typename T, int LOGM, int LOGL>
__device__ __forceinline__ void c2r_from_global(const T* __restrict__ row, T mid_scale,
                                                V<T, Geo<LOGM, LOGL>::P>& x, const Tw<T, LOGM, LOGL>& tw, int gl) {
    using G = Geo<LOGM, LOGL>;
    const T inv_sqrt_n = T(1) / Sc<T>::sqrt_((T)G::N);
#pragma unroll
    for (int p = 0; p < G::P; ++p) {
        const int k = p * G::L + gl;
        T ar = row[k], ai = row[G::R + k];
        T br = row[G::M - k], bi = -row[G::R + G::M - k];       // conj(X_{M-k})
        T sc = mid_scale * inv_sqrt_n;
        if (k == 0) { ai = T(0); bi = T(0); sc = inv_sqrt_n; }
        T sr = ar + br, si = ai + bi, dr = ar - br, di = ai - bi;
        T c = tw.nc[p], s = tw.ns[p];
        x.a[p] = (sr - (dr * s + di * c)) * sc;
        x.b[p] = (si + (dr * c - di * s)) * sc;
    }
    fft_dif_inv<T, LOGM, LOGL>(x, tw, gl);
}

// R2C, ortho: packed real pairs (bit-reversed layout) -> spectrum row in global memory.
// mid_scale multiplies X_1..X_{M-1} (1 for rfft, 2 for the adjoint of irfft).
template <typename T, int LOGM, int LOGL>
__device__ __forceinline__ void r2c_to_global(V<T, Geo<LOGM, LOGL>::P>& x, T mid_scale, T* __restrict__ row,
                                              const Tw<T, LOGM, LOGL>& tw, int gl, bool do_store) {
    using G = Geo<LOGM, LOGL>;
    fft_dit_fwd<T, LOGM, LOGL>(x, tw, gl);
    const T inv_sqrt_n = T(1) / Sc<T>::sqrt_((T)G::N);
    const int src = (G::L - gl) & (G::L - 1);
    T outr[G::P], outi[G::P];
#pragma unroll
    for (int p = 0; p < G::P; ++p) {
        // partner position (M - e) mod M: lane (L-gl)&(L-1), register P-1-p (gl>0) or (P-p)%P (gl==0)
        T pr = __shfl_sync(CHK_FULL, x.a[G::P - 1 - p], src, G::L);
        T pi = __shfl_sync(CHK_FULL, x.b[G::P - 1 - p], src, G::L);
        if (gl == 0) { pr = x.a[(G::P - p) % G::P]; pi = x.b[(G::P - p) % G::P]; }
        T zr = x.a[p], zi = x.b[p];
        T er = zr + pr, ei = zi - pi, dr = zr - pr, di = zi + pi;
        T c = tw.nc[p], s = tw.ns[p];
        const int k = p * G::L + gl;
        T scl = (k == 0 ? T(1) : mid_scale) * T(0.5) * inv_sqrt_n;
        outr[p] = (er + (di * c - dr * s)) * scl;
        outi[p] = (ei - (dr * c + di * s)) * scl;
    }
    if (do_store) {
#pragma unroll
        for (int p = 0; p < G::P; ++p) {
            const int k = p * G::L + gl;
            row[k] = outr[p];
            row[G::R + k] = (k == 0) ? T(0) : outi[p];
        }
        if (gl == 0) {
            row[G::M] = (x.a[0] - x.b[0]) * inv_sqrt_n;
            row[G::R + G::M] = T(0);
        }
    }
}

template <int LOGM>
__device__ __forceinline__ int bitrev(int e) { return (int)(__brev((unsigned)e) >> (32 - LOGM)); }

// row pairs in the bit-reversed layout
template <typename T, int LOGM, int LOGL>
__device__ __forceinline__ void load_pairs(const T* __restrict__ row, V<T, Geo<LOGM, LOGL>::P>& x, int gl) {
    using G = Geo<LOGM, LOGL>;
#pragma unroll
    for (int p = 0; p < G::P; ++p) {
        const int j = bitrev<LOGM>(p * G::L + gl);
        x.a[p] = row[2 * j]; x.b[p] = row[2 * j + 1];
    }
}
template <typename T, int LOGM, int LOGL>
__device__ __forceinline__ void store_pairs(T* __restrict__ row, const V<T, Geo<LOGM, LOGL>::P>& x, int gl, bool ok) {
    using G = Geo<LOGM, LOGL>;
    if (!ok) return;
#pragma unroll
    for (int p = 0; p < G::P; ++p) {
        const int j = bitrev<LOGM>(p * G::L + gl);
        row[2 * j] = x.a[p]; row[2 * j + 1] = x.b[p];
    }
}

// ---- hyperbolic ops (forward) ----------------------------------------------------------------
// project, utils/complexhyperbolic.py:72-87.  Returns through refs what the adjoint needs.
template <typename T, int P, int LOGL>
__device__ __forceinline__ void project_(V<T, P>& x, T sc) {
    T nrm = Sc<T>::max_(Sc<T>::sqrt_(vdot<T, P, LOGL>(x, x)), Sc<T>::min_norm);
    T maxnorm = Sc<T>::proj_top / sc;
    if (nrm > maxnorm) {
        T f = maxnorm / nrm;
#pragma unroll
        for (int p = 0; p < P; ++p) { x.a[p] *= f; x.b[p] *= f; }
    }
}
// expmap0, :41-54 (tanh clamp :36-37)
template <typename T, int P, int LOGL>
__device__ __forceinline__ void expmap0_(V<T, P>& u, T sc) {
    T nu = Sc<T>::max_(Sc<T>::sqrt_(vdot<T, P, LOGL>(u, u)), Sc<T>::min_norm);
    T a = sc * nu;
    T th = Sc<T>::tanh_(Sc<T>::min_(Sc<T>::max_(a, T(-15)), T(15)));
    T f = th / a;
#pragma unroll
    for (int p = 0; p < P; ++p) { u.a[p] *= f; u.b[p] *= f; }
    project_<T, P, LOGL>(u, sc);
}
// real_mobius_add, :90-106: out = x (+) y
template <typename T, int P, int LOGL>
__device__ __forceinline__ void mobius_(const V<T, P>& x, const V<T, P>& y, T c, V<T, P>& o) {
    T x2 = T(0), y2 = T(0), xy = T(0);
#pragma unroll
    for (int p = 0; p < P; ++p) {
        x2 = Sc<T>::fma_(x.a[p], x.a[p], x2); x2 = Sc<T>::fma_(x.b[p], x.b[p], x2);
        y2 = Sc<T>::fma_(y.a[p], y.a[p], y2); y2 = Sc<T>::fma_(y.b[p], y.b[p], y2);
        xy = Sc<T>::fma_(x.a[p], y.a[p], xy); xy = Sc<T>::fma_(x.b[p], y.b[p], xy);
    }
    gsum3<T, LOGL>(x2, y2, xy);
    T A = T(1) + T(2) * c * xy + c * y2;
    T B = T(1) - c * x2;
    T den = Sc<T>::max_(T(1) + T(2) * c * xy + c * c * x2 * y2, Sc<T>::min_norm);
    T ia = A / den, ib = B / den;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        o.a[p] = ia * x.a[p] + ib * y.a[p];
        o.b[p] = ia * x.b[p] + ib * y.b[p];
    }
}
// givens_rotations scale=None (utils/euclidean.py:39-42,55-57): o = ghat * x as complex numbers
template <typename T, int P>
__device__ __forceinline__ void rot_(const V<T, P>& g, const V<T, P>& x, V<T, P>& o) {
#pragma unroll
    for (int p = 0; p < P; ++p) {
        T inv = T(1) / Sc<T>::sqrt_(g.a[p] * g.a[p] + g.b[p] * g.b[p]);
        T g0 = g.a[p] * inv, g1 = g.b[p] * inv;
        T x0 = x.a[p], x1 = x.b[p];
        o.a[p] = g0 * x0 - g1 * x1;
        o.b[p] = g0 * x1 + g1 * x0;
    }
}
// givens_reflection AS CODED (utils/euclidean.py:60-75): o0 = g0 x0 + g1 x1, o1 = (g1 - g0) x0
template <typename T, int P>
__device__ __forceinline__ void refl_(const V<T, P>& g, const V<T, P>& x, V<T, P>& o) {
#pragma unroll
    for (int p = 0; p < P; ++p) {
        T inv = T(1) / Sc<T>::sqrt_(g.a[p] * g.a[p] + g.b[p] * g.b[p]);
        T g0 = g.a[p] * inv, g1 = g.b[p] * inv;
        T x0 = x.a[p], x1 = x.b[p];
        o.a[p] = g0 * x0 + g1 * x1;
        o.b[p] = g0 * (-x0) + g1 * x0;
    }
}

// ---- adjoints ---------------------------------------------------------------------------------
// project adjoint: g (in: grad of output, out: grad of input); returns d/dc contribution.
template <typename T, int P, int LOGL>
__device__ __forceinline__ T project_bwd_(const V<T, P>& x, T c, T sc, V<T, P>& g) {
    T raw = Sc<T>::sqrt_(vdot<T, P, LOGL>(x, x));
    T nrm = Sc<T>::max_(raw, Sc<T>::min_norm);
    T maxnorm = Sc<T>::proj_top / sc;
    if (!(nrm > maxnorm)) return T(0);
    T gx = vdot<T, P, LOGL>(g, x);
    T f = maxnorm / nrm;
    T k = (raw >= Sc<T>::min_norm) ? gx * maxnorm / (nrm * nrm * nrm) : T(0);
#pragma unroll
    for (int p = 0; p < P; ++p) {
        g.a[p] = g.a[p] * f - k * x.a[p];
        g.b[p] = g.b[p] * f - k * x.b[p];
    }
    return -(gx / nrm) * maxnorm / (T(2) * c);
}
// expmap0 adjoint at input u: g (grad of expmap0 output) -> grad of u; returns d/dc contribution.
template <typename T, int P, int LOGL>
__device__ __forceinline__ T expmap0_bwd_(const V<T, P>& u, T c, T sc, V<T, P>& g) {
    T raw = Sc<T>::sqrt_(vdot<T, P, LOGL>(u, u));
    T nu = Sc<T>::max_(raw, Sc<T>::min_norm);
    T a = sc * nu;
    bool inside = (a >= T(-15)) && (a <= T(15));
    T th = Sc<T>::tanh_(Sc<T>::min_(Sc<T>::max_(a, T(-15)), T(15)));
    T f = th / a;
    V<T, P> gam;
#pragma unroll
    for (int p = 0; p < P; ++p) { gam.a[p] = u.a[p] * f; gam.b[p] = u.b[p] * f; }
    T gc = project_bwd_<T, P, LOGL>(gam, c, sc, g);       // g is now grad of gamma
    T fprime = (inside ? (T(1) - th * th) / a : T(0)) - f / a;
    T gu = vdot<T, P, LOGL>(g, u);
    T k = (raw >= Sc<T>::min_norm) ? gu * fprime * sc / nu : T(0);
#pragma unroll
    for (int p = 0; p < P; ++p) {
        g.a[p] = f * g.a[p] + k * u.a[p];
        g.b[p] = f * g.b[p] + k * u.b[p];
    }
    return gc + gu * fprime * nu / (T(2) * sc);
}
// mobius adjoint: g = grad of out; writes gx, gy; returns d/dc.
template <typename T, int P, int LOGL>
__device__ __forceinline__ T mobius_bwd_(const V<T, P>& x, const V<T, P>& y, T c, const V<T, P>& g,
                                         V<T, P>& gx, V<T, P>& gy) {
    T x2 = T(0), y2 = T(0), xy = T(0), gdx = T(0), gdy = T(0), dummy = T(0);
#pragma unroll
    for (int p = 0; p < P; ++p) {
        x2 = Sc<T>::fma_(x.a[p], x.a[p], x2); x2 = Sc<T>::fma_(x.b[p], x.b[p], x2);
        y2 = Sc<T>::fma_(y.a[p], y.a[p], y2); y2 = Sc<T>::fma_(y.b[p], y.b[p], y2);
        xy = Sc<T>::fma_(x.a[p], y.a[p], xy); xy = Sc<T>::fma_(x.b[p], y.b[p], xy);
        gdx = Sc<T>::fma_(g.a[p], x.a[p], gdx); gdx = Sc<T>::fma_(g.b[p], x.b[p], gdx);
        gdy = Sc<T>::fma_(g.a[p], y.a[p], gdy); gdy = Sc<T>::fma_(g.b[p], y.b[p], gdy);
    }
    gsum3<T, LOGL>(x2, y2, xy);
    gsum3<T, LOGL>(gdx, gdy, dummy);
    T A = T(1) + T(2) * c * xy + c * y2;
    T B = T(1) - c * x2;
    T den_raw = T(1) + T(2) * c * xy + c * c * x2 * y2;
    T den = Sc<T>::max_(den_raw, Sc<T>::min_norm);
    T gA = gdx / den, gB = gdy / den;                       // g_num . x, g_num . y
    T gnum_dot_num = A * gA + B * gB;                        // (g/den) . num
    T gden = (den_raw >= Sc<T>::min_norm) ? -gnum_dot_num / den : T(0);
    T gxy = T(2) * c * (gA + gden);
    T gx2 = -c * gB + gden * c * c * y2;
    T gy2 = c * gA + gden * c * c * x2;
    T ia = A / den, ib = B / den;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        T ga = g.a[p], gb = g.b[p];
        gx.a[p] = ia * ga + gxy * y.a[p] + T(2) * gx2 * x.a[p];
        gx.b[p] = ia * gb + gxy * y.b[p] + T(2) * gx2 * x.b[p];
        gy.a[p] = ib * ga + gxy * x.a[p] + T(2) * gy2 * y.a[p];
        gy.b[p] = ib * gb + gxy * x.b[p] + T(2) * gy2 * y.b[p];
    }
    return gA * (T(2) * xy + y2) - gB * x2 + gden * (T(2) * xy + T(2) * c * x2 * y2);
}
// through g/|g|: given raw pair (r0,r1) and grad wrt ghat
template <typename T>
__device__ __forceinline__ void norm_bwd_(T r0, T r1, T gh0, T gh1, T& o0, T& o1) {
    T inv = T(1) / Sc<T>::sqrt_(r0 * r0 + r1 * r1);
    T h0 = r0 * inv, h1 = r1 * inv;
    T d = gh0 * h0 + gh1 * h1;
    o0 = (gh0 - d * h0) * inv;
    o1 = (gh1 - d * h1) * inv;
}
template <typename T, int P>
__device__ __forceinline__ void rot_bwd_(const V<T, P>& g, const V<T, P>& x, const V<T, P>& go, V<T, P>& gg, V<T, P>& gx) {
#pragma unroll
    for (int p = 0; p < P; ++p) {
        T inv = T(1) / Sc<T>::sqrt_(g.a[p] * g.a[p] + g.b[p] * g.b[p]);
        T g0 = g.a[p] * inv, g1 = g.b[p] * inv;
        T o0 = go.a[p], o1 = go.b[p], x0 = x.a[p], x1 = x.b[p];
        gx.a[p] = g0 * o0 + g1 * o1;
        gx.b[p] = -g1 * o0 + g0 * o1;
        norm_bwd_<T>(g.a[p], g.b[p], o0 * x0 + o1 * x1, -o0 * x1 + o1 * x0, gg.a[p], gg.b[p]);
    }
}
template <typename T, int P>
__device__ __forceinline__ void refl_bwd_(const V<T, P>& g, const V<T, P>& x, const V<T, P>& go, V<T, P>& gg, V<T, P>& gx) {
#pragma unroll
    for (int p = 0; p < P; ++p) {
        T inv = T(1) / Sc<T>::sqrt_(g.a[p] * g.a[p] + g.b[p] * g.b[p]);
        T g0 = g.a[p] * inv, g1 = g.b[p] * inv;
        T o0 = go.a[p], o1 = go.b[p], x0 = x.a[p], x1 = x.b[p];
        gx.a[p] = g0 * o0 + (g1 - g0) * o1;
        gx.b[p] = g1 * o0;
        norm_bwd_<T>(g.a[p], g.b[p], o0 * x0 - o1 * x0, o0 * x1 + o1 * x0, gg.a[p], gg.b[p]);
    }
}

template <typename T>
__device__ __forceinline__ T softplus_(T x) { return x > T(20) ? x : Sc<T>::log1p_(Sc<T>::exp_(x)); }
template <typename T>
__device__ __forceinline__ T sigmoid_(T x) { return T(1) / (T(1) + Sc<T>::exp_(-x)); }

template <typename T> struct QArgs {
    const T* entity; const T* rel; const T* rel_diag; const T* ctx; const T* c_table;
    const int64_t* head_idx; const int64_t* rel_idx;
    int64_t nq; int multi_c;
    // fwd outputs
    T* out_q; T* out_c;
    // bwd in/out
    const T* grad_q; T* g_entity_rows; T* g_rel_rows; T* g_rd_rows; T* g_ctx_rows; T* g_c;
};

// attention over [ref, rot] candidates, models/complexhyperbolic.py:150-158
template <typename T, int P, int LOGL>
__device__ __forceinline__ void att_fwd_(const V<T, P>& ctx, const V<T, P>& refq, const V<T, P>& rotq, T scale,
                                         T& w_ref, T& w_rot, V<T, P>& att) {
    T l_ref = T(0), l_rot = T(0), dummy = T(0);
#pragma unroll
    for (int p = 0; p < P; ++p) {
        l_ref += (ctx.a[p] * refq.a[p]) * scale + (ctx.b[p] * refq.b[p]) * scale;
        l_rot += (ctx.a[p] * rotq.a[p]) * scale + (ctx.b[p] * rotq.b[p]) * scale;
    }
    gsum3<T, LOGL>(l_ref, l_rot, dummy);
    T mx = Sc<T>::max_(l_ref, l_rot);
    T e_ref = Sc<T>::exp_(l_ref - mx), e_rot = Sc<T>::exp_(l_rot - mx);
    T inv = T(1) / (e_ref + e_rot);
    w_ref = e_ref * inv; w_rot = e_rot * inv;
#pragma unroll
    for (int p = 0; p < P; ++p) {
        att.a[p] = w_ref * refq.a[p] + w_rot * rotq.a[p];
        att.b[p] = w_ref * refq.b[p] + w_rot * rotq.b[p];
    }
}

template <typename T, int LOGM, int LOGL, int KIND, bool BWD>
__global__ void __launch_bounds__(128) query_kernel(QArgs<T> A) {
    using G = Geo<LOGM, LOGL>;
    constexpr int P = G::P;
    typedef V<T, P> Vt;
    const int lane = threadIdx.x & 31;
    const int gl = lane & (G::L - 1);
    const int grp = lane >> LOGL;
    Tw<T, LOGM, LOGL> tw;
    tw.init(gl);
    const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t warp_id = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t n_iter = (A.nq + G::QPW - 1) / G::QPW;
    constexpr int N = G::N, R2W = 2 * G::R;
    constexpr int RDW = (KIND == CHK_ATT) ? 2 * N : N;

    for (int64_t it = warp_id; it < n_iter; it += warps_total) {
        int64_t qi = it * G::QPW + grp;
        const bool ok = qi < A.nq;
        if (!ok) qi = A.nq - 1;                       // keep the group convergent; stores are masked
        const int64_t h = A.head_idx[qi], rl = A.rel_idx[qi];
        const T c_raw = A.multi_c ? A.c_table[rl] : A.c_table[0];
        const T c = A.multi_c ? softplus_<T>(c_raw) : c_raw;
        const T sc = Sc<T>::sqrt_(c);
        const T* relrow = A.rel + rl * (2 * N);
        const T* rdrow = A.rel_diag + rl * RDW;

        Vt u;
        c2r_from_global<T, LOGM, LOGL>(A.entity + h * R2W, T(1), u, tw, gl);
        Vt v;      // vector that feeds the final rfft
        // saved for the adjoint
        Vt hu, t1, t2, m1, lhs, rd, refq, rotq, cx, att;
        T w_ref = T(0), w_rot = T(0);
        const T att_scale = T(1) / Sc<T>::sqrt_((T)G::R);

        if (KIND == CHK_ROT) {
            hu = u; expmap0_<T, P, LOGL>(hu, sc);
            load_pairs<T, LOGM, LOGL>(relrow, t1, gl); expmap0_<T, P, LOGL>(t1, sc);
            load_pairs<T, LOGM, LOGL>(relrow + N, t2, gl); expmap0_<T, P, LOGL>(t2, sc);
            mobius_<T, P, LOGL>(hu, t1, c, m1);
            lhs = m1; project_<T, P, LOGL>(lhs, sc);
            load_pairs<T, LOGM, LOGL>(rdrow, rd, gl);
            Vt res1; rot_<T, P>(rd, lhs, res1);
            if (!BWD) { mobius_<T, P, LOGL>(res1, t2, c, v); }
            else {
                // ---- adjoint ----
                Vt g; c2r_from_global<T, LOGM, LOGL>(A.grad_q + qi * R2W, T(0.5), g, tw, gl);   // grad of v
                Vt g_res1, g_t2;
                T gc = mobius_bwd_<T, P, LOGL>(res1, t2, c, g, g_res1, g_t2);
                Vt r2; load_pairs<T, LOGM, LOGL>(relrow + N, r2, gl);
                gc += expmap0_bwd_<T, P, LOGL>(r2, c, sc, g_t2);
                store_pairs<T, LOGM, LOGL>(A.g_rel_rows + qi * (2 * N) + N, g_t2, gl, ok);
                Vt g_rd, g_lhs; rot_bwd_<T, P>(rd, lhs, g_res1, g_rd, g_lhs);
                store_pairs<T, LOGM, LOGL>(A.g_rd_rows + qi * RDW, g_rd, gl, ok);
                gc += project_bwd_<T, P, LOGL>(m1, c, sc, g_lhs);                  // g_lhs -> grad of m1
                Vt g_hu, g_t1;
                gc += mobius_bwd_<T, P, LOGL>(hu, t1, c, g_lhs, g_hu, g_t1);
                Vt r1; load_pairs<T, LOGM, LOGL>(relrow, r1, gl);
                gc += expmap0_bwd_<T, P, LOGL>(r1, c, sc, g_t1);
                store_pairs<T, LOGM, LOGL>(A.g_rel_rows + qi * (2 * N), g_t1, gl, ok);
                gc += expmap0_bwd_<T, P, LOGL>(u, c, sc, g_hu);                    // g_hu -> grad of u
                r2c_to_global<T, LOGM, LOGL>(g_hu, T(2), A.g_entity_rows + qi * R2W, tw, gl, ok);
                if (ok && gl == 0) A.g_c[qi] = A.multi_c ? gc * sigmoid_<T>(c_raw) : gc;
            }
        } else {
            // REF and ATT share the tail: v = project(mobius(expmap0(pre), expmap0(rel[:n])))
            Vt pre;
            if (KIND == CHK_REF) {
                load_pairs<T, LOGM, LOGL>(rdrow, rd, gl);
                refl_<T, P>(rd, u, pre);
            } else {
                load_pairs<T, LOGM, LOGL>(rdrow, rd, gl);           // rot half
                load_pairs<T, LOGM, LOGL>(rdrow + N, t2, gl);       // ref half (reuse t2 storage)
                rot_<T, P>(rd, u, rotq);
                refl_<T, P>(t2, u, refq);
                load_pairs<T, LOGM, LOGL>(A.ctx + rl * N, cx, gl);
                att_fwd_<T, P, LOGL>(cx, refq, rotq, att_scale, w_ref, w_rot, att);
                pre = att;
            }
            lhs = pre; expmap0_<T, P, LOGL>(lhs, sc);
            load_pairs<T, LOGM, LOGL>(relrow, t1, gl); expmap0_<T, P, LOGL>(t1, sc);
            mobius_<T, P, LOGL>(lhs, t1, c, m1);
            if (!BWD) { v = m1; project_<T, P, LOGL>(v, sc); }
            else {
                Vt g; c2r_from_global<T, LOGM, LOGL>(A.grad_q + qi * R2W, T(0.5), g, tw, gl);   // grad of v
                T gc = project_bwd_<T, P, LOGL>(m1, c, sc, g);                     // -> grad of m1
                Vt g_lhs, g_t1;
                gc += mobius_bwd_<T, P, LOGL>(lhs, t1, c, g, g_lhs, g_t1);
                Vt r1; load_pairs<T, LOGM, LOGL>(relrow, r1, gl);
                gc += expmap0_bwd_<T, P, LOGL>(r1, c, sc, g_t1);
                store_pairs<T, LOGM, LOGL>(A.g_rel_rows + qi * (2 * N), g_t1, gl, ok);
                Vt zero;
#pragma unroll
                for (int p = 0; p < P; ++p) { zero.a[p] = T(0); zero.b[p] = T(0); }
                store_pairs<T, LOGM, LOGL>(A.g_rel_rows + qi * (2 * N) + N, zero, gl, ok);
                gc += expmap0_bwd_<T, P, LOGL>(pre, c, sc, g_lhs);                 // -> grad of pre
                Vt g_u;
                if (KIND == CHK_REF) {
                    Vt g_rd; refl_bwd_<T, P>(rd, u, g_lhs, g_rd, g_u);
                    store_pairs<T, LOGM, LOGL>(A.g_rd_rows + qi * RDW, g_rd, gl, ok);
                } else {
                    T gw_ref = T(0), gw_rot = T(0), dummy = T(0);
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        gw_ref += g_lhs.a[p] * refq.a[p] + g_lhs.b[p] * refq.b[p];
                        gw_rot += g_lhs.a[p] * rotq.a[p] + g_lhs.b[p] * rotq.b[p];
                    }
                    gsum3<T, LOGL>(gw_ref, gw_rot, dummy);
                    T avg = w_ref * gw_ref + w_rot * gw_rot;
                    T gl_ref = w_ref * (gw_ref - avg), gl_rot = w_rot * (gw_rot - avg);
                    Vt g_refq, g_rotq, g_cx;
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        g_refq.a[p] = w_ref * g_lhs.a[p] + gl_ref * att_scale * cx.a[p];
                        g_refq.b[p] = w_ref * g_lhs.b[p] + gl_ref * att_scale * cx.b[p];
                        g_rotq.a[p] = w_rot * g_lhs.a[p] + gl_rot * att_scale * cx.a[p];
                        g_rotq.b[p] = w_rot * g_lhs.b[p] + gl_rot * att_scale * cx.b[p];
                        g_cx.a[p] = att_scale * (gl_ref * refq.a[p] + gl_rot * rotq.a[p]);
                        g_cx.b[p] = att_scale * (gl_ref * refq.b[p] + gl_rot * rotq.b[p]);
                    }
                    store_pairs<T, LOGM, LOGL>(A.g_ctx_rows + qi * N, g_cx, gl, ok);
                    Vt g_rd_rot, g_rd_ref, g_u2;
                    rot_bwd_<T, P>(rd, u, g_rotq, g_rd_rot, g_u);
                    refl_bwd_<T, P>(t2, u, g_refq, g_rd_ref, g_u2);
                    store_pairs<T, LOGM, LOGL>(A.g_rd_rows + qi * RDW, g_rd_rot, gl, ok);
                    store_pairs<T, LOGM, LOGL>(A.g_rd_rows + qi * RDW + N, g_rd_ref, gl, ok);
#pragma unroll
                    for (int p = 0; p < P; ++p) { g_u.a[p] += g_u2.a[p]; g_u.b[p] += g_u2.b[p]; }
                }
                r2c_to_global<T, LOGM, LOGL>(g_u, T(2), A.g_entity_rows + qi * R2W, tw, gl, ok);
                if (ok && gl == 0) A.g_c[qi] = A.multi_c ? gc * sigmoid_<T>(c_raw) : gc;
            }
        }
        if (!BWD) {
            r2c_to_global<T, LOGM, LOGL>(v, T(1), A.out_q + qi * R2W, tw, gl, ok);
            if (ok && gl == 0) A.out_c[qi] = c;
        }
    }
}

template <typename T, int LOGM, int LOGL, int KIND, bool BWD>
int launch_query(const QArgs<T>& A, cudaStream_t st) {
    using G = Geo<LOGM, LOGL>;
    const int warps_per_block = 4;
    int64_t n_iter = (A.nq + G::QPW - 1) / G::QPW;
    int64_t blocks = (n_iter + warps_per_block - 1) / warps_per_block;
    const int64_t cap = 148 * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    query_kernel<T, LOGM, LOGL, KIND, BWD><<<(unsigned)blocks, warps_per_block * 32, 0, st>>>(A);
    CHK_CUDA_LAUNCH_CHECK("query_kernel");
    return CHK_OK;
}

template <typename T, int KIND, bool BWD>
int dispatch_rank(int rank, const QArgs<T>& A, cudaStream_t st) {
    static const bool alt = [] { const char* e = getenv("CHK_K1_ALT"); return e && e[0] == '1'; }();
    if (alt) {      // experiment: fewer lanes per query (more points per lane, fewer shuffle stages)
        if (rank == 33) return launch_query<T, 5, 2, KIND, BWD>(A, st);
        if (rank == 65) return launch_query<T, 6, 3, KIND, BWD>(A, st);
    }
    switch (rank) {
        case 9: return launch_query<T, 3, 3, KIND, BWD>(A, st);      // n=16
        case 17: return launch_query<T, 4, 3, KIND, BWD>(A, st);     // n=32
        case 33: return launch_query<T, 5, 3, KIND, BWD>(A, st);     // n=64:  8 lanes x 4 points, 4 queries/warp
        case 65: return launch_query<T, 6, 4, KIND, BWD>(A, st);     // n=128: 16 lanes x 4 points
        case 129: return launch_query<T, 7, 4, KIND, BWD>(A, st);    // n=256: 16 lanes x 8 points
        case 257: return launch_query<T, 8, 5, KIND, BWD>(A, st);    // n=512: 32 lanes x 8 points
        default:
            chk_set_error("rank %d unsupported: 2(rank-1) must be a power of two in [16,512]", rank);
            return CHK_EUNSUPPORTED;
    }
}

template <typename T, bool BWD>
int dispatch_kind(int kind, int rank, const QArgs<T>& A, cudaStream_t st) {
    switch (kind) {
        case CHK_ROT: return dispatch_rank<T, CHK_ROT, BWD>(rank, A, st);
        case CHK_REF: return dispatch_rank<T, CHK_REF, BWD>(rank, A, st);
        case CHK_ATT:
            if (!A.ctx) { chk_set_error("FFTAttH needs context_vec"); return CHK_EINVAL; }
            return dispatch_rank<T, CHK_ATT, BWD>(rank, A, st);
        default: chk_set_error("unknown model kind %d", kind); return CHK_EINVAL;
    }
}

template <typename T>
QArgs<T> make_args(int64_t nq, int multi_c, const void* entity, const void* rel, const void* rel_diag, const void* ctx,
                   const void* c_table, const int64_t* head_idx, const int64_t* rel_idx) {
    QArgs<T> A{};
    A.entity = (const T*)entity; A.rel = (const T*)rel; A.rel_diag = (const T*)rel_diag; A.ctx = (const T*)ctx;
    A.c_table = (const T*)c_table; A.head_idx = head_idx; A.rel_idx = rel_idx; A.nq = nq; A.multi_c = multi_c;
    return A;
}

}  // namespace

int chk_query_fwd_tpq(int kind, int rank, int64_t nq, int multi_c, const void* entity, const void* rel, const void* rel_diag,
                      const void* ctx, const void* c_table, const int64_t* head_idx, const int64_t* rel_idx,
                      const int32_t* perm, void* out_q, void* out_c, cudaStream_t st);

extern "C" int chk_query_fwd(int kind, int dtype, int rank, int64_t nq, int multi_c,
                             const void* entity, const void* rel, const void* rel_diag, const void* ctx,
                             const void* c_table, const int64_t* head_idx, const int64_t* rel_idx,
                             void* out_q, void* out_c, void* stream) {
    if (nq == 0) return CHK_OK;
    if (nq < 0 || !entity || !rel || !rel_diag || !c_table || !head_idx || !rel_idx || !out_q || !out_c) {
        chk_set_error("chk_query_fwd: null pointer or negative size"); return CHK_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CHK_F32) {
        auto A = make_args<float>(nq, multi_c, entity, rel, rel_diag, ctx, c_table, head_idx, rel_idx);
        A.out_q = (float*)out_q; A.out_c = (float*)out_c;
        return dispatch_kind<float, false>(kind, rank, A, st);
    } else if (dtype == CHK_F64) {
        auto A = make_args<double>(nq, multi_c, entity, rel, rel_diag, ctx, c_table, head_idx, rel_idx);
        A.out_q = (double*)out_q; A.out_c = (double*)out_c;
        return dispatch_kind<double, false>(kind, rank, A, st);
    }
    chk_set_error("unknown dtype %d", dtype);
    return CHK_EINVAL;
}

int chk_group_by_key_impl(const int64_t* keys, int64_t n, int n_keys, int32_t* perm, int32_t* counts, cudaStream_t st);

extern "C" int chk_group_by_key(const int64_t* keys, int64_t n, int n_keys, int32_t* perm, int32_t* counts_scratch, void* stream) {
    if (n == 0) return CHK_OK;
    if (n < 0 || !keys || !perm || !counts_scratch) { chk_set_error("chk_group_by_key: bad argument"); return CHK_EINVAL; }
    return chk_group_by_key_impl(keys, n, n_keys, perm, counts_scratch, (cudaStream_t)stream);
}

extern "C" int chk_query_fwd_grouped(int kind, int dtype, int rank, int64_t nq, int multi_c,
                                     const void* entity, const void* rel, const void* rel_diag, const void* ctx,
                                     const void* c_table, const int64_t* head_idx, const int64_t* rel_idx, const int32_t* perm,
                                     void* out_q, void* out_c, void* stream) {
    if (nq == 0) return CHK_OK;
    if (nq < 0 || !entity || !rel || !rel_diag || !c_table || !head_idx || !rel_idx || !out_q || !out_c) {
        chk_set_error("chk_query_fwd_grouped: null pointer or negative size"); return CHK_EINVAL;
    }
    if (dtype != CHK_F32 || !(rank == 9 || rank == 17 || rank == 33)) {
        chk_set_error("chk_query_fwd_grouped: fp32 and rank in {9,17,33} only (use chk_query_fwd)"); return CHK_EUNSUPPORTED;
    }
    return chk_query_fwd_tpq(kind, rank, nq, multi_c, entity, rel, rel_diag, ctx, c_table, head_idx, rel_idx, perm, out_q, out_c,
                             (cudaStream_t)stream);
}

extern "C" int chk_query_bwd(int kind, int dtype, int rank, int64_t nq, int multi_c,
                             const void* entity, const void* rel, const void* rel_diag, const void* ctx,
                             const void* c_table, const int64_t* head_idx, const int64_t* rel_idx,
                             const void* grad_q, void* g_entity_rows, void* g_rel_rows, void* g_rel_diag_rows,
                             void* g_ctx_rows, void* g_c, void* stream) {
    if (nq == 0) return CHK_OK;
    if (nq < 0 || !entity || !rel || !rel_diag || !c_table || !head_idx || !rel_idx || !grad_q || !g_entity_rows ||
        !g_rel_rows || !g_rel_diag_rows || !g_c || (kind == CHK_ATT && !g_ctx_rows)) {
        chk_set_error("chk_query_bwd: null pointer or negative size"); return CHK_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CHK_F32) {
        auto A = make_args<float>(nq, multi_c, entity, rel, rel_diag, ctx, c_table, head_idx, rel_idx);
        A.grad_q = (const float*)grad_q; A.g_entity_rows = (float*)g_entity_rows; A.g_rel_rows = (float*)g_rel_rows;
        A.g_rd_rows = (float*)g_rel_diag_rows; A.g_ctx_rows = (float*)g_ctx_rows; A.g_c = (float*)g_c;
        return dispatch_kind<float, true>(kind, rank, A, st);
    } else if (dtype == CHK_F64) {
        auto A = make_args<double>(nq, multi_c, entity, rel, rel_diag, ctx, c_table, head_idx, rel_idx);
        A.grad_q = (const double*)grad_q; A.g_entity_rows = (double*)g_entity_rows; A.g_rel_rows = (double*)g_rel_rows;
        A.g_rd_rows = (double*)g_rel_diag_rows; A.g_ctx_rows = (double*)g_ctx_rows; A.g_c = (double*)g_c;
        return dispatch_kind<double, true>(kind, rank, A, st);
    }
    chk_set_error("unknown dtype %d", dtype);
    return CHK_EINVAL;
}
