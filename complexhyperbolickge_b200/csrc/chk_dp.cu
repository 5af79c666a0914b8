// Data-parallel dense-gradient step over NVLink peer memory (no reference counterpart — the reference is
// single-device, optimizers/kg_optimizer.py:51).
//
// What the host used to do with one NCCL all_reduce of the flat dense gradient followed by chk_dense_apply on every replica
// (every rank applying the whole update) is ONE kernel here, plus a small second one:
//
//   chk_dp_fused_apply   every rank owns a 1/world slice of the flat index space.  For its slice it sums the `world` gradient
//                        buffers in ascending rank order through their peer addresses (symmetric memory: the same layout on
//                        every rank; deterministic, the same bits on every replica), applies torch.optim.Adagrad / Adam to
//                        its LOCAL parameter / state values and stores the results into EVERY replica (peer stores):
//                        reduce-scatter + optimizer + all-gather without a gradient ever being written back.
//   chk_dp_wait_clear    waits until every peer has finished its slice (so this replica is complete and nobody reads this
//                        rank's gradients any more) and clears the local gradient buffer for the next step.
//
// Cross-GPU ordering uses monotonic 32-bit flags in a symmetric signal array (st.release.sys / ld.acquire.sys): slot
// [r] "rank r's gradients of step E are complete", slot [world + r] "rank r has stored its slice of step E everywhere".  No
// block ever waits for another block of its own grid (the last block to finish sends the second flag), so the kernels cannot
// deadlock on residency; a peer that never arrives trips a ~4 s timeout that sets a status word instead of hanging the GPU.
#include "chk_common.cuh"

namespace {

__device__ __forceinline__ void st_release_sys(int* p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

constexpr int DP_MAX_WORLD = 8;                        // one NVSwitch box
struct DpArgs {
    const void* const* grad; void* const* param; void* const* s0; void* const* s1;   // device arrays of `world` peer base pointers
    int* const* sig;                                                                  // peer signal arrays, int[2 * world] each
    int world, rank; int64_t n;
    const double* hyper; const int* step_id;
    int* local;                                                                       // [0] epoch  [1] ticket A  [2] status  [3] ticket C  ([4..7]: chk_dp_all_gather)
};

// threads t < world of the block poll slot `base + t` of this rank's signal array until it reaches v
__device__ __forceinline__ void wait_all(const DpArgs& A, int base, int v) {
    if ((int)threadIdx.x < A.world) {
        const int* s = A.sig[A.rank] + base + threadIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(s) < v) {
            if (clock64() - t0 > 8000000000LL) { A.local[2] = 1; break; }            // a peer never arrived: flag it, do not hang
        }
    }
    __syncthreads();
}

template <typename T>
__device__ __forceinline__ void adagrad_elem(T& p, T g, T& a, T lr, T eps) {          // same arithmetic as chk_step.cu's adagrad_apply
    a = Sc<T>::fma_(g, g, a);
    p -= lr * g / (Sc<T>::sqrt_(a) + eps);
}
template <>
__device__ __forceinline__ void adagrad_elem<float>(float& p, float g, float& a, float lr, float eps) {
    a = __fmaf_rn(g, g, a);
    float sq, rc;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(sq) : "f"(a));
    asm("rcp.approx.f32 %0, %1;" : "=f"(rc) : "f"(sq + eps));
    p = __fmaf_rn(-(lr * g), rc, p);
}

// four consecutive elements, 16-byte aligned, through L2 (never this SM's L1: the data is written by other GPUs / kernels)
template <typename T> __device__ __forceinline__ void load4(const T* p, T (&v)[4]);
template <> __device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
    const float4 t = __ldcg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void load4<double>(const double* p, double (&v)[4]) {
    const double2 a = __ldcg(reinterpret_cast<const double2*>(p)), b = __ldcg(reinterpret_cast<const double2*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <typename T> __device__ __forceinline__ void store4(T* p, const T (&v)[4]);
template <> __device__ __forceinline__ void store4<float>(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void store4<double>(double* p, const double (&v)[4]) {
    reinterpret_cast<double2*>(p)[0] = make_double2(v[0], v[1]); reinterpret_cast<double2*>(p)[1] = make_double2(v[2], v[3]);
}

template <typename T, int OPT>
__global__ void __launch_bounds__(256) dp_fused_apply_kernel(DpArgs A) {
    const int E = *reinterpret_cast<volatile int*>(A.local);
    const int v = E + 1;
    if (blockIdx.x == 0 && (int)threadIdx.x < A.world) st_release_sys(A.sig[threadIdx.x] + A.rank, v);   // the kernel that produced my gradients is done
    wait_all(A, 0, v);
    const T lr = (T)A.hyper[0], eps = (T)A.hyper[1];
    T b2 = T(0), w1 = T(0), w2 = T(0), step_size = T(0), bc2s = T(1);
    if (OPT == CHK_OPT_ADAM) {
        const double beta1 = A.hyper[4], beta2 = A.hyper[5];
        const double t = (double)(*A.step_id);
        b2 = (T)beta2; w1 = (T)(1.0 - beta1); w2 = (T)(1.0 - beta2);
        step_size = (T)(A.hyper[0] / (1.0 - pow(beta1, t)));
        bc2s = (T)sqrt(1.0 - pow(beta2, t));
    }
    // n is a multiple of 4 * world (the host pads the flat buffers): every rank's slice is a whole number of 4-element vectors,
    // one per thread and pass, so a thread has `world` 16-byte (fp64: 32-byte) peer loads in flight and the whole slice is
    // covered in one or two passes of the grid (r2: one element per thread and pass left 14 serial NVLink round trips per thread)
    const int W = A.world;
    const int64_t per = A.n / W;
    const int64_t lo = (int64_t)A.rank * per, nvec = per / 4;
    T* const p_loc = (T*)A.param[A.rank];
    T* const s0_loc = (T*)A.s0[A.rank];
    T* const s1_loc = OPT == CHK_OPT_ADAM ? (T*)A.s1[A.rank] : nullptr;
    for (int64_t vi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; vi < nvec; vi += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = lo + 4 * vi;
        T gk[DP_MAX_WORLD][4];
#pragma unroll
        for (int k = 0; k < DP_MAX_WORLD; ++k) {
            if (k < W) load4<T>((const T*)A.grad[k] + i, gk[k]);
            else { gk[k][0] = gk[k][1] = gk[k][2] = gk[k][3] = T(0); }
        }
        T pv[4], a0[4], a1[4];
        load4<T>(p_loc + i, pv);
        load4<T>(s0_loc + i, a0);
        if (OPT == CHK_OPT_ADAM) load4<T>(s1_loc + i, a1);
        bool any = OPT == CHK_OPT_ADAM;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            T g = gk[0][e];
#pragma unroll
            for (int k = 1; k < DP_MAX_WORLD; ++k) if (k < W) g += gk[k][e];                         // ascending rank order
            if (OPT == CHK_OPT_ADAGRAD) {
                if (g != T(0)) { adagrad_elem<T>(pv[e], g, a0[e], lr, eps); any = true; }
            } else {
                T m = a0[e], vv = a1[e];
                m = m + w1 * (g - m);
                vv = Sc<T>::fma_(w2 * g, g, vv * b2);
                const T denom = Sc<T>::sqrt_(vv) / bc2s + eps;
                pv[e] = pv[e] - step_size * (m / denom);
                a0[e] = m; a1[e] = vv;
            }
        }
        if (any) {                                                     // untouched vectors (Adagrad, zero gradient) are already identical everywhere
#pragma unroll
            for (int k = 0; k < DP_MAX_WORLD; ++k) {
                if (k < W) {
                    store4<T>((T*)A.param[k] + i, pv);
                    store4<T>((T*)A.s0[k] + i, a0);
                    if (OPT == CHK_OPT_ADAM) store4<T>((T*)A.s1[k] + i, a1);
                }
            }
        }
    }
    __threadfence_system();                                            // my peer stores are visible system-wide before the flag
    __syncthreads();
    if (threadIdx.x == 0) {
        const bool last = atomicAdd(A.local + 1, 1) == (int)gridDim.x - 1;           // the block that finishes last speaks for the rank
        if (last) {
            A.local[1] = 0;
            __threadfence_system();
            for (int k = 0; k < W; ++k) st_release_sys(A.sig[k] + W + A.rank, v);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) dp_wait_clear_kernel(DpArgs A) {
    const int E = *reinterpret_cast<volatile int*>(A.local);
    wait_all(A, A.world, E + 1);                                        // every slice of this replica has been written; nobody reads my gradients
    T* g = (T*)const_cast<void*>(A.grad[A.rank]);
    const T z[4] = {T(0), T(0), T(0), T(0)};
    for (int64_t vi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; vi < A.n / 4; vi += (int64_t)gridDim.x * blockDim.x) store4<T>(g + 4 * vi, z);
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(A.local + 3, 1) == (int)gridDim.x - 1) { A.local[3] = 0; A.local[0] = E + 1; }
}

// all_gather by peer reads: after a flag barrier ("my block is complete") every rank copies the `world` blocks from their owners'
// symmetric buffers into its local [world, bytes] buffer.  The barrier also orders everything before it on every rank against
// everything after it on every other rank (the owner-sharded tables rely on that, parallel.py).
struct AgArgs {
    const void* const* src; void* dst; int64_t bytes; int* const* sig; int slot; int world, rank; int* local; int epoch_idx, ticket_idx;
};
template <typename V>
__global__ void __launch_bounds__(256) dp_all_gather_kernel(AgArgs A) {
    const int E = *reinterpret_cast<volatile int*>(A.local + A.epoch_idx);
    const int v = E + 1;
    if (blockIdx.x == 0 && (int)threadIdx.x < A.world) st_release_sys(A.sig[threadIdx.x] + A.slot + A.rank, v);
    if ((int)threadIdx.x < A.world) {
        const int* s = A.sig[A.rank] + A.slot + threadIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(s) < v) {
            if (clock64() - t0 > 8000000000LL) { A.local[2] = 1; break; }
        }
    }
    __syncthreads();
    const int64_t nv = A.bytes / (int64_t)sizeof(V), total = nv * A.world;
    V* dst = reinterpret_cast<V*>(A.dst);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(i / nv);
        dst[i] = __ldcg(reinterpret_cast<const V*>(A.src[k]) + (i - (int64_t)k * nv));
    }
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(A.local + A.ticket_idx, 1) == (int)gridDim.x - 1) { A.local[A.ticket_idx] = 0; A.local[A.epoch_idx] = v; }
}

}  // namespace

extern "C" int chk_dp_all_gather(int world, int rank, const void* const* peer_src, int64_t bytes_per_rank, void* dst,
                                 int32_t* const* peer_signal, int channel, int32_t* local_state, void* stream) {
    if (bytes_per_rank == 0) return CHK_OK;
    if (world < 2 || world > DP_MAX_WORLD || rank < 0 || rank >= world || bytes_per_rank < 0 || (bytes_per_rank & 7) || !peer_src || !dst ||
        !peer_signal || !local_state || channel < 0 || channel > 1) {
        chk_set_error("chk_dp_all_gather: bad argument (bytes_per_rank must be a multiple of 8, channel 0 or 1)"); return CHK_EINVAL;
    }
    AgArgs A{peer_src, dst, bytes_per_rank, (int* const*)peer_signal, (2 + channel) * world, world, rank, (int*)local_state, 4 + 2 * channel, 5 + 2 * channel};
    const bool v16 = (bytes_per_rank & 15) == 0 && (((uintptr_t)dst) & 15) == 0;
    const int64_t nvec = bytes_per_rank / (v16 ? 16 : 8) * world;
    // at most two 256-thread blocks per SM: the blocks spin until every peer has arrived and must never fill an SM (a rank's
    // other peer-memory kernel may have to start beside them for the peers to get its flag)
    int grid = (int)((nvec + 255) / 256); if (grid > 148 * 2) grid = 148 * 2; if (grid < 1) grid = 1;
    if (v16) dp_all_gather_kernel<uint4><<<grid, 256, 0, (cudaStream_t)stream>>>(A);
    else dp_all_gather_kernel<uint2><<<grid, 256, 0, (cudaStream_t)stream>>>(A);
    CHK_CUDA_LAUNCH_CHECK("dp_all_gather_kernel");
    return CHK_OK;
}

extern "C" int chk_dp_fused_apply(int dtype, int opt, int world, int rank, const void* const* peer_grad, void* const* peer_param,
                                  void* const* peer_state0, void* const* peer_state1, int32_t* const* peer_signal, int64_t n,
                                  const double* hyper, const int32_t* step_id, int32_t* local_state, void* stream) {
    if (n == 0) return CHK_OK;
    if (world < 2 || world > DP_MAX_WORLD || rank < 0 || rank >= world || n < 0 || n % (4 * world) != 0 || !peer_grad || !peer_param || !peer_state0 || !peer_signal || !hyper ||
        !local_state || (opt != CHK_OPT_ADAGRAD && opt != CHK_OPT_ADAM) || (opt == CHK_OPT_ADAM && (!peer_state1 || !step_id))) {
        chk_set_error("chk_dp_fused_apply: bad argument (n must be a multiple of 4 * world)"); return CHK_EINVAL;
    }
    DpArgs A{peer_grad, peer_param, peer_state0, peer_state1, (int* const*)peer_signal, world, rank, n, hyper, (const int*)step_id, (int*)local_state};
    const int64_t nvec = n / world / 4;
    int grid = (int)((nvec + 255) / 256); if (grid > 148 * 4) grid = 148 * 4; if (grid < 1) grid = 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CHK_F32 && opt == CHK_OPT_ADAGRAD) dp_fused_apply_kernel<float, CHK_OPT_ADAGRAD><<<grid, 256, 0, st>>>(A);
    else if (dtype == CHK_F32) dp_fused_apply_kernel<float, CHK_OPT_ADAM><<<grid, 256, 0, st>>>(A);
    else if (dtype == CHK_F64 && opt == CHK_OPT_ADAGRAD) dp_fused_apply_kernel<double, CHK_OPT_ADAGRAD><<<grid, 256, 0, st>>>(A);
    else if (dtype == CHK_F64) dp_fused_apply_kernel<double, CHK_OPT_ADAM><<<grid, 256, 0, st>>>(A);
    else { chk_set_error("unknown dtype %d", dtype); return CHK_EINVAL; }
    CHK_CUDA_LAUNCH_CHECK("dp_fused_apply_kernel");
    int grid2 = (int)((n / 4 + 255) / 256); if (grid2 > 148 * 4) grid2 = 148 * 4; if (grid2 < 1) grid2 = 1;
    if (dtype == CHK_F32) dp_wait_clear_kernel<float><<<grid2, 256, 0, st>>>(A);
    else dp_wait_clear_kernel<double><<<grid2, 256, 0, st>>>(A);
    CHK_CUDA_LAUNCH_CHECK("dp_wait_clear_kernel");
    return CHK_OK;
}
