// Training-loop kernels next to the hot path: the negative-sampling loss with its
// gradient, and a row-sparse Adagrad step that is exactly the dense torch.optim.Adagrad update.
//
//   chk_nsloss          KGOptimizer.neg_sampling_loss (reference optimizers/kg_optimizer.py:115-122):
//                       loss = -mean(cat[logsigmoid(s[:,0]), logsigmoid(-s[:,1:])]) over B*(1+neg) terms, and d loss / d s.
//   chk_sparse_adagrad  torch.optim.Adagrad (lr_decay = 0, weight_decay = 0): sum += g*g; p -= lr * g / (sqrt(sum) + eps)
//                       applied only to the listed rows (a zero gradient row is a no-op in Adagrad, so sparse == dense,
//                       SURVEY §7D), each row once per step however often it is listed, and the gradient row is cleared
//                       so the dense .grad buffer is all-zero again without an N x 2r memset.
#include "chk_common.cuh"

namespace {

template <typename T>
__device__ __forceinline__ T logsigmoid_(T x) {           // min(x,0) - log1p(exp(-|x|)), as ATen
    return Sc<T>::min_(x, T(0)) - Sc<T>::log1p_(Sc<T>::exp_(-Sc<T>::abs_(x)));
}
template <typename T>
__device__ __forceinline__ T sigmoid_t(T x) { return T(1) / (T(1) + Sc<T>::exp_(-x)); }

template <typename T>
__global__ void __launch_bounds__(256) nsloss_kernel(const T* __restrict__ s, int64_t B, int64_t nt, T* __restrict__ loss,
                                                     T* __restrict__ grad) {
    __shared__ T red[8];
    const int64_t total = B * nt;
    const T inv = T(1) / (T)total;
    T acc = T(0);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const bool pos = (i % nt) == 0;
        const T x = pos ? s[i] : -s[i];                   // term = logsigmoid(x)
        acc -= logsigmoid_<T>(x);
        const T g = -sigmoid_t<T>(-x) * inv;              // d(-logsigmoid(x))/dx / total
        grad[i] = pos ? g : -g;
    }
    acc = warp_sum<T>(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        T v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : T(0);
        v = warp_sum<T>(v);
        if (threadIdx.x == 0) atomicAdd(loss, v * inv);
    }
}

// Adagrad on one row by one warp.  Even widths (every 2r-wide / n-wide table: rows are 8-byte aligned) go two elements
// per lane with all loads of an iteration pair issued before the arithmetic; odd widths (bh, bt, c) take the scalar loop.
template <typename T>
__device__ __forceinline__ void adagrad_elem(T& pv, T& gv, T& av, T lr, T eps) {
    av = Sc<T>::fma_(gv, gv, av);
    pv -= lr * gv / (Sc<T>::sqrt_(av) + eps);
    gv = T(0);
}
template <typename T> struct Vec2;
template <> struct Vec2<float> { using type = float2; };
template <> struct Vec2<double> { using type = double2; };
template <typename T>
__device__ __forceinline__ void adagrad_row(T* __restrict__ p, T* __restrict__ g, T* __restrict__ a, int64_t width, int lane, T lr, T eps) {
    using V = typename Vec2<T>::type;
    if ((width & 1) == 0) {
        V* p2 = reinterpret_cast<V*>(p); V* g2 = reinterpret_cast<V*>(g); V* a2 = reinterpret_cast<V*>(a);
        const int64_t w2 = width >> 1;
        int64_t c = lane;
        for (; c + 32 < w2; c += 64) {
            V g0 = g2[c], g1 = g2[c + 32], a0 = a2[c], a1 = a2[c + 32], p0 = p2[c], p1 = p2[c + 32];
            adagrad_elem<T>(p0.x, g0.x, a0.x, lr, eps); adagrad_elem<T>(p0.y, g0.y, a0.y, lr, eps);
            adagrad_elem<T>(p1.x, g1.x, a1.x, lr, eps); adagrad_elem<T>(p1.y, g1.y, a1.y, lr, eps);
            a2[c] = a0; a2[c + 32] = a1; p2[c] = p0; p2[c + 32] = p1; g2[c] = g0; g2[c + 32] = g1;
        }
        for (; c < w2; c += 32) {
            V g0 = g2[c], a0 = a2[c], p0 = p2[c];
            adagrad_elem<T>(p0.x, g0.x, a0.x, lr, eps); adagrad_elem<T>(p0.y, g0.y, a0.y, lr, eps);
            a2[c] = a0; p2[c] = p0; g2[c] = g0;
        }
    } else {
        for (int64_t c = lane; c < width; c += 32) {
            T gv = g[c], av = a[c], pv = p[c];
            adagrad_elem<T>(pv, gv, av, lr, eps);
            a[c] = av; p[c] = pv; g[c] = gv;
        }
    }
}

// one warp per list entry; stamp[row] == *step_id marks "already updated in this step"
template <typename T>
__global__ void __launch_bounds__(256) sparse_adagrad_kernel(T* __restrict__ param, T* __restrict__ grad, T* __restrict__ sum,
                                                             const int64_t* __restrict__ rows, int64_t m, int64_t width,
                                                             T lr, T eps, int* __restrict__ stamp, const int* __restrict__ step_id) {
    const int lane = threadIdx.x & 31;
    const int cur = *step_id;
    for (int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < m; t += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        const int64_t row = rows[t];
        int first = 0;
        if (lane == 0) first = atomicExch(stamp + row, cur) != cur;
        first = __shfl_sync(CHK_FULL, first, 0);
        if (!first) continue;
        T* p = param + row * width; T* g = grad + row * width; T* a = sum + row * width;
        adagrad_row<T>(p, g, a, width, lane, lr, eps);
    }
}

__global__ void bump_kernel(int* c) { *c += 1; }

// Data-parallel sparse exchange, send side: slot t carries the WHOLE accumulated gradient row rows[t] if it is the
// first slot of this step naming that row (claimed through stamp[row] = -step, a token the Adagrad kernel never
// writes), zeros otherwise; the claimed row of the dense gradient is cleared (the sum over ranks is scattered back).
template <typename T>
__global__ void __launch_bounds__(256) claim_gather_kernel(T* __restrict__ grad, const int64_t* __restrict__ rows, int64_t m,
                                                           int64_t width, int* __restrict__ stamp, const int* __restrict__ step_id,
                                                           T* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int token = -(*step_id);
    for (int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < m; t += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        const int64_t row = rows[t];
        int first = 0;
        if (lane == 0) first = atomicExch(stamp + row, token) != token;
        first = __shfl_sync(CHK_FULL, first, 0);
        T* g = grad + row * width; T* o = out + t * width;
        for (int64_t c = lane; c < width; c += 32) {
            T v = T(0);
            if (first) { v = g[c]; g[c] = T(0); }
            o[c] = v;
        }
    }
}

// ---- multi-table variants: one launch walks up to CHK_MAX_TABLES tables (blockIdx.y = table) ---------------------
struct TabList { chk_table_desc t[CHK_MAX_TABLES]; };

template <typename T>
__global__ void __launch_bounds__(256) multi_scatter_kernel(TabList L) {
    const chk_table_desc d = L.t[blockIdx.y];
    if (!d.src_rows) return;
    T* dense = (T*)d.grad; const T* src = (const T*)d.src_rows;
    if (d.width < 32) {                                  // narrow tables (bh, bt, c): one thread per element
        const int64_t total = d.m * d.width;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
            const int64_t rrow = i / d.width, c = i - rrow * d.width;
            atomicAdd(dense + d.rows[rrow] * d.width + c, src[i]);
        }
        return;
    }
    const int lane = threadIdx.x & 31;                   // wide tables: one warp per row, coalesced reductions
    for (int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < d.m; t += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        T* g = dense + d.rows[t] * d.width; const T* v = src + t * d.width;
        for (int64_t c = lane; c < d.width; c += 32) atomicAdd(g + c, v[c]);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) multi_adagrad_kernel(TabList L, T lr, T eps, const int* __restrict__ step_id) {
    const chk_table_desc d = L.t[blockIdx.y];
    const int lane = threadIdx.x & 31;
    const int cur = *step_id;
    T* param = (T*)d.param; T* grad = (T*)d.grad; T* sum = (T*)d.state_sum;
    for (int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < d.m; t += (int64_t)gridDim.x * (blockDim.x >> 5)) {
        const int64_t row = d.rows[t];
        int first = 0;
        if (lane == 0) first = atomicExch(d.stamp + row, cur) != cur;
        first = __shfl_sync(CHK_FULL, first, 0);
        if (!first) continue;
        T* p = param + row * d.width; T* g = grad + row * d.width; T* a = sum + row * d.width;
        adagrad_row<T>(p, g, a, d.width, lane, lr, eps);
    }
}

}  // namespace

extern "C" int chk_nsloss(int dtype, int64_t B, int64_t nt, const void* scores, void* loss_accum, void* grad_scores, void* stream) {
    if (B == 0 || nt == 0) return CHK_OK;
    if (B < 0 || nt < 0 || !scores || !loss_accum || !grad_scores) { chk_set_error("chk_nsloss: bad argument"); return CHK_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    int64_t blocks = (B * nt + 255) / 256;
    if (blocks > 148 * 4) blocks = 148 * 4;
    if (dtype == CHK_F32) nsloss_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((const float*)scores, B, nt, (float*)loss_accum, (float*)grad_scores);
    else if (dtype == CHK_F64) nsloss_kernel<double><<<(unsigned)blocks, 256, 0, st>>>((const double*)scores, B, nt, (double*)loss_accum, (double*)grad_scores);
    else { chk_set_error("unknown dtype %d", dtype); return CHK_EINVAL; }
    CHK_CUDA_LAUNCH_CHECK("nsloss_kernel");
    return CHK_OK;
}

extern "C" int chk_sparse_adagrad(int dtype, void* param, void* grad, void* state_sum, const int64_t* rows, int64_t m,
                                  int64_t width, double lr, double eps, int32_t* stamp, const int32_t* step_id, void* stream) {
    if (m == 0 || width == 0) return CHK_OK;
    if (m < 0 || width < 0 || !param || !grad || !state_sum || !rows || !stamp || !step_id) { chk_set_error("chk_sparse_adagrad: bad argument"); return CHK_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    int64_t blocks = (m + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (dtype == CHK_F32) sparse_adagrad_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((float*)param, (float*)grad, (float*)state_sum, rows, m, width, (float)lr, (float)eps, stamp, step_id);
    else if (dtype == CHK_F64) sparse_adagrad_kernel<double><<<(unsigned)blocks, 256, 0, st>>>((double*)param, (double*)grad, (double*)state_sum, rows, m, width, lr, eps, stamp, step_id);
    else { chk_set_error("unknown dtype %d", dtype); return CHK_EINVAL; }
    CHK_CUDA_LAUNCH_CHECK("sparse_adagrad_kernel");
    return CHK_OK;
}

extern "C" int chk_claim_gather_rows(int dtype, void* grad, const int64_t* rows, int64_t m, int64_t width, int32_t* stamp,
                                     const int32_t* step_id, void* out_rows, void* stream) {
    if (m == 0 || width == 0) return CHK_OK;
    if (m < 0 || width < 0 || !grad || !rows || !stamp || !step_id || !out_rows) { chk_set_error("chk_claim_gather_rows: bad argument"); return CHK_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    int64_t blocks = (m + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (dtype == CHK_F32) claim_gather_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((float*)grad, rows, m, width, stamp, step_id, (float*)out_rows);
    else if (dtype == CHK_F64) claim_gather_kernel<double><<<(unsigned)blocks, 256, 0, st>>>((double*)grad, rows, m, width, stamp, step_id, (double*)out_rows);
    else { chk_set_error("unknown dtype %d", dtype); return CHK_EINVAL; }
    CHK_CUDA_LAUNCH_CHECK("claim_gather_kernel");
    return CHK_OK;
}

extern "C" int chk_step_counter_bump(int32_t* counter, void* stream) {
    if (!counter) { chk_set_error("chk_step_counter_bump: null"); return CHK_EINVAL; }
    bump_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter);
    CHK_CUDA_LAUNCH_CHECK("bump_kernel");
    return CHK_OK;
}

static int check_tabs(const chk_table_desc* tabs, int n, bool need_src, bool need_opt, TabList& L, int64_t& max_items, int64_t per_block) {
    if (n < 1 || n > CHK_MAX_TABLES || !tabs) { chk_set_error("multi-table call: 1..%d tables expected", CHK_MAX_TABLES); return CHK_EINVAL; }
    max_items = 0;
    for (int i = 0; i < n; ++i) {
        const chk_table_desc& d = tabs[i];
        if (d.m < 0 || d.width <= 0 || !d.grad || (d.m > 0 && !d.rows) || (need_opt && (!d.param || !d.state_sum || !d.stamp))) {
            chk_set_error("multi-table call: bad descriptor %d", i); return CHK_EINVAL;
        }
        L.t[i] = d;
        const int64_t items = need_src ? (d.src_rows ? d.m * d.width : 0) : d.m;
        if (items > max_items) max_items = items;
    }
    (void)per_block;
    return CHK_OK;
}

extern "C" int chk_multi_scatter_add(int dtype, const chk_table_desc* tabs, int n_tables, void* stream) {
    TabList L{}; int64_t mx = 0;
    int rc = check_tabs(tabs, n_tables, true, false, L, mx, 256);
    if (rc != CHK_OK) return rc;
    if (mx == 0) return CHK_OK;
    int64_t bx = (mx + 255) / 256; if (bx > 148 * 8) bx = 148 * 8;
    dim3 grid((unsigned)bx, (unsigned)n_tables);
    if (dtype == CHK_F32) multi_scatter_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(L);
    else if (dtype == CHK_F64) multi_scatter_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>(L);
    else { chk_set_error("unknown dtype %d", dtype); return CHK_EINVAL; }
    CHK_CUDA_LAUNCH_CHECK("multi_scatter_kernel");
    return CHK_OK;
}

extern "C" int chk_multi_sparse_adagrad(int dtype, const chk_table_desc* tabs, int n_tables, double lr, double eps,
                                        const int32_t* step_id, void* stream) {
    TabList L{}; int64_t mx = 0;
    if (!step_id) { chk_set_error("chk_multi_sparse_adagrad: null step_id"); return CHK_EINVAL; }
    int rc = check_tabs(tabs, n_tables, false, true, L, mx, 8);
    if (rc != CHK_OK) return rc;
    if (mx == 0) return CHK_OK;
    int64_t bx = (mx + 7) / 8; if (bx > 148 * 8) bx = 148 * 8;
    dim3 grid((unsigned)bx, (unsigned)n_tables);
    if (dtype == CHK_F32) multi_adagrad_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(L, (float)lr, (float)eps, step_id);
    else if (dtype == CHK_F64) multi_adagrad_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>(L, lr, eps, step_id);
    else { chk_set_error("unknown dtype %d", dtype); return CHK_EINVAL; }
    CHK_CUDA_LAUNCH_CHECK("multi_adagrad_kernel");
    return CHK_OK;
}
