// Evaluation batch as ONE call: the body of KGModel.get_ranking's loop (reference models/base.py:243-271)
// for one batch, enqueued from C with the filter index resident on the device.
//
//   chk_filter_lookup   the reference's per-query `filters[(h, r)]` dictionary lookup + `+= [t]` (models/base.py:264-268,
//                       datasets/process.py:55-77) against a sorted key table + CSR that lives in HBM: binary search of the
//                       key, the list's range, membership of the true tail — no per-batch host work, no CSR upload.
//   chk_eval_batch      split ids -> K1 -> query norms -> target scores -> rank counts (exact or tcgen05 tier) -> filter pass.
//                       One host call per batch instead of ~20 (the per-batch host cost was what limited 8-GPU scaling and the
//                       end-to-end number of the small graphs in round 1).
#include "chk_common.cuh"

int chk_rank_counts_fma(int dtype, int rank, int64_t b, const void* q, const void* qn, const void* bh_vals,
                        const void* target, const void* entity, const void* hn, const void* bt,
                        int64_t n_rows, int64_t* counts, cudaStream_t st);
int chk_rank_counts_mma(int dtype, int rank, int64_t b, const void* q, const void* qn, const void* bh_vals, const void* target,
                        const void* entity, const void* hn, const void* bt, int64_t n_rows, const void* shadow, void* workspace,
                        int64_t workspace_bytes, int64_t* counts, cudaStream_t st);

namespace {

// ids of the batch, bh values, zeroed counters
template <typename T>
__global__ void __launch_bounds__(256) eval_split_kernel(const int64_t* __restrict__ queries, int64_t b, const T* __restrict__ bh,
                                                         int64_t* __restrict__ heads, int64_t* __restrict__ rels, int64_t* __restrict__ tails,
                                                         T* __restrict__ bhv, int64_t* __restrict__ counts) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < b; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t h = queries[3 * i];
        heads[i] = h; rels[i] = queries[3 * i + 1]; tails[i] = queries[3 * i + 2];
        if (bh) bhv[i] = bh[h];
        counts[i] = 0;
    }
}

// target[i] = score(q_i, entity[tails[i]]) with the canonical arithmetic, straight from the full table
template <typename T>
__global__ void __launch_bounds__(32) target_pairs_kernel(RArgs<T> A, const int64_t* __restrict__ tails, T* __restrict__ target) {
    __shared__ PairTiles<T> S;
    for (int64_t i0 = (int64_t)blockIdx.x * 32; i0 < A.b; i0 += (int64_t)gridDim.x * 32) {
        const int64_t i = i0 + threadIdx.x;
        const bool valid = i < A.b;
        const T s = warp_exact_pairs<T>(A, (unsigned)(valid ? i : 0), (unsigned)(valid ? tails[i] : 0), valid, S);
        if (valid) target[i] = s;
    }
}

// One CTA: per query the key's list range in the device-resident CSR, the running offsets of the batch's lists, and the
// true tail's own contribution (it always "outranks" itself: subtract 1 when it is not already in the list and lives in
// this shard — the list entries are handled by the filter pass).
__global__ void __launch_bounds__(1024) filter_lookup_kernel(const int64_t* __restrict__ queries, int64_t b, int64_t n_rel2,
                                                             const int64_t* __restrict__ keys, int64_t nk, const int64_t* __restrict__ indptr,
                                                             const int64_t* __restrict__ vals, int64_t shard_offset, int64_t n_rows,
                                                             int64_t* __restrict__ flt_indptr, int64_t* __restrict__ flt_start,
                                                             unsigned long long* __restrict__ counts, int* __restrict__ flags) {
    __shared__ int64_t wsum[32];
    __shared__ int64_t carry;
    if (threadIdx.x == 0) { carry = 0; flt_indptr[0] = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t i0 = 0; i0 < b; i0 += blockDim.x) {
        const int64_t i = i0 + threadIdx.x;
        int64_t len = 0;
        if (i < b) {
            const int64_t code = queries[3 * i] * n_rel2 + queries[3 * i + 1], t = queries[3 * i + 2];
            int64_t lo = 0, hi = nk;                                   // first key >= code
            while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (keys[mid] < code) lo = mid + 1; else hi = mid; }
            int64_t start = 0;
            bool t_in = false;
            if (lo < nk && keys[lo] == code) {
                start = indptr[lo]; len = indptr[lo + 1] - start;
                int64_t a = start, z = start + len;                    // the lists are sorted: is the true tail in it?
                while (a < z) { const int64_t mid = (a + z) >> 1; if (vals[mid] < t) a = mid + 1; else z = mid; }
                t_in = a < start + len && vals[a] == t;
            } else {
                atomicOr(flags, 1);                                    // the reference raises KeyError (models/base.py:266)
            }
            flt_start[i] = start;
            const int64_t tl = t - shard_offset;
            if (!t_in && tl >= 0 && tl < n_rows) atomicAdd(counts + i, (unsigned long long)(-1LL));
        }
        // block-wide inclusive scan of len
        int64_t x = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int64_t y = __shfl_up_sync(CHK_FULL, x, o); if (lane >= o) x += y; }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int64_t w = lane < (int)(blockDim.x >> 5) ? wsum[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int64_t y = __shfl_up_sync(CHK_FULL, w, o); if (lane >= o) w += y; }
            wsum[lane] = w;
        }
        __syncthreads();
        const int64_t before = carry + (warp ? wsum[warp - 1] : 0);
        if (i < b) flt_indptr[i + 1] = before + x;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = before + x;
        __syncthreads();
    }
}

// counts[i] -= #{ e in list_i within the shard : score(i,e) >= target[i] }; the lists are ranges of the device-resident CSR
template <typename T>
__global__ void __launch_bounds__(32) filter_sub_ranged_kernel(RArgs<T> A, const int64_t* __restrict__ flt_indptr,
                                                               const int64_t* __restrict__ flt_start, const int64_t* __restrict__ vals,
                                                               int64_t shard_offset) {
    __shared__ PairTiles<T> S;
    const int64_t total = flt_indptr[A.b];
    for (int64_t t0 = (int64_t)blockIdx.x * 32; t0 < total; t0 += (int64_t)gridDim.x * 32) {
        const int64_t t = t0 + threadIdx.x;
        bool valid = t < total;
        int64_t i = 0, e = 0;
        if (valid) {
            int64_t lo = 0, hi = A.b;                                  // largest i with flt_indptr[i] <= t
            while (hi - lo > 1) { const int64_t mid = (lo + hi) >> 1; if (flt_indptr[mid] <= t) lo = mid; else hi = mid; }
            i = lo;
            e = vals[flt_start[i] + (t - flt_indptr[i])] - shard_offset;
            valid = e >= 0 && e < A.n_rows;
        }
        if (!__any_sync(CHK_FULL, valid)) continue;
        const T s = warp_exact_pairs<T>(A, (unsigned)i, (unsigned)e, valid, S);
        if (valid && s >= A.target[i]) atomicAdd(A.counts + i, (unsigned long long)(-1LL));
    }
}

template <typename T>
RArgs<T> rargs(int rank, int64_t b, const void* q, const void* qn, const void* bhv, const void* target, const void* entity,
               const void* hn, const void* bt, int64_t n_rows) {
    RArgs<T> A{};
    A.q = (const T*)q; A.qn = (const T*)qn; A.bh_vals = (const T*)bhv; A.target = (const T*)target;
    A.entity = (const T*)entity; A.hn = (const T*)hn; A.bt = (const T*)bt; A.b = b; A.n_rows = n_rows; A.r = rank;
    return A;
}

struct Scratch { int64_t *heads, *rels, *tails, *flt_indptr, *flt_start; char *q, *c_out, *qn, *bhv; };

int64_t carve(char* base, int dtype, int rank, int64_t b, Scratch* s) {
    const int64_t es = dtype == CHK_F64 ? 8 : 4;
    int64_t off = 0;
    auto take = [&](int64_t bytes) { char* p = base ? base + off : nullptr; off += (bytes + 255) & ~int64_t(255); return p; };
    char* h = take(8 * b); char* r = take(8 * b); char* t = take(8 * b); char* fi = take(8 * (b + 1)); char* fs = take(8 * b);
    char* q = take(es * b * 2 * rank); char* c = take(es * b); char* qn = take(es * b); char* bhv = take(es * b);
    if (s) { s->heads = (int64_t*)h; s->rels = (int64_t*)r; s->tails = (int64_t*)t; s->flt_indptr = (int64_t*)fi; s->flt_start = (int64_t*)fs;
             s->q = q; s->c_out = c; s->qn = qn; s->bhv = bhv; }
    return off;
}

}  // namespace

extern "C" int chk_filter_lookup(const int64_t* queries, int64_t b, int64_t n_rel2, const int64_t* keys, int64_t n_keys,
                                 const int64_t* indptr, const int64_t* vals, int64_t shard_offset, int64_t n_rows,
                                 int64_t* flt_indptr, int64_t* flt_start, int64_t* counts, int32_t* flags, void* stream) {
    if (b == 0) return CHK_OK;
    if (b < 0 || n_keys < 0 || !queries || (n_keys > 0 && (!keys || !indptr)) || !flt_indptr || !flt_start || !counts || !flags) {
        chk_set_error("chk_filter_lookup: bad argument"); return CHK_EINVAL;
    }
    const int threads = b >= 1024 ? 1024 : (int)((b + 31) / 32 * 32);
    filter_lookup_kernel<<<1, threads, 0, (cudaStream_t)stream>>>(queries, b, n_rel2, keys, n_keys, indptr, vals, shard_offset, n_rows,
                                                                  flt_indptr, flt_start, (unsigned long long*)counts, flags);
    CHK_CUDA_LAUNCH_CHECK("filter_lookup_kernel");
    return CHK_OK;
}

extern "C" int64_t chk_eval_scratch_bytes(int dtype, int rank, int64_t b) {
    if (b < 0 || rank < 2 || (dtype != CHK_F32 && dtype != CHK_F64)) return -1;
    return carve(nullptr, dtype, rank, b, nullptr);
}

extern "C" int chk_eval_batch(const chk_eval_args* a, void* stream) {
    if (!a) { chk_set_error("chk_eval_batch: null args"); return CHK_EINVAL; }
    const int64_t b = a->b;
    if (b == 0) return CHK_OK;
    if (b < 0 || !a->queries || !a->entity || !a->rel || !a->rel_diag || !a->c_table || !a->hn_full || !a->counts || !a->target || !a->flags ||
        !a->scratch || a->scratch_bytes < chk_eval_scratch_bytes(a->dtype, a->rank, b) || ((a->bh == nullptr) != (a->bt == nullptr)) ||
        (a->shard_rows > 0 && (!a->shard_entity || !a->shard_hn)) || ((a->shard_bt == nullptr) != (a->bt == nullptr) && a->shard_rows > 0) ||
        (a->f_nkeys > 0 && (!a->f_keys || !a->f_indptr))) {
        chk_set_error("chk_eval_batch: bad argument (null table / scratch too small / bias tables inconsistent)"); return CHK_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    Scratch s;
    carve((char*)a->scratch, a->dtype, a->rank, b, &s);
    const bool f32 = a->dtype == CHK_F32;
    const int blocks = (int)((b + 255) / 256);
    if (f32) eval_split_kernel<float><<<blocks, 256, 0, st>>>(a->queries, b, (const float*)a->bh, s.heads, s.rels, s.tails, (float*)s.bhv, a->counts);
    else eval_split_kernel<double><<<blocks, 256, 0, st>>>(a->queries, b, (const double*)a->bh, s.heads, s.rels, s.tails, (double*)s.bhv, a->counts);
    CHK_CUDA_LAUNCH_CHECK("eval_split_kernel");
    int rc = chk_query_fwd(a->kind, a->dtype, a->rank, b, a->multi_c, a->entity, a->rel, a->rel_diag, a->ctx, a->c_table, s.heads, s.rels, s.q, s.c_out, stream);
    if (rc != CHK_OK) return rc;
    rc = chk_row_hnorm(a->dtype, a->rank, b, s.q, s.qn, stream);
    if (rc != CHK_OK) return rc;
    const void* bhv = a->bh ? (const void*)s.bhv : nullptr;
    unsigned tb = (unsigned)((b + 31) / 32);
    if (f32) { auto A = rargs<float>(a->rank, b, s.q, s.qn, bhv, nullptr, a->entity, a->hn_full, a->bt, a->n_entities); target_pairs_kernel<float><<<tb, 32, 0, st>>>(A, s.tails, (float*)a->target); }
    else { auto A = rargs<double>(a->rank, b, s.q, s.qn, bhv, nullptr, a->entity, a->hn_full, a->bt, a->n_entities); target_pairs_kernel<double><<<tb, 32, 0, st>>>(A, s.tails, (double*)a->target); }
    CHK_CUDA_LAUNCH_CHECK("target_pairs_kernel");
    if (a->shard_rows > 0) {
        if (a->algo == CHK_RANK_MMA) {
            if (!a->shadow || !a->workspace) { chk_set_error("chk_eval_batch: CHK_RANK_MMA needs shadow and workspace"); return CHK_EINVAL; }
            rc = chk_rank_counts_mma(a->dtype, a->rank, b, s.q, s.qn, bhv, a->target, a->shard_entity, a->shard_hn, a->shard_bt, a->shard_rows,
                                     a->shadow, a->workspace, a->workspace_bytes, a->counts, st);
        } else if (a->algo == CHK_RANK_FMA) {
            rc = chk_rank_counts_fma(a->dtype, a->rank, b, s.q, s.qn, bhv, a->target, a->shard_entity, a->shard_hn, a->shard_bt, a->shard_rows, a->counts, st);
        } else { chk_set_error("unknown rank algorithm %d", a->algo); return CHK_EINVAL; }
        if (rc != CHK_OK) return rc;
    }
    rc = chk_filter_lookup(a->queries, b, a->n_rel2, a->f_keys, a->f_nkeys, a->f_indptr, a->f_vals, a->shard_offset, a->shard_rows,
                           s.flt_indptr, s.flt_start, a->counts, a->flags, stream);
    if (rc != CHK_OK) return rc;
    if (a->shard_rows > 0 && a->f_nkeys > 0) {
        if (f32) { auto A = rargs<float>(a->rank, b, s.q, s.qn, bhv, a->target, a->shard_entity, a->shard_hn, a->shard_bt, a->shard_rows); A.counts = (unsigned long long*)a->counts;
                   filter_sub_ranged_kernel<float><<<148 * 8, 32, 0, st>>>(A, s.flt_indptr, s.flt_start, a->f_vals, a->shard_offset); }
        else { auto A = rargs<double>(a->rank, b, s.q, s.qn, bhv, a->target, a->shard_entity, a->shard_hn, a->shard_bt, a->shard_rows); A.counts = (unsigned long long*)a->counts;
               filter_sub_ranged_kernel<double><<<148 * 8, 32, 0, st>>>(A, s.flt_indptr, s.flt_start, a->f_vals, a->shard_offset); }
        CHK_CUDA_LAUNCH_CHECK("filter_sub_ranged_kernel");
    }
    return CHK_OK;
}
