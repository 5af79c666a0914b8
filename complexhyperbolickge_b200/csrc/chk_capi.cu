// C-ABI glue: error reporting and the chk_rank_counts dispatcher.
#include <cstdarg>
#include <cstdio>
#include "chk_common.cuh"

static thread_local char g_err[512] = "";

void chk_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int chk_rank_counts_fma(int dtype, int rank, int64_t b, const void* q, const void* qn, const void* bh_vals,
                        const void* target, const void* entity, const void* hn, const void* bt,
                        int64_t n_rows, int64_t* counts, cudaStream_t st);
int chk_filter_subtract(int dtype, int rank, int64_t b, const void* q, const void* qn, const void* bh_vals,
                        const void* target, const void* entity, const void* hn, const void* bt,
                        int64_t n_rows, int64_t shard_offset, const int64_t* indptr, const int64_t* fidx,
                        int64_t total, int64_t* counts, cudaStream_t st);
int chk_rank_counts_mma(int dtype, int rank, int64_t b, const void* q, const void* qn, const void* bh_vals,
                        const void* target, const void* entity, const void* hn, const void* bt,
                        int64_t n_rows, const void* shadow, void* workspace, int64_t workspace_bytes,
                        int64_t* counts, cudaStream_t st);

extern "C" int chk_abi_version(void) { return 4; }
extern "C" const char* chk_last_error(void) { return g_err; }

extern "C" int chk_rank_counts(int algo, int dtype, int rank, int64_t b, const void* q, const void* qn,
                               const void* bh_vals, const void* target, const void* entity, const void* hn,
                               const void* bt, int64_t n_rows, int64_t shard_offset,
                               const int64_t* filter_indptr, const int64_t* filter_idx, int64_t filter_total,
                               const void* shadow, void* workspace, int64_t workspace_bytes,
                               int64_t* counts, void* stream) {
    if (b == 0 || n_rows == 0) return CHK_OK;
    if (b < 0 || n_rows < 0 || rank < 2 || !q || !qn || !target || !entity || !hn || !counts ||
        ((bh_vals == nullptr) != (bt == nullptr)) || (filter_total > 0 && (!filter_indptr || !filter_idx))) {
        chk_set_error("chk_rank_counts: bad argument");
        return CHK_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (algo == CHK_RANK_FMA) {
        rc = chk_rank_counts_fma(dtype, rank, b, q, qn, bh_vals, target, entity, hn, bt, n_rows, counts, st);
    } else if (algo == CHK_RANK_MMA) {
        if (!shadow || !workspace) { chk_set_error("CHK_RANK_MMA needs shadow and workspace"); return CHK_EINVAL; }
        rc = chk_rank_counts_mma(dtype, rank, b, q, qn, bh_vals, target, entity, hn, bt, n_rows, shadow, workspace,
                                 workspace_bytes, counts, st);
    } else {
        chk_set_error("unknown rank algorithm %d", algo);
        return CHK_EINVAL;
    }
    if (rc != CHK_OK) return rc;
    return chk_filter_subtract(dtype, rank, b, q, qn, bh_vals, target, entity, hn, bt, n_rows, shard_offset,
                               filter_indptr, filter_idx, filter_total, counts, st);
}
