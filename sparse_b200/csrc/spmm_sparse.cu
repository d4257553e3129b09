// spmm_sparse.cu -- K3: sparse-output forms of sparse x dense, and dense <-> CSR compaction.
//
// Replaces (sparse/numba_backend/_common.py):
//   _csr_ndarray_count_nnz :573-600 + _dot_csr_ndarray_sparse :758-804  -> b2s_spmm_csr_dense_flagged + b2s_dense_to_csr
//   _dot_coo_ndarray (sparse out) :1017-1072, _dot_ndarray_coo (sparse out) :1106-1158
//                                                           -> K1 (dense, exact) + b2s_dense_to_csr(mode = value != 0)
//   _csc_ndarray_count_nnz :603-632 + _dot_csc_ndarray_sparse :807-866
//                                                           -> b2s_dense_to_csr(sparsify) + wide-accumulate SpGEMM (spgemm.cu)
//   GCXS._prune  _compressed/compressed.py:816-842          -> b2s_flag_not_fill + b2s_scan_flags + b2s_compact + b2s_indptr_remap
#include <cub/cub.cuh>
#include <type_traits>

#include "common.cuh"

namespace b2s {

// out[i,j] = sum_k a[i,k]*b[k,j] with the reference's sparse-output arithmetic: the product is rounded to T,
// the running sum is kept in W (numba unifies `val = 0` with the product: int64 (+) T -> f64 for floats,
// i64 for ints), stored back as T.  flag[i,j] = any stored k of row i has b[k,j] != 0 (structural test on B).
template <typename T, typename W, typename I>
__global__ void __launch_bounds__(128)
spmm_flagged_kernel(int64_t M, int64_t N, const T *__restrict__ a_data, const I *__restrict__ a_indices,
                    const I *__restrict__ a_indptr, const T *__restrict__ B, int64_t ldb, T *__restrict__ out,
                    uint8_t *__restrict__ flags) {
    const int64_t row = blockIdx.x;
    const int64_t j = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    if (row >= M || j >= N) return;
    const int64_t s = (int64_t)a_indptr[row], e = (int64_t)a_indptr[row + 1];
    W acc = W(0);
    bool any = false;
    for (int64_t p = s; p < e; ++p) {
        const T bv = B[(int64_t)a_indices[p] * ldb + j];
        const T prod = mul_rn(a_data[p], bv);
        acc = add_rn(acc, (W)prod);
        any |= (bv != T(0));
    }
    out[row * N + j] = (T)acc;
    flags[row * N + j] = any ? 1 : 0;
}

// flags from values: mode 0 -> x != 0 (value compare: drops +-0, keeps NaN), mode 1 -> bits(x) != bits(+0)
template <typename T>
__global__ void dense_flags_kernel(const T *__restrict__ x, int64_t n, int mode, uint8_t *__restrict__ flags) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const T v = x[i];
        bool keep;
        if (mode == 0) keep = (v != T(0));
        else {
            if constexpr (sizeof(T) == 1) {
                keep = v != T(0);
            } else if constexpr (sizeof(T) == 4) {
                uint32_t u;
                memcpy(&u, &v, 4);
                keep = u != 0u;
            } else {
                uint64_t u;
                memcpy(&u, &v, 8);
                keep = u != 0ull;
            }
        }
        flags[i] = keep ? 1 : 0;
    }
}

template <typename T>
__global__ void dense_fill_kernel(const T *__restrict__ x, const uint8_t *__restrict__ flags,
                                  const int64_t *__restrict__ pos, int64_t M, int64_t N, int64_t *__restrict__ rows,
                                  int64_t *__restrict__ cols, T *__restrict__ data, int64_t *__restrict__ indptr,
                                  int64_t total) {
    const int64_t n = M * N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / N;
        const int64_t c = i - r * N;
        if (indptr && c == 0) indptr[r] = pos[i];
        if (flags[i]) {
            const int64_t p = pos[i];
            if (rows) rows[p] = r;
            cols[p] = c;
            data[p] = x[i];
        }
    }
    if (indptr && blockIdx.x == 0 && threadIdx.x == 0) indptr[M] = total;
}

// new_indptr[r] = pos[old_indptr[r]] (pos = exclusive scan of keep flags, pos[n] := total)
template <typename I>
__global__ void indptr_remap_kernel(const I *__restrict__ old_indptr, int64_t nrows, const int64_t *__restrict__ pos,
                                    int64_t n, int64_t total, I *__restrict__ new_indptr) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r <= nrows; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t o = (int64_t)old_indptr[r];
        new_indptr[r] = (I)(o >= n ? total : pos[o]);
    }
}

struct DensePlan {
    int dtype;
    int64_t M, N, total;
    const void *x;
    uint8_t *flags;
    int64_t *pos;
    bool own_flags;
    cudaStream_t stream;
};

static unsigned grid_n(int64_t n) {
    int64_t b = (n + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

}  // namespace b2s

using namespace b2s;

extern "C" {

int b2s_spmm_csr_dense_flagged(int dtype, int idx_bytes, int64_t M, int64_t K, int64_t N, const void *a_data_dev,
                               const void *a_indices_dev, const void *a_indptr_dev, const void *b_dev, int64_t ldb,
                               void *out_dev, uint8_t *flags_out_dev, void *stream) {
    B2S_REQUIRE(idx_bytes == 4 || idx_bytes == 8, B2S_ERR_INVALID, "spmm_flagged: idx_bytes");
    if (M == 0 || N == 0) return B2S_OK;
    B2S_REQUIRE(M <= 2147483647LL && (N + 127) / 128 <= 65535, B2S_ERR_OVERFLOW, "spmm_flagged: grid too large");
    cudaStream_t s = (cudaStream_t)stream;
    dim3 grid((unsigned)M, (unsigned)((N + 127) / 128));
#define B2S_F(T, W, I)                                                                                          \
    spmm_flagged_kernel<T, W, I><<<grid, 128, 0, s>>>(M, N, (const T *)a_data_dev, (const I *)a_indices_dev,    \
                                                      (const I *)a_indptr_dev, (const T *)b_dev, ldb, (T *)out_dev, \
                                                      flags_out_dev)
    if (idx_bytes == 4) {
        switch (dtype) {
            case B2S_F32: B2S_F(float, double, int32_t); break;
            case B2S_F64: B2S_F(double, double, int32_t); break;
            case B2S_I32: B2S_F(int32_t, int64_t, int32_t); break;
            case B2S_I64: B2S_F(int64_t, int64_t, int32_t); break;
            default: set_error("spmm_flagged: dtype %d", dtype); return B2S_ERR_UNSUPPORTED;
        }
    } else {
        switch (dtype) {
            case B2S_F32: B2S_F(float, double, int64_t); break;
            case B2S_F64: B2S_F(double, double, int64_t); break;
            case B2S_I32: B2S_F(int32_t, int64_t, int64_t); break;
            case B2S_I64: B2S_F(int64_t, int64_t, int64_t); break;
            default: set_error("spmm_flagged: dtype %d", dtype); return B2S_ERR_UNSUPPORTED;
        }
    }
#undef B2S_F
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

/* dense (M x N, contiguous) -> CSR/COO entries.  flags_or_null: externally computed keep flags (structural test),
 * otherwise derived from the values (mode 0: x != 0, mode 1: bits != +0).  Two-phase: begin returns the count. */
int b2s_dense_to_csr_begin(int dtype, int64_t M, int64_t N, const void *x_dev, const uint8_t *flags_or_null_dev,
                           int mode, void **plan_out, int64_t *nnz_out, void *stream) {
    B2S_REQUIRE(plan_out && nnz_out, B2S_ERR_INVALID, "dense_to_csr_begin: NULL output");
    const int64_t n = M * N;
    cudaStream_t s = (cudaStream_t)stream;
    DensePlan *pl = new DensePlan();
    pl->dtype = dtype;
    pl->M = M;
    pl->N = N;
    pl->x = x_dev;
    pl->stream = s;
    pl->flags = nullptr;
    pl->pos = nullptr;
    pl->own_flags = false;
    pl->total = 0;
    *plan_out = pl;
    *nnz_out = 0;
    if (n == 0) return B2S_OK;
    int rc;
    if (flags_or_null_dev) {
        pl->flags = const_cast<uint8_t *>(flags_or_null_dev);
    } else {
        if ((rc = scratch_alloc((void **)&pl->flags, (size_t)n, s))) return rc;
        pl->own_flags = true;
        switch (dtype) {
            case B2S_F32: dense_flags_kernel<float><<<grid_n(n), 256, 0, s>>>((const float *)x_dev, n, mode, pl->flags); break;
            case B2S_F64: dense_flags_kernel<double><<<grid_n(n), 256, 0, s>>>((const double *)x_dev, n, mode, pl->flags); break;
            case B2S_I32: dense_flags_kernel<int32_t><<<grid_n(n), 256, 0, s>>>((const int32_t *)x_dev, n, mode, pl->flags); break;
            case B2S_I64: dense_flags_kernel<int64_t><<<grid_n(n), 256, 0, s>>>((const int64_t *)x_dev, n, mode, pl->flags); break;
            case B2S_BOOL: dense_flags_kernel<uint8_t><<<grid_n(n), 256, 0, s>>>((const uint8_t *)x_dev, n, mode, pl->flags); break;
            default: set_error("dense_to_csr: dtype %d", dtype); return B2S_ERR_UNSUPPORTED;
        }
        B2S_CHECK_LAUNCH();
    }
    if ((rc = scratch_alloc((void **)&pl->pos, (size_t)n * 8, s))) return rc;
    size_t tb = 0;
    void *tmp = nullptr;
    for (int pass = 0; pass < 2; ++pass) {  // size query, then the scan; 64-bit offsets only when M*N needs them
        if (n >= 2147483647LL) B2S_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, pl->flags, pl->pos, (int64_t)n, s));
        else B2S_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, pl->flags, pl->pos, (int)n, s));
        if (pass == 0 && (rc = scratch_alloc(&tmp, tb, s))) return rc;
    }
    count_launch(2);
    int64_t lp = 0;
    uint8_t lf = 0;
    B2S_CUDA(cudaMemcpyAsync(&lp, pl->pos + (n - 1), 8, cudaMemcpyDeviceToHost, s));
    B2S_CUDA(cudaMemcpyAsync(&lf, pl->flags + (n - 1), 1, cudaMemcpyDeviceToHost, s));
    B2S_CUDA(cudaStreamSynchronize(s));
    scratch_free(tmp, s);
    pl->total = lp + (lf ? 1 : 0);
    *nnz_out = pl->total;
    return B2S_OK;
}

int b2s_dense_to_csr_finish(void *plan, int64_t *rows_out_or_null_dev, int64_t *cols_out_dev, void *data_out_dev,
                            int64_t *indptr_out_or_null_dev) {
    B2S_REQUIRE(plan != nullptr, B2S_ERR_INVALID, "dense_to_csr_finish: NULL plan");
    DensePlan *pl = (DensePlan *)plan;
    cudaStream_t s = pl->stream;
    const int64_t n = pl->M * pl->N;
    int rc = B2S_OK;
    if (n > 0) {
#define B2S_DF(T)                                                                                                     \
    dense_fill_kernel<T><<<grid_n(n), 256, 0, s>>>((const T *)pl->x, pl->flags, pl->pos, pl->M, pl->N,                  \
                                                   rows_out_or_null_dev, cols_out_dev, (T *)data_out_dev,               \
                                                   indptr_out_or_null_dev, pl->total)
        switch (pl->dtype) {
            case B2S_F32: B2S_DF(float); break;
            case B2S_F64: B2S_DF(double); break;
            case B2S_I32: B2S_DF(int32_t); break;
            case B2S_I64: B2S_DF(int64_t); break;
            case B2S_BOOL: B2S_DF(uint8_t); break;
            default: rc = B2S_ERR_UNSUPPORTED;
        }
#undef B2S_DF
        count_launch();
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            set_error("dense_fill launch: %s", cudaGetErrorString(e));
            rc = B2S_ERR_CUDA;
        }
    } else if (indptr_out_or_null_dev && pl->M >= 0) {
        cudaMemsetAsync(indptr_out_or_null_dev, 0, (size_t)(pl->M + 1) * 8, s);
    }
    if (pl->own_flags) scratch_free(pl->flags, s);
    scratch_free(pl->pos, s);
    delete pl;
    return rc;
}

int b2s_indptr_remap(int idx_bytes, const void *old_indptr_dev, int64_t nrows, const int64_t *pos_dev, int64_t n,
                     int64_t total, void *new_indptr_dev, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (idx_bytes == 4)
        indptr_remap_kernel<int32_t><<<grid_n(nrows + 1), 256, 0, s>>>((const int32_t *)old_indptr_dev, nrows, pos_dev, n, total, (int32_t *)new_indptr_dev);
    else if (idx_bytes == 8)
        indptr_remap_kernel<int64_t><<<grid_n(nrows + 1), 256, 0, s>>>((const int64_t *)old_indptr_dev, nrows, pos_dev, n, total, (int64_t *)new_indptr_dev);
    else {
        set_error("indptr_remap: idx_bytes %d", idx_bytes);
        return B2S_ERR_INVALID;
    }
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

}  // extern "C"
