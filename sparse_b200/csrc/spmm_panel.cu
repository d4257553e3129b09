// spmm_panel.cu -- K1p: CSR x dense with COLUMN-PANEL passes that keep the active slice of B resident in L2.
//
// Same contract as K1 (bit-identical to _dot_csr_ndarray, sparse/numba_backend/_common.py:720-755).  At C2 the
// one-pass kernel is DRAM-bound: B (512 MB) cannot live in the 126 MB L2, so 89 % of the 512-byte row gathers
// go to HBM (44.6 GB of traffic).  Here the K axis is cut into P panels of <= ~48 MB of B rows; pass p processes,
// for every row of A, only the stored entries whose column falls into panel p.  Rows are sorted by column, so these
// are a contiguous run that starts where pass p-1 stopped (a per-row cursor), and the partial sums are carried
// through C in fp32 -- a store followed by a load of the same float is exact, and panels are visited in ascending
// column order, so every out[i,j] still sees the reference's operation sequence.
// DRAM traffic becomes  B once + A ~once + C (2P-1) x  ~= 10 GB instead of 44.6 GB; the gathers are served by L2.
// B rows carry an evict_last policy, the A stream, the cursor and the C read-modify-write are streaming accesses.
#include "common.cuh"

namespace b2s {

template <typename T, int VEC>
struct alignas(sizeof(T) * VEC) PPack {
    T v[VEC];
};

__device__ __forceinline__ uint4 ldg_cs_v4(const void *p) {
    uint4 r;
    asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

template <typename T, typename I, int U>
__global__ void __launch_bounds__(256)
spmm_panel_kernel(int64_t M, int64_t N, const T *__restrict__ a_data, const I *__restrict__ a_indices,
                  const I *__restrict__ a_indptr, const I *__restrict__ cursor_in, I *__restrict__ cursor_out,
                  const T *__restrict__ B, int64_t ldb,
                  T *__restrict__ C, int64_t ldc, int64_t col_hi, int first, int last) {
    constexpr int VEC = 16 / sizeof(T);
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int64_t col0 = ((int64_t)blockIdx.y * 32 + lane) * VEC;
    const bool col_ok = col0 < N;
    int64_t base = first ? (int64_t)a_indptr[row] : (int64_t)cursor_in[row];
    const int64_t end = (int64_t)a_indptr[row + 1];
    const uint64_t pol_b = policy_evict_last();
    const T *bcol = B + col0;
    T *crow = C + row * ldc + col0;

    PPack<T, VEC> acc;
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc.v[k] = T(0);
    // nothing of this row in this panel (and not the initialising pass): leave C and the cursor untouched
    I col = 0;
    T val = T(0);
    int cnt = 0;
    {
        const int64_t rem = end - base;
        cnt = rem > 32 ? 32 : (rem > 0 ? (int)rem : 0);
        if (lane < cnt) {
            col = ldg_stream(a_indices + base + lane);
            val = ldg_stream(a_data + base + lane);
        }
    }
    unsigned m = __ballot_sync(FULL, lane < cnt && (int64_t)col < col_hi);
    int take = __popc(m);
    if (take == 0 && !first) {
        // column tiles of the same row read cursor_in concurrently: the cursor is double-buffered per pass
        if (!last && blockIdx.y == 0 && lane == 0) cursor_out[row] = (I)base;
        return;
    }
    if (!first && col_ok) {
        const uint4 u = ldg_cs_v4(crow);
        memcpy(&acc, &u, 16);
    }
    while (true) {
#pragma unroll 1
        for (int j = 0; j < take; j += U) {
            PPack<T, VEC> bv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const I c = __shfl_sync(FULL, col, (j + u) & 31);
                if (j + u < take && col_ok) {
                    const uint4 raw = ldg_nc_v4_hint(bcol + (int64_t)c * ldb, pol_b);
                    memcpy(&bv[u], &raw, 16);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const T v = __shfl_sync(FULL, val, (j + u) & 31);
                if (j + u < take) {
#pragma unroll
                    for (int k = 0; k < VEC; ++k) acc.v[k] = add_rn(acc.v[k], mul_rn(v, bv[u].v[k]));
                }
            }
        }
        base += take;
        if (take < 32 || base >= end) break;  // reached the panel boundary or the end of the row
        const int64_t rem = end - base;
        cnt = rem > 32 ? 32 : (int)rem;
        col = 0;
        if (lane < cnt) {
            col = ldg_stream(a_indices + base + lane);
            val = ldg_stream(a_data + base + lane);
        }
        m = __ballot_sync(FULL, lane < cnt && (int64_t)col < col_hi);
        take = __popc(m);
        if (take == 0) break;
    }
    if (!last && blockIdx.y == 0 && lane == 0) cursor_out[row] = (I)base;
    if (col_ok) {
        uint4 u;
        memcpy(&u, &acc, 16);
        stg_cs_v4(crow, u);
    }
}

// rows sorted by column? (flag[0] |= 1 when some row has a descending neighbour pair)
template <typename I>
__global__ void rows_sorted_kernel(int64_t M, const I *__restrict__ indptr, const I *__restrict__ indices,
                                   int *__restrict__ flag) {
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    bool bad = false;
    for (int64_t row = warp; row < M; row += nwarps) {
        const int64_t s = (int64_t)indptr[row], e = (int64_t)indptr[row + 1];
        for (int64_t p = s + lane; p + 1 < e; p += 32) bad |= indices[p + 1] < indices[p];
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(flag, 1);
}

template <typename I>
__global__ void max_row_kernel(int64_t M, const I *__restrict__ indptr, unsigned long long *__restrict__ out) {
    unsigned long long m = 0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < M; r += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long n = (unsigned long long)((int64_t)indptr[r + 1] - (int64_t)indptr[r]);
        m = n > m ? n : m;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long v = __shfl_xor_sync(0xffffffffu, m, o);
        m = v > m ? v : m;
    }
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

template <typename T, typename I>
static int run_panels(int64_t M, int64_t K, int64_t N, const void *ad, const void *ai, const void *ap, const void *b,
                      int64_t ldb, void *out, int64_t ldc, int n_panels, cudaStream_t s) {
    constexpr int VEC = 16 / sizeof(T);
    I *cursor = nullptr;
    int rc = scratch_alloc((void **)&cursor, (size_t)M * sizeof(I) * 2, s);
    if (rc) return rc;
    const int64_t gx = (M + 7) / 8;
    const int64_t gy = (N + 32 * VEC - 1) / (32 * VEC);
    B2S_REQUIRE(gx <= 2147483647LL && gy <= 65535, B2S_ERR_OVERFLOW, "spmm_panels: grid too large");
    dim3 grid((unsigned)gx, (unsigned)gy);
    const int64_t width = (K + n_panels - 1) / n_panels;
    for (int p = 0; p < n_panels; ++p) {
        const int64_t hi = (p + 1 == n_panels) ? (int64_t)1 << 62 : (int64_t)(p + 1) * width;
        spmm_panel_kernel<T, I, 8><<<grid, 256, 0, s>>>(M, N, (const T *)ad, (const I *)ai, (const I *)ap,
                                                       cursor + (size_t)((p + 1) & 1) * M, cursor + (size_t)(p & 1) * M,
                                                       (const T *)b, ldb, (T *)out, ldc, hi, p == 0 ? 1 : 0,
                                                       p + 1 == n_panels ? 1 : 0);
        B2S_CHECK_LAUNCH();
    }
    return scratch_free(cursor, s);
}

}  // namespace b2s

using namespace b2s;

extern "C" {

/* 1 if every row of the CSR has non-decreasing column indices (precondition of the panel passes). Synchronises. */
int b2s_csr_rows_sorted(int idx_bytes, int64_t M, const void *indptr_dev, const void *indices_dev, int *sorted_host,
                        void *stream) {
    B2S_REQUIRE(sorted_host != nullptr, B2S_ERR_INVALID, "csr_rows_sorted: NULL output");
    *sorted_host = 1;
    if (M == 0) return B2S_OK;
    cudaStream_t s = (cudaStream_t)stream;
    int *flag = nullptr;
    int rc = scratch_alloc((void **)&flag, 4, s);
    if (rc) return rc;
    B2S_CUDA(cudaMemsetAsync(flag, 0, 4, s));
    int64_t blocks = (M * 32 + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    if (idx_bytes == 4) rows_sorted_kernel<int32_t><<<(unsigned)blocks, 256, 0, s>>>(M, (const int32_t *)indptr_dev, (const int32_t *)indices_dev, flag);
    else rows_sorted_kernel<int64_t><<<(unsigned)blocks, 256, 0, s>>>(M, (const int64_t *)indptr_dev, (const int64_t *)indices_dev, flag);
    B2S_CHECK_LAUNCH();
    int h = 0;
    B2S_CUDA(cudaMemcpyAsync(&h, flag, 4, cudaMemcpyDeviceToHost, s));
    B2S_CUDA(cudaStreamSynchronize(s));
    scratch_free(flag, s);
    *sorted_host = h ? 0 : 1;
    return B2S_OK;
}

/* Largest number of stored entries in a row (decides whether the nnz-balanced long-row path is worth enabling). */
int b2s_csr_max_row_nnz(int idx_bytes, int64_t M, const void *indptr_dev, int64_t *max_host, void *stream) {
    B2S_REQUIRE(max_host != nullptr, B2S_ERR_INVALID, "csr_max_row_nnz: NULL output");
    *max_host = 0;
    if (M == 0) return B2S_OK;
    cudaStream_t s = (cudaStream_t)stream;
    unsigned long long *d = nullptr;
    int rc = scratch_alloc((void **)&d, 8, s);
    if (rc) return rc;
    B2S_CUDA(cudaMemsetAsync(d, 0, 8, s));
    int64_t blocks = (M + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    if (idx_bytes == 4) max_row_kernel<int32_t><<<(unsigned)blocks, 256, 0, s>>>(M, (const int32_t *)indptr_dev, d);
    else max_row_kernel<int64_t><<<(unsigned)blocks, 256, 0, s>>>(M, (const int64_t *)indptr_dev, d);
    B2S_CHECK_LAUNCH();
    unsigned long long h = 0;
    B2S_CUDA(cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, s));
    B2S_CUDA(cudaStreamSynchronize(s));
    scratch_free(d, s);
    *max_host = (int64_t)h;
    return B2S_OK;
}

/*
 * K1 with explicit scheduling control.  n_panels <= 1: the one-pass kernel (same as b2s_spmm_csr_dense).
 * n_panels >= 2: column-panel passes (rows MUST be sorted by column -- check with b2s_csr_rows_sorted; 16-byte aligned
 * fp32/fp64 operands with N * sizeof(T) a multiple of 16); n_panels == 0: choose automatically from the size of B
 * (about 48 MB of B per panel) when `rows_sorted` is 1 and the matrix has enough entries per row (`nnz` >= 0 known).
 */
int b2s_spmm_csr_dense_ex(int dtype, int idx_bytes, int64_t M, int64_t K, int64_t N, int64_t nnz,
                          const void *a_data_dev, const void *a_indices_dev, const void *a_indptr_dev,
                          const void *b_dev, int64_t ldb, void *out_dev, int64_t ldc, int n_panels, int rows_sorted,
                          int long_rows, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const size_t es = dtype_size(dtype);
    const bool vec_ok = (dtype == B2S_F32 || dtype == B2S_F64) && es != 0 && (N * es) % 16 == 0 &&
                        (ldb * es) % 16 == 0 && (ldc * es) % 16 == 0 && (((uintptr_t)b_dev | (uintptr_t)out_dev) & 15) == 0;
    if (n_panels == 0) {
        // Measured on B200 (profiles/r01_k1_panels.json): 7.74 ms one-pass vs 8.5 / 9.2 / 13.4 ms with 2 / 8 / 16 panels
        // at C2.  The one-pass kernel already runs at the L2 (LTS) throughput ceiling (~6.8 TB/s of gathered bytes), so
        // turning DRAM misses into L2 hits buys nothing and every extra pass adds its C read-modify-write.  The
        // automatic choice is therefore always the one-pass kernel; panels stay available on explicit request.
        (void)nnz;
        n_panels = 1;
    }
    if (n_panels <= 1 || !vec_ok || M == 0 || N == 0) {
        set_call_skew(long_rows ? 1 : 0);
        const int rc = spmm_csr_dense_impl(dtype, idx_bytes, M, K, N, a_data_dev, a_indices_dev, a_indptr_dev, b_dev,
                                           ldb, out_dev, ldc, s);
        set_call_skew(-1);
        return rc;
    }
    B2S_REQUIRE(rows_sorted == 1, B2S_ERR_INVALID, "spmm panels: rows must be sorted by column (rows_sorted=1)");
    B2S_REQUIRE(idx_bytes == 4 || idx_bytes == 8, B2S_ERR_INVALID, "spmm panels: idx_bytes");
    if (dtype == B2S_F32) {
        return idx_bytes == 4 ? run_panels<float, int32_t>(M, K, N, a_data_dev, a_indices_dev, a_indptr_dev, b_dev, ldb, out_dev, ldc, n_panels, s)
                              : run_panels<float, int64_t>(M, K, N, a_data_dev, a_indices_dev, a_indptr_dev, b_dev, ldb, out_dev, ldc, n_panels, s);
    }
    return idx_bytes == 4 ? run_panels<double, int32_t>(M, K, N, a_data_dev, a_indices_dev, a_indptr_dev, b_dev, ldb, out_dev, ldc, n_panels, s)
                          : run_panels<double, int64_t>(M, K, N, a_data_dev, a_indices_dev, a_indptr_dev, b_dev, ldb, out_dev, ldc, n_panels, s);
}

}  // extern "C"
