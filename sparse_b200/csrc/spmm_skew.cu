// spmm_skew.cu -- nnz-balanced handling of long rows for K1 (power-law matrices).
//
// In exact-order mode a row is inherently sequential per output column, so a row-split kernel makes a warp spend
// row_nnz x latency / U on one row: a row of 85 k entries alone takes ~20 ms.  Rows longer than kLongRow are
// therefore taken out of the row-split grid (its warps skip rows flagged in `skip`) and given to a column-split
// kernel: one CTA per long row, one LANE per output column (so a 128-column row gets 4 warps), 16 independent
// 4-byte gathers in flight per lane.  The long-row kernel is launched first on a side stream and runs concurrently
// with the row-split kernel (longest-processing-time-first), and every output element still accumulates in stored
// order with separate product / sum roundings -> results stay bit-identical.
#include "common.cuh"

namespace b2s {

constexpr int kLongRow = 4096;   // stored entries
constexpr int kLongCap = 2048;   // long rows handled by the column-split kernel (further ones stay row-split)

template <typename I>
__global__ void mark_long_rows_kernel(int64_t M, const I *__restrict__ indptr, uint8_t *__restrict__ skip,
                                      int64_t *__restrict__ list, unsigned int *__restrict__ count) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= M) return;
    const int64_t n = (int64_t)indptr[row + 1] - (int64_t)indptr[row];
    uint8_t flag = 0;
    if (n > kLongRow) {
        const unsigned slot = atomicAdd(count, 1u);
        if (slot < (unsigned)kLongCap) {
            list[slot] = row;
            flag = 1;
        }
    }
    skip[row] = flag;
}

template <typename T, typename I>
__global__ void __launch_bounds__(1024)
spmm_long_rows_kernel(int64_t N, const T *__restrict__ a_data, const I *__restrict__ a_indices,
                      const I *__restrict__ a_indptr, const T *__restrict__ B, int64_t ldb, T *__restrict__ C,
                      int64_t ldc, const int64_t *__restrict__ list, const unsigned int *__restrict__ count) {
    constexpr int U = 16;
    constexpr unsigned FULL = 0xffffffffu;
    unsigned n_long = *count;
    if (n_long > (unsigned)kLongCap) n_long = kLongCap;
    if (blockIdx.x >= n_long) return;
    const int64_t row = list[blockIdx.x];
    const int lane = threadIdx.x & 31;
    const int64_t col = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;  // one output column per lane
    const bool col_ok = col < N;
    int64_t base = (int64_t)a_indptr[row];
    const int64_t end = (int64_t)a_indptr[row + 1];
    const uint64_t pol_b = policy_evict_last();
    const T *bcol = B + col;
    T acc = T(0);
    I cn = 0;
    T vn = T(0);
    if (base + lane < end) {
        cn = ldg_stream(a_indices + base + lane);
        vn = ldg_stream(a_data + base + lane);
    }
    while (base < end) {
        const I c32 = cn;
        const T v32 = vn;
        const int64_t rem = end - base;
        const int cnt = rem > 32 ? 32 : (int)rem;
        const int64_t nb = base + 32;
        if (nb + lane < end) {
            cn = ldg_stream(a_indices + nb + lane);
            vn = ldg_stream(a_data + nb + lane);
        }
#pragma unroll 1
        for (int j = 0; j < cnt; j += U) {
            T bv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const I c = __shfl_sync(FULL, c32, (j + u) & 31);
                bv[u] = T(0);
                if (j + u < cnt && col_ok) {
                    if constexpr (sizeof(T) == 4) {
                        unsigned r;
                        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;"
                                     : "=r"(r) : "l"(bcol + (int64_t)c * ldb), "l"(pol_b));
                        memcpy(&bv[u], &r, 4);
                    } else {
                        unsigned long long r;
                        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;"
                                     : "=l"(r) : "l"(bcol + (int64_t)c * ldb), "l"(pol_b));
                        memcpy(&bv[u], &r, 8);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const T v = __shfl_sync(FULL, v32, (j + u) & 31);
                if (j + u < cnt) acc = add_rn(acc, mul_rn(v, bv[u]));
            }
        }
        base = nb;
    }
    if (col_ok) C[row * ldc + col] = acc;
}

struct SkewState {
    cudaStream_t side = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    bool ok = false;
};
static SkewState &skew_state() {
    static SkewState st;
    if (!st.ok) {
        if (cudaStreamCreateWithFlags(&st.side, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&st.fork, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&st.join, cudaEventDisableTiming) == cudaSuccess)
            st.ok = true;
    }
    return st;
}

// scratch of the in-flight call (freed in skew_end)
static thread_local int64_t *t_list = nullptr;
static thread_local unsigned int *t_count = nullptr;

template <typename T, typename I>
int skew_begin(int64_t M, int64_t N, const void *ad, const void *ai, const void *ap, const void *b, int64_t ldb,
               void *out, int64_t ldc, cudaStream_t s, uint8_t **skip_out) {
    SkewState &st = skew_state();
    *skip_out = nullptr;
    if (!st.ok) return B2S_OK;  // no side stream: plain row-split
    uint8_t *skip = nullptr;
    int rc;
    if ((rc = scratch_alloc((void **)&skip, (size_t)M, s))) return rc;
    if ((rc = scratch_alloc((void **)&t_list, (size_t)kLongCap * 8, s))) return rc;
    if ((rc = scratch_alloc((void **)&t_count, 4, s))) return rc;
    B2S_CUDA(cudaMemsetAsync(t_count, 0, 4, s));
    mark_long_rows_kernel<I><<<(unsigned)((M + 255) / 256), 256, 0, s>>>(M, (const I *)ap, skip, t_list, t_count);
    B2S_CHECK_LAUNCH();
    // fork: the long-row kernel runs on the side stream, concurrently with the row-split kernel on `s`
    B2S_CUDA(cudaEventRecord(st.fork, s));
    B2S_CUDA(cudaStreamWaitEvent(st.side, st.fork, 0));
    int threads = (int)((N + 31) / 32) * 32;
    if (threads > 1024) threads = 1024;
    const unsigned gy = (unsigned)((N + threads - 1) / threads);
    dim3 grid((unsigned)kLongCap, gy);
    spmm_long_rows_kernel<T, I><<<grid, threads, 0, st.side>>>(N, (const T *)ad, (const I *)ai, (const I *)ap,
                                                              (const T *)b, ldb, (T *)out, ldc, t_list, t_count);
    B2S_CHECK_LAUNCH();
    B2S_CUDA(cudaEventRecord(st.join, st.side));
    *skip_out = skip;
    return B2S_OK;
}

int skew_end(cudaStream_t s, uint8_t *skip) {
    SkewState &st = skew_state();
    if (skip == nullptr) return B2S_OK;
    B2S_CUDA(cudaStreamWaitEvent(s, st.join, 0));  // join before anything later on `s` (incl. the frees) proceeds
    scratch_free(skip, s);
    scratch_free(t_list, s);
    scratch_free(t_count, s);
    t_list = nullptr;
    t_count = nullptr;
    return B2S_OK;
}

template int skew_begin<float, int32_t>(int64_t, int64_t, const void *, const void *, const void *, const void *, int64_t, void *, int64_t, cudaStream_t, uint8_t **);
template int skew_begin<float, int64_t>(int64_t, int64_t, const void *, const void *, const void *, const void *, int64_t, void *, int64_t, cudaStream_t, uint8_t **);
template int skew_begin<double, int32_t>(int64_t, int64_t, const void *, const void *, const void *, const void *, int64_t, void *, int64_t, cudaStream_t, uint8_t **);
template int skew_begin<double, int64_t>(int64_t, int64_t, const void *, const void *, const void *, const void *, int64_t, void *, int64_t, cudaStream_t, uint8_t **);
template int skew_begin<int32_t, int32_t>(int64_t, int64_t, const void *, const void *, const void *, const void *, int64_t, void *, int64_t, cudaStream_t, uint8_t **);
template int skew_begin<int32_t, int64_t>(int64_t, int64_t, const void *, const void *, const void *, const void *, int64_t, void *, int64_t, cudaStream_t, uint8_t **);
template int skew_begin<int64_t, int32_t>(int64_t, int64_t, const void *, const void *, const void *, const void *, int64_t, void *, int64_t, cudaStream_t, uint8_t **);
template int skew_begin<int64_t, int64_t>(int64_t, int64_t, const void *, const void *, const void *, const void *, int64_t, void *, int64_t, cudaStream_t, uint8_t **);

}  // namespace b2s
