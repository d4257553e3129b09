// spmm_skew.cu -- nnz-balanced handling of long rows for K1 (power-law matrices).
//
// In exact-order mode a row is inherently sequential per output column, so a row-split kernel makes a warp spend
// row_nnz x latency / U on one row (a row of 85 k entries alone takes ~20 ms), and a CTA of 8 rows lives as long as its
// longest row (the other seven warps idle: occupancy collapses on power-law matrices).  Rows longer than
// max(512, 4 x the mean row length) are therefore taken out of the row-split grid (its warps skip rows flagged in
// `skip`) and given to a column-split kernel: one CTA per long row, one THREAD per output column, and the gathered B
// rows of the next 112 stored entries always in flight in a shared-memory ring filled by 16-byte cp.async copies
// (LDGSTS: no registers, no warp stalls; 57 KB in flight per CTA instead of 16 loads per lane), consumed in stored
// order.  The long-row kernel is launched first on a side stream and runs concurrently with the row-split kernel
// (longest-processing-time-first), and every output element still accumulates in stored order with separate
// product / sum roundings -> results stay bit-identical.  B rows that are not 16-byte aligned take the older
// register-staged kernel (16 independent 4-byte gathers in flight per lane).
#include "common.cuh"

namespace b2s {

constexpr int kLongRowMin = 512;   // stored entries: never treat shorter rows as long
constexpr int kLongCap = 1 << 18;  // long rows handled by the column-split kernel (further ones stay row-split)

template <typename I>
__global__ void mark_long_rows_kernel(int64_t M, const I *__restrict__ indptr, uint8_t *__restrict__ skip,
                                      int64_t *__restrict__ list, unsigned int *__restrict__ count) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= M) return;
    const int64_t n = (int64_t)indptr[row + 1] - (int64_t)indptr[row];
    const int64_t mean4 = 4 * (((int64_t)indptr[M] - (int64_t)indptr[0]) / (M > 0 ? M : 1));
    const int64_t thr = mean4 > kLongRowMin ? mean4 : kLongRowMin;
    uint8_t flag = 0;
    if (n > thr) {
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
    const int lane = threadIdx.x & 31;
    for (unsigned li = blockIdx.x; li < n_long; li += gridDim.x) {
    const int64_t row = list[li];
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
}

// ---- shared-memory ring variant ---------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_16(void *smem_dst, const void *gmem_src, int src_bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <typename T, typename I>
struct RingCfg {
    static constexpr int ROWB = 512;                       // bytes of a B row one CTA stages (its column panel)
    static constexpr int CW = ROWB / (int)sizeof(T);       // columns per CTA = threads per CTA
    static constexpr int CPR = ROWB / 16;                  // 16-byte chunks per staged row
    static constexpr int RPP = CW / CPR;                   // rows filled by one pass of the CTA
    static constexpr int G = 16;                           // rows per commit group
    static constexpr int NG = 8;                           // groups in the ring (NG - 1 in flight)
    static constexpr int ABLK = 2048;                      // stored entries of the row staged at a time (cols + vals)
    static constexpr size_t ring_b = (size_t)NG * G * ROWB;
    static constexpr size_t smem = ring_b + (size_t)ABLK * (sizeof(I) + sizeof(T));
};

template <typename T, typename I>
__global__ void __launch_bounds__(RingCfg<T, I>::CW)
spmm_long_rows_ring_kernel(int64_t N, const T *__restrict__ a_data, const I *__restrict__ a_indices,
                           const I *__restrict__ a_indptr, const T *__restrict__ B, int64_t ldb, T *__restrict__ C,
                           int64_t ldc, const int64_t *__restrict__ list, const unsigned int *__restrict__ count) {
    using R = RingCfg<T, I>;
    constexpr int G = R::G, NG = R::NG, ROWB = R::ROWB, CPR = R::CPR, RPP = R::RPP, NT = R::CW;
    extern __shared__ __align__(16) unsigned char sm[];
    unsigned char *ring = sm;
    T *s_vals = reinterpret_cast<T *>(sm + R::ring_b);
    I *s_cols = reinterpret_cast<I *>(sm + R::ring_b + (size_t)R::ABLK * sizeof(T));
    unsigned n_long = *count;
    if (n_long > (unsigned)kLongCap) n_long = kLongCap;
    const int tid = threadIdx.x;
    const int64_t col0 = (int64_t)blockIdx.y * NT;
    const int64_t col = col0 + tid;
    const bool col_ok = col < N;
    const int chunk = tid % CPR, rsub = tid / CPR;
    // bytes of this thread's 16-byte chunk that lie inside the row (0 .. 16): the tail of the last panel is zero-filled
    int64_t rem = (N - col0) * (int64_t)sizeof(T) - (int64_t)chunk * 16;
    const int src_bytes = rem >= 16 ? 16 : (rem > 0 ? (int)rem : 0);
    const unsigned char *bbase = reinterpret_cast<const unsigned char *>(B + col0) + (src_bytes ? chunk * 16 : 0);

    for (unsigned li = blockIdx.x; li < n_long; li += gridDim.x) {
        const int64_t row = list[li];
        const int64_t start = (int64_t)a_indptr[row], end = (int64_t)a_indptr[row + 1];
        T acc = T(0);
        for (int64_t blk0 = start; blk0 < end; blk0 += R::ABLK) {
            const int nb = (int)((end - blk0) < R::ABLK ? (end - blk0) : R::ABLK);
            __syncthreads();  // the previous block of entries is fully consumed
            for (int e = tid; e < nb; e += NT) {
                s_cols[e] = ldg_stream(a_indices + blk0 + e);
                s_vals[e] = ldg_stream(a_data + blk0 + e);
            }
            __syncthreads();
            const int ngroups = (nb + G - 1) / G;
            auto issue = [&](int g) {
                if (g < ngroups) {
                    unsigned char *slot = ring + (size_t)(g % NG) * G * ROWB;
#pragma unroll
                    for (int p = 0; p < G / RPP; ++p) {
                        const int r = p * RPP + rsub;
                        const int e = g * G + r;
                        if (e < nb)
                            cp_async_16(slot + (size_t)r * ROWB + chunk * 16,
                                        bbase + (size_t)s_cols[e] * (size_t)ldb * sizeof(T), src_bytes);
                    }
                }
                cp_async_commit();  // committed even when empty: the group count stays uniform
            };
            for (int g = 0; g < NG - 1; ++g) issue(g);
            for (int g = 0; g < ngroups; ++g) {
                issue(g + NG - 1);
                cp_async_wait<NG - 1>();  // group g has landed (this thread's copies)...
                __syncthreads();          // ... and everybody else's
                const T *slot = reinterpret_cast<const T *>(ring + (size_t)(g % NG) * G * ROWB);
                const int e0 = g * G;
#pragma unroll
                for (int r = 0; r < G; ++r)
                    if (e0 + r < nb) acc = add_rn(acc, mul_rn(s_vals[e0 + r], slot[(size_t)r * NT + tid]));
                __syncthreads();  // the slot is refilled by the next iteration's issue
            }
        }
        if (col_ok) C[row * ldc + col] = acc;
    }
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
static thread_local unsigned int *t_rowctr = nullptr;
unsigned int *skew_row_counter() { return t_rowctr; }

template <typename T, typename I>
int skew_begin(int64_t M, int64_t N, const void *ad, const void *ai, const void *ap, const void *b, int64_t ldb,
               void *out, int64_t ldc, cudaStream_t s, uint8_t **skip_out, bool aligned16) {
    SkewState &st = skew_state();
    *skip_out = nullptr;
    if (!st.ok) return B2S_OK;  // no side stream: plain row-split
    uint8_t *skip = nullptr;
    int rc;
    if ((rc = scratch_alloc((void **)&skip, (size_t)M, s))) return rc;
    if ((rc = scratch_alloc((void **)&t_list, (size_t)kLongCap * 8, s))) return rc;
    if ((rc = scratch_alloc((void **)&t_count, 4, s))) return rc;
    if ((rc = scratch_alloc((void **)&t_rowctr, 64 * sizeof(unsigned int), s))) return rc;
    B2S_CUDA(cudaMemsetAsync(t_count, 0, 4, s));
    mark_long_rows_kernel<I><<<(unsigned)((M + 255) / 256), 256, 0, s>>>(M, (const I *)ap, skip, t_list, t_count);
    B2S_CHECK_LAUNCH();
    // fork: the long-row kernel runs on the side stream, concurrently with the row-split kernel on `s`
    B2S_CUDA(cudaEventRecord(st.fork, s));
    B2S_CUDA(cudaStreamWaitEvent(st.side, st.fork, 0));
    if (aligned16) {
        using R = RingCfg<T, I>;
        auto kern = spmm_long_rows_ring_kernel<T, I>;
        B2S_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)R::smem));
        const unsigned gy = (unsigned)((N + R::CW - 1) / R::CW);
        // persistent over the list (the count lives on the device): enough CTAs for every SM's shared memory
        dim3 grid((unsigned)(num_sms() * 2), gy);
        kern<<<grid, R::CW, R::smem, st.side>>>(N, (const T *)ad, (const I *)ai, (const I *)ap, (const T *)b, ldb,
                                                (T *)out, ldc, t_list, t_count);
    } else {
        int threads = (int)((N + 31) / 32) * 32;
        if (threads > 1024) threads = 1024;
        const unsigned gy = (unsigned)((N + threads - 1) / threads);
        dim3 grid(4096u, gy);
        spmm_long_rows_kernel<T, I><<<grid, threads, 0, st.side>>>(N, (const T *)ad, (const I *)ai, (const I *)ap,
                                                                  (const T *)b, ldb, (T *)out, ldc, t_list, t_count);
    }
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
    scratch_free(t_rowctr, s);
    t_list = nullptr;
    t_count = nullptr;
    t_rowctr = nullptr;
    return B2S_OK;
}

template int skew_begin<float, int32_t>(int64_t, int64_t, const void *, const void *, const void *, const void *, int64_t, void *, int64_t, cudaStream_t, uint8_t **, bool);
template int skew_begin<float, int64_t>(int64_t, int64_t, const void *, const void *, const void *, const void *, int64_t, void *, int64_t, cudaStream_t, uint8_t **, bool);
template int skew_begin<double, int32_t>(int64_t, int64_t, const void *, const void *, const void *, const void *, int64_t, void *, int64_t, cudaStream_t, uint8_t **, bool);
template int skew_begin<double, int64_t>(int64_t, int64_t, const void *, const void *, const void *, const void *, int64_t, void *, int64_t, cudaStream_t, uint8_t **, bool);
template int skew_begin<int32_t, int32_t>(int64_t, int64_t, const void *, const void *, const void *, const void *, int64_t, void *, int64_t, cudaStream_t, uint8_t **, bool);
template int skew_begin<int32_t, int64_t>(int64_t, int64_t, const void *, const void *, const void *, const void *, int64_t, void *, int64_t, cudaStream_t, uint8_t **, bool);
template int skew_begin<int64_t, int32_t>(int64_t, int64_t, const void *, const void *, const void *, const void *, int64_t, void *, int64_t, cudaStream_t, uint8_t **, bool);
template int skew_begin<int64_t, int64_t>(int64_t, int64_t, const void *, const void *, const void *, const void *, int64_t, void *, int64_t, cudaStream_t, uint8_t **, bool);

}  // namespace b2s
