// fused.cu -- K8 SDDMM and K9 MTTKRP: the two example paths of the reference, fused.
//
// SDDMM   examples/sddmm_example.py:51-52   out = s * (a @ b)
//   The reference forms the full dense product with BLAS and gathers it at s's coordinates
//   (_umath.py:606-608); at 1e6 x 1e6 that intermediate is 4 TB.  Here one warp owns one row i of the mask:
//   A[i,:] is staged once into shared memory with a 1-D bulk-TMA copy (cp.async.bulk + mbarrier) and kept in
//   registers; every stored (i,j) gathers the K-contiguous row j of B^T with 16-byte loads, 4 rows in flight,
//   and the 32 per-lane partial dot products of a chunk are combined with a 31-shuffle transpose-reduce so lane l
//   ends up with the result of entry l (coalesced store).  out = s.data * dot (one rounding, like s * dense).
//   HBM-bound: K*sizeof(T) bytes of B^T per stored entry; tensor cores do not pay below ~0.4 % mask density.
//
// MTTKRP  examples/mttkrp_example.py:51-52  out[i,j] = sum_{k,l} B[i,k,l] * D[l,j] * C[k,j]
//   The reference materialises two nnz x J broadcast products and reduces them; here one warp owns output row i
//   (entries of B sorted by (i,k,l)), lanes own columns j, and each entry gathers row l of D and row k of C.
#include "common.cuh"

namespace b2s {

constexpr unsigned FULLM = 0xffffffffu;

__device__ __forceinline__ uint32_t f_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <typename T, int VEC>
struct alignas(sizeof(T) * VEC) FPack {
    T v[VEC];
};

template <typename T, typename I, int KV>
__global__ void __launch_bounds__(256)
sddmm_kernel(int64_t M, const I *__restrict__ indptr, const I *__restrict__ cols, const T *__restrict__ svals,
             const T *__restrict__ A, int64_t lda, const T *__restrict__ Bt, int64_t ldbt, T *__restrict__ out) {
    constexpr int VEC = 16 / sizeof(T);
    constexpr int KROW = KV * 32 * VEC;  // K
    constexpr int WARPS = 8;
    constexpr int U = 4;
    __shared__ __align__(128) T s_a[WARPS][KROW];
    __shared__ uint64_t s_bar[WARPS];
    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    const uint32_t bar = f_smem_u32(&s_bar[w]);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t phase = 0;
    const uint64_t pol_b = policy_evict_last();
    const uint64_t pol_a = policy_evict_first();
    const int64_t warps_total = (int64_t)gridDim.x * WARPS;
    for (int64_t row = (int64_t)blockIdx.x * WARPS + w; row < M; row += warps_total) {
        const int64_t s = (int64_t)indptr[row], e = (int64_t)indptr[row + 1];
        if (s == e) continue;
        // stage A[row, :] with one bulk copy (TMA 1-D), then pull this lane's slices into registers
        if (lane == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(KROW * sizeof(T))) : "memory");
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
                    "r"(f_smem_u32(&s_a[w][0])),
                "l"(A + row * lda), "r"((uint32_t)(KROW * sizeof(T))), "r"(bar), "l"(pol_a)
                : "memory");
        }
        {
            asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                "SDDMM_WAIT:\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                "@p bra SDDMM_DONE;\n\t"
                "bra SDDMM_WAIT;\n\t"
                "SDDMM_DONE:\n\t"
                "}\n" ::"r"(bar),
                "r"(phase)
                : "memory");
            phase ^= 1;
        }
        FPack<T, VEC> a[KV];
#pragma unroll
        for (int q = 0; q < KV; ++q) a[q] = *reinterpret_cast<const FPack<T, VEC> *>(&s_a[w][(q * 32 + lane) * VEC]);
        __syncwarp();  // all lanes done reading s_a before the next row's copy may overwrite it

        for (int64_t base = s; base < e; base += 32) {
            const int64_t rem = e - base;
            const int cnt = rem > 32 ? 32 : (int)rem;
            I j = 0;
            T sv = T(0);
            if (lane < cnt) {
                j = ldg_stream(cols + base + lane);
                sv = ldg_stream(svals + base + lane);
            }
            T part[32];
#pragma unroll
            for (int n = 0; n < 32; n += U) {
                FPack<T, VEC> b[U][KV];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const I jj = __shfl_sync(FULLM, j, n + u);
                    if (n + u < cnt) {
                        const T *brow = Bt + (int64_t)jj * ldbt;
#pragma unroll
                        for (int q = 0; q < KV; ++q) {
                            const uint4 raw = ldg_nc_v4_hint(brow + (q * 32 + lane) * VEC, pol_b);
                            memcpy(&b[u][q], &raw, 16);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    T acc = T(0);
                    if (n + u < cnt) {
#pragma unroll
                        for (int q = 0; q < KV; ++q)
#pragma unroll
                            for (int v = 0; v < VEC; ++v) acc = fma(a[q].v[v], b[u][q].v[v], acc);
                    }
                    part[n + u] = acc;
                }
            }
            // transpose-reduce: 31 shuffles, lane l ends with the full dot product of entry l
#pragma unroll
            for (int off = 16, half = 16; off >= 1; off >>= 1, half >>= 1) {
                const bool up = (lane & off) != 0;
#pragma unroll
                for (int i = 0; i < half; ++i) {
                    const T send = up ? part[i] : part[i + half];
                    const T keep = up ? part[i + half] : part[i];
                    part[i] = keep + __shfl_xor_sync(FULLM, send, off);
                }
            }
            if (lane < cnt) out[base + lane] = mul_rn(sv, part[0]);
        }
    }
}

// generic-K fallback (any K, any alignment): lanes stride over k
template <typename T, typename I>
__global__ void __launch_bounds__(256)
sddmm_generic_kernel(int64_t M, int64_t K, const I *__restrict__ indptr, const I *__restrict__ cols,
                     const T *__restrict__ svals, const T *__restrict__ A, int64_t lda, const T *__restrict__ Bt,
                     int64_t ldbt, T *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < M; row += warps_total) {
        const int64_t s = (int64_t)indptr[row], e = (int64_t)indptr[row + 1];
        const T *arow = A + row * lda;
        for (int64_t p = s; p < e; ++p) {
            const T *brow = Bt + (int64_t)cols[p] * ldbt;
            T acc = T(0);
            for (int64_t k = lane; k < K; k += 32) acc = fma(arow[k], brow[k], acc);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULLM, acc, o);
            if (lane == 0) out[p] = mul_rn(svals[p], acc);
        }
    }
}

// MTTKRP: one warp per output row i, lanes over columns j (tiles of 32), 4 entries in flight
template <typename T, typename I>
__global__ void __launch_bounds__(256)
mttkrp_kernel(int64_t Mi, int64_t J, const I *__restrict__ indptr, const I *__restrict__ kk, const I *__restrict__ ll,
              const T *__restrict__ vals, const T *__restrict__ D, int64_t ldd, const T *__restrict__ C, int64_t ldc,
              T *__restrict__ out, int64_t ldo) {
    constexpr int U = 4;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= Mi) return;
    const int64_t row = warp;
    const int64_t j = (int64_t)blockIdx.y * 32 + lane;
    const bool jok = j < J;
    const int64_t s = (int64_t)indptr[row], e = (int64_t)indptr[row + 1];
    T acc = T(0);
    for (int64_t base = s; base < e; base += 32) {
        const int64_t rem = e - base;
        const int cnt = rem > 32 ? 32 : (int)rem;
        I k = 0, l = 0;
        T v = T(0);
        if (lane < cnt) {
            k = kk[base + lane];
            l = ll[base + lane];
            v = vals[base + lane];
        }
        for (int n = 0; n < cnt; n += U) {
            T dv[U], cv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const I lk = __shfl_sync(FULLM, l, (n + u) & 31);
                const I ck = __shfl_sync(FULLM, k, (n + u) & 31);
                dv[u] = T(0);
                cv[u] = T(0);
                if (n + u < cnt && jok) {
                    dv[u] = __ldg(D + (int64_t)lk * ldd + j);
                    cv[u] = __ldg(C + (int64_t)ck * ldc + j);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const T bv = __shfl_sync(FULLM, v, (n + u) & 31);
                if (n + u < cnt) acc = add_rn(acc, mul_rn(mul_rn(bv, dv[u]), cv[u]));
            }
        }
    }
    if (jok) out[row * ldo + j] = acc;
}

}  // namespace b2s

using namespace b2s;

extern "C" {

/*
 * SDDMM: out_vals[p] = s_vals[p] * dot(A[i_p, :], Bt[j_p, :]) for every stored (i_p, j_p) of the mask, given as CSR
 * (indptr over rows i, cols j).  Bt is b transposed, (N x K) row-major, so the gathered vectors are contiguous.
 */
int b2s_sddmm(int dtype, int idx_bytes, int64_t M, int64_t N, int64_t K, const void *indptr_dev,
              const void *cols_dev, const void *s_vals_dev, const void *a_dev, int64_t lda, const void *bt_dev,
              int64_t ldbt, void *out_vals_dev, void *stream) {
    B2S_REQUIRE(idx_bytes == 4 || idx_bytes == 8, B2S_ERR_INVALID, "sddmm: idx_bytes");
    B2S_REQUIRE(dtype == B2S_F32 || dtype == B2S_F64, B2S_ERR_UNSUPPORTED, "sddmm: dtype %d (f32/f64 only)", dtype);
    if (M == 0) return B2S_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t es = dtype_size(dtype);
    const int vec = (int)(16 / es);
    const bool aligned = (((uintptr_t)a_dev | (uintptr_t)bt_dev) & 15) == 0 && (lda * es) % 16 == 0 &&
                         (ldbt * es) % 16 == 0;
    int kv = 0;
    if (aligned && K % (32 * vec) == 0) kv = (int)(K / (32 * vec));
    int64_t blocks = (M + 7) / 8;
    const int64_t cap = (int64_t)num_sms() * 8;
    if (blocks > cap) blocks = cap;
#define B2S_SD(T, I, KV)                                                                                              \
    sddmm_kernel<T, I, KV><<<(unsigned)blocks, 256, 0, s>>>(M, (const I *)indptr_dev, (const I *)cols_dev,            \
                                                            (const T *)s_vals_dev, (const T *)a_dev, lda,              \
                                                            (const T *)bt_dev, ldbt, (T *)out_vals_dev)
#define B2S_SDG(T, I)                                                                                                 \
    sddmm_generic_kernel<T, I><<<(unsigned)blocks, 256, 0, s>>>(M, K, (const I *)indptr_dev, (const I *)cols_dev,     \
                                                                (const T *)s_vals_dev, (const T *)a_dev, lda,          \
                                                                (const T *)bt_dev, ldbt, (T *)out_vals_dev)
#define B2S_SDK(T, I)                      \
    switch (kv) {                          \
        case 1: B2S_SD(T, I, 1); break;    \
        case 2: B2S_SD(T, I, 2); break;    \
        case 4: B2S_SD(T, I, 4); break;    \
        default: B2S_SDG(T, I); break;     \
    }
    if (dtype == B2S_F32) {
        if (idx_bytes == 4) { B2S_SDK(float, int32_t) } else { B2S_SDK(float, int64_t) }
    } else {
        if (idx_bytes == 4) { B2S_SDK(double, int32_t) } else { B2S_SDK(double, int64_t) }
    }
#undef B2S_SDK
#undef B2S_SDG
#undef B2S_SD
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

/*
 * MTTKRP: out[i, j] = sum over stored (i,k,l) of B: val * D[l, j] * C[k, j].  B is given sorted by (i,k,l) as
 * indptr over i plus the k and l coordinate arrays.  out is dense (I x J), fully written.
 */
int b2s_mttkrp(int dtype, int idx_bytes, int64_t I_, int64_t J, const void *indptr_dev, const void *k_dev,
               const void *l_dev, const void *vals_dev, const void *d_dev, int64_t ldd, const void *c_dev, int64_t ldc,
               void *out_dev, int64_t ldo, void *stream) {
    B2S_REQUIRE(idx_bytes == 4 || idx_bytes == 8, B2S_ERR_INVALID, "mttkrp: idx_bytes");
    B2S_REQUIRE(dtype == B2S_F32 || dtype == B2S_F64, B2S_ERR_UNSUPPORTED, "mttkrp: dtype %d (f32/f64 only)", dtype);
    if (I_ == 0 || J == 0) return B2S_OK;
    cudaStream_t s = (cudaStream_t)stream;
    dim3 grid((unsigned)((I_ + 7) / 8), (unsigned)((J + 31) / 32));
#define B2S_MT(T, I)                                                                                                 \
    mttkrp_kernel<T, I><<<grid, 256, 0, s>>>(I_, J, (const I *)indptr_dev, (const I *)k_dev, (const I *)l_dev,        \
                                             (const T *)vals_dev, (const T *)d_dev, ldd, (const T *)c_dev, ldc,       \
                                             (T *)out_dev, ldo)
    if (dtype == B2S_F32) {
        if (idx_bytes == 4) B2S_MT(float, int32_t); else B2S_MT(float, int64_t);
    } else {
        if (idx_bytes == 4) B2S_MT(double, int32_t); else B2S_MT(double, int64_t);
    }
#undef B2S_MT
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

}  // extern "C"
