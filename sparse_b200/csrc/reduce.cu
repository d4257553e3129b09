// reduce.cu -- K7: grouped (segmented) reductions over sorted COO / GCXS entries.
//
// Replaces _calc_counts_invidx + ufunc.reduceat (sparse/numba_backend/_coo/core.py:1601-1661), the GCXS
// reduceat over indptr (_compressed/compressed.py:354-372) and the fill-value correction of
// SparseArray.reduce (_sparse_array.py:405-422).
//
// Input: group ids (non-decreasing; produced from the kept-axes linear index of every stored entry) and the
// matching values.  One streaming pass (CUB ReduceByKey: decoupled look-back segmented scan, deterministic for
// a given input size) yields per-group value, first position and count.  Floating-point association differs
// from NumPy's reduceat (whose order is itself unspecified), so parity is tolerance-based (tests use 1e-6 f32).
#include <cub/cub.cuh>
#include <type_traits>

#include "common.cuh"

namespace b2s {

enum RedOp { R_ADD = 0, R_MUL = 1, R_MAX = 2, R_MIN = 3, R_AND = 4, R_OR = 5, R_BAND = 6, R_BOR = 7, R_BXOR = 8 };

template <typename T>
__device__ __forceinline__ bool isnan_t(T x) {
    return x != x;
}

struct FAdd {
    template <typename T>
    __device__ __forceinline__ T operator()(const T &a, const T &b) const { return add_rn(a, b); }
};
struct FMul {
    template <typename T>
    __device__ __forceinline__ T operator()(const T &a, const T &b) const { return mul_rn(a, b); }
};
struct FMax {  // np.maximum: NaN propagates
    template <typename T>
    __device__ __forceinline__ T operator()(const T &a, const T &b) const {
        if (isnan_t(a)) return a;
        if (isnan_t(b)) return b;
        return a >= b ? a : b;
    }
};
struct FMin {
    template <typename T>
    __device__ __forceinline__ T operator()(const T &a, const T &b) const {
        if (isnan_t(a)) return a;
        if (isnan_t(b)) return b;
        return a <= b ? a : b;
    }
};
struct FAnd {
    template <typename T>
    __device__ __forceinline__ T operator()(const T &a, const T &b) const { return (T)((a != T(0)) && (b != T(0))); }
};
struct FOr {
    template <typename T>
    __device__ __forceinline__ T operator()(const T &a, const T &b) const { return (T)((a != T(0)) || (b != T(0))); }
};
struct FBand {
    template <typename T>
    __device__ __forceinline__ T operator()(const T &a, const T &b) const { return a & b; }
};
struct FBor {
    template <typename T>
    __device__ __forceinline__ T operator()(const T &a, const T &b) const { return a | b; }
};
struct FBxor {
    template <typename T>
    __device__ __forceinline__ T operator()(const T &a, const T &b) const { return a ^ b; }
};

__global__ void group_ids_kernel(const int64_t *__restrict__ keys, int64_t n, int64_t ncols, int64_t *__restrict__ g) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        g[i] = keys[i] / ncols;
}

// counts from run starts: counts[g] = start[g+1] - start[g]
__global__ void run_starts_kernel(const int64_t *__restrict__ gid, int64_t n, const int64_t *__restrict__ pos,
                                  int64_t *__restrict__ starts, int64_t ngroups) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (i == 0 || gid[i] != gid[i - 1]) starts[pos[i]] = i;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) starts[ngroups] = n;
}

__global__ void counts_from_starts_kernel(const int64_t *__restrict__ starts, int64_t ngroups,
                                          int64_t *__restrict__ counts) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ngroups; i += (int64_t)gridDim.x * blockDim.x)
        counts[i] = starts[i + 1] - starts[i];
}

// fill-value correction (_sparse_array.py:405-422).
//  add: v = v + (n_fill == 0 ? 0 : fill * n_fill)        (reduce_super_ufunc = multiply)
//  mul: v = v * (n_fill == 0 ? 1 : fill ** n_fill)       (reduce_super_ufunc = power)
//  others: if count != n_cols: v = op(v, fill)
template <typename T, typename Op>
__global__ void fill_fix_kernel(T *__restrict__ vals, const int64_t *__restrict__ counts, int64_t ngroups,
                                int64_t ncols, T fill, int op, Op f) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ngroups;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t nf = ncols - counts[i];
        T v = vals[i];
        if (op == R_ADD) {
            const T contrib = nf == 0 ? T(0) : mul_rn(fill, (T)nf);
            v = add_rn(v, contrib);
        } else if (op == R_MUL) {
            T contrib = T(1);
            if (nf != 0) {
                if constexpr (std::is_floating_point<T>::value) contrib = (T)pow((double)fill, (double)nf);
                else {
                    T b = fill;
                    int64_t e = nf;
                    T r = 1;
                    while (e > 0) {
                        if (e & 1) r = mul_rn(r, b);
                        b = mul_rn(b, b);
                        e >>= 1;
                    }
                    contrib = r;
                }
            }
            v = mul_rn(v, contrib);
        } else if (nf != 0) {
            v = f(v, fill);
        }
        vals[i] = v;
    }
}

static unsigned rgrid(int64_t n) {
    int64_t b = (n + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

template <typename T, typename Op>
static int reduce_by_key_t(const int64_t *gid, const T *vals, int64_t n, int64_t *uniq_out, T *agg_out,
                           int64_t *nruns_dev, cudaStream_t s) {
    size_t tb = 0;
    B2S_CUDA(cub::DeviceReduce::ReduceByKey(nullptr, tb, gid, uniq_out, vals, agg_out, nruns_dev, Op(), (int)n, s));
    void *tmp = nullptr;
    int rc = scratch_alloc(&tmp, tb, s);
    if (rc) return rc;
    B2S_CUDA(cub::DeviceReduce::ReduceByKey(tmp, tb, gid, uniq_out, vals, agg_out, nruns_dev, Op(), (int)n, s));
    count_launch(2);
    return scratch_free(tmp, s);
}

template <typename T>
static int reduce_dispatch_op(int op, const int64_t *gid, const T *vals, int64_t n, int64_t *uniq_out, T *agg_out,
                              int64_t *nruns_dev, cudaStream_t s) {
    switch (op) {
        case R_ADD: return reduce_by_key_t<T, FAdd>(gid, vals, n, uniq_out, agg_out, nruns_dev, s);
        case R_MUL: return reduce_by_key_t<T, FMul>(gid, vals, n, uniq_out, agg_out, nruns_dev, s);
        case R_MAX: return reduce_by_key_t<T, FMax>(gid, vals, n, uniq_out, agg_out, nruns_dev, s);
        case R_MIN: return reduce_by_key_t<T, FMin>(gid, vals, n, uniq_out, agg_out, nruns_dev, s);
        default: break;
    }
    if constexpr (std::is_integral<T>::value) {
        switch (op) {
            case R_AND: return reduce_by_key_t<T, FAnd>(gid, vals, n, uniq_out, agg_out, nruns_dev, s);
            case R_OR: return reduce_by_key_t<T, FOr>(gid, vals, n, uniq_out, agg_out, nruns_dev, s);
            case R_BAND: return reduce_by_key_t<T, FBand>(gid, vals, n, uniq_out, agg_out, nruns_dev, s);
            case R_BOR: return reduce_by_key_t<T, FBor>(gid, vals, n, uniq_out, agg_out, nruns_dev, s);
            case R_BXOR: return reduce_by_key_t<T, FBxor>(gid, vals, n, uniq_out, agg_out, nruns_dev, s);
            default: break;
        }
    }
    set_error("reduce: op %d unsupported for this dtype", op);
    return B2S_ERR_UNSUPPORTED;
}

}  // namespace b2s

using namespace b2s;

extern "C" {

/* group id of every entry = key / ncols (keys = linear index over (kept axes..., reduced axes...)). */
int b2s_group_ids(const int64_t *keys_dev, int64_t n, int64_t ncols, int64_t *gid_out_dev, void *stream) {
    if (n == 0) return B2S_OK;
    B2S_REQUIRE(ncols >= 1, B2S_ERR_INVALID, "group_ids: ncols must be >= 1");
    group_ids_kernel<<<rgrid(n), 256, 0, (cudaStream_t)stream>>>(keys_dev, n, ncols, gid_out_dev);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

/*
 * Reduce `vals` over runs of equal group id.  Outputs are caller-allocated with room for n entries
 * (n_groups <= n); n_groups is returned on the host (one stream sync).  counts_out[g] = run length.
 */
int b2s_reduce_by_key(int dtype, int op, const int64_t *gid_dev, const void *vals_dev, int64_t n,
                      int64_t *groups_out_dev, void *vals_out_dev, int64_t *counts_out_dev, int64_t *n_groups_host,
                      void *stream) {
    B2S_REQUIRE(n_groups_host != nullptr, B2S_ERR_INVALID, "reduce_by_key: NULL n_groups");
    *n_groups_host = 0;
    if (n == 0) return B2S_OK;
    B2S_REQUIRE(n < 2147483647LL, B2S_ERR_OVERFLOW, "reduce_by_key: n=%lld exceeds 2^31", (long long)n);
    cudaStream_t s = (cudaStream_t)stream;
    int64_t *nruns = nullptr;
    int rc = scratch_alloc((void **)&nruns, 8, s);
    if (rc) return rc;
    switch (dtype) {
        case B2S_F32: rc = reduce_dispatch_op<float>(op, gid_dev, (const float *)vals_dev, n, groups_out_dev, (float *)vals_out_dev, nruns, s); break;
        case B2S_F64: rc = reduce_dispatch_op<double>(op, gid_dev, (const double *)vals_dev, n, groups_out_dev, (double *)vals_out_dev, nruns, s); break;
        case B2S_I32: rc = reduce_dispatch_op<int32_t>(op, gid_dev, (const int32_t *)vals_dev, n, groups_out_dev, (int32_t *)vals_out_dev, nruns, s); break;
        case B2S_I64: rc = reduce_dispatch_op<int64_t>(op, gid_dev, (const int64_t *)vals_dev, n, groups_out_dev, (int64_t *)vals_out_dev, nruns, s); break;
        case B2S_BOOL: rc = reduce_dispatch_op<uint8_t>(op, gid_dev, (const uint8_t *)vals_dev, n, groups_out_dev, (uint8_t *)vals_out_dev, nruns, s); break;
        default: set_error("reduce_by_key: dtype %d", dtype); rc = B2S_ERR_UNSUPPORTED;
    }
    if (rc) return rc;
    int64_t h = 0;
    B2S_CUDA(cudaMemcpyAsync(&h, nruns, 8, cudaMemcpyDeviceToHost, s));
    B2S_CUDA(cudaStreamSynchronize(s));
    scratch_free(nruns, s);
    *n_groups_host = h;
    // run lengths: reduce-by-key of ones would be a second pass over the keys; derive them from the run starts
    if (counts_out_dev && h > 0) {
        uint8_t *heads = nullptr;
        int64_t *pos = nullptr, *starts = nullptr;
        if ((rc = scratch_alloc((void **)&pos, (size_t)n * 8, s))) return rc;
        if ((rc = scratch_alloc((void **)&starts, (size_t)(h + 1) * 8, s))) return rc;
        if ((rc = scratch_alloc((void **)&heads, (size_t)n, s))) return rc;
        // heads[i] = gid[i] != gid[i-1]; pos = exclusive scan; starts[pos[i]] = i
        if ((rc = b2s_flag_heads(gid_dev, n, heads, stream))) return rc;
        size_t tb = 0;
        B2S_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, heads, pos, (int)n, s));
        void *tmp = nullptr;
        if ((rc = scratch_alloc(&tmp, tb, s))) return rc;
        B2S_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, heads, pos, (int)n, s));
        count_launch(2);
        run_starts_kernel<<<rgrid(n), 256, 0, s>>>(gid_dev, n, pos, starts, h);
        B2S_CHECK_LAUNCH();
        counts_from_starts_kernel<<<rgrid(h), 256, 0, s>>>(starts, h, counts_out_dev);
        B2S_CHECK_LAUNCH();
        scratch_free(tmp, s);
        scratch_free(heads, s);
        scratch_free(pos, s);
        scratch_free(starts, s);
    }
    return B2S_OK;
}

/* Apply the fill-value contribution of SparseArray.reduce to the per-group results, in place. */
int b2s_reduce_fill_fix(int dtype, int op, void *vals_dev, const int64_t *counts_dev, int64_t n_groups, int64_t ncols,
                        const void *fill_host, void *stream) {
    if (n_groups == 0) return B2S_OK;
    cudaStream_t s = (cudaStream_t)stream;
#define B2S_FF(T, OPF)                                                                                       \
    {                                                                                                        \
        T f;                                                                                                 \
        memcpy(&f, fill_host, sizeof(T));                                                                    \
        fill_fix_kernel<T, OPF><<<rgrid(n_groups), 256, 0, s>>>((T *)vals_dev, counts_dev, n_groups, ncols, f, op, OPF()); \
    }
#define B2S_FF_T(T)                                             \
    switch (op) {                                               \
        case R_MAX: B2S_FF(T, FMax) break;                      \
        case R_MIN: B2S_FF(T, FMin) break;                      \
        case R_AND: B2S_FF(T, FAnd) break;                      \
        case R_OR: B2S_FF(T, FOr) break;                        \
        default: B2S_FF(T, FAdd) break; /* add / mul handled by `op` inside the kernel */ \
    }
    switch (dtype) {
        case B2S_F32: B2S_FF_T(float) break;
        case B2S_F64: B2S_FF_T(double) break;
        case B2S_I32: B2S_FF_T(int32_t) break;
        case B2S_I64: B2S_FF_T(int64_t) break;
        case B2S_BOOL: B2S_FF_T(uint8_t) break;
        default: set_error("reduce_fill_fix: dtype %d", dtype); return B2S_ERR_UNSUPPORTED;
    }
#undef B2S_FF_T
#undef B2S_FF
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

}  // extern "C"
