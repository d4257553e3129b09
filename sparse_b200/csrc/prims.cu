// prims.cu -- streaming device primitives shared by COO canonicalisation, format conversion,
// elemwise and reductions: linearise / unravel coordinates, stable key sort,
// head flags, compaction, gathers, casts, indptr construction, dense transpose.
//
// Reference functions these replace (sparse/numba_backend/):
//   linear_loc                      _coo/common.py:56-64      -> b2s_coo_linearize
//   COO._sort_indices               _coo/core.py:1294-1317    -> b2s_keys_flags + b2s_sort_keys + b2s_gather
//   COO._sum_duplicates             _coo/core.py:1319-1353    -> b2s_flag_heads + b2s_scan_flags + b2s_segment_sum
//   COO._prune / GCXS._prune        _coo/core.py:1355-1371    -> b2s_flag_not_fill + b2s_scan_flags + b2s_compact
//   _from_coo (bincount + cumsum)   _compressed/compressed.py:25-77 -> b2s_indptr_from_sorted
//   uncompress_dimension            _compressed/convert.py:81-87    -> b2s_rows_from_indptr
// Device-wide sort and scan use CUB (header-only CCCL shipped with the toolkit).
#include <cub/cub.cuh>

#include "common.cuh"
#include "fastdiv.cuh"

namespace b2s {

// 16-byte element (complex128) for the element-size generic movers: moved as one 128-bit word, compared bitwise
struct alignas(16) U128 {
    uint64_t lo, hi;
    __host__ __device__ bool operator!=(const U128 &o) const { return lo != o.lo || hi != o.hi; }
};

constexpr int kMaxDims = 16;
struct DimPack {
    int64_t stride[kMaxDims];  // multiplier of each *input* row (0 for dropped rows)
    FastDiv fext[kMaxDims];    // extents, prepared for multiply-high division
};

static inline unsigned grid_for(int64_t n, int threads = 256, int per_sm = 16) {
    int64_t b = (n + threads - 1) / threads;
    const int64_t cap = (int64_t)num_sms() * per_sm;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

#define B2S_GRID_STRIDE(i, n)                                                          \
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n);          \
         i += (int64_t)gridDim.x * blockDim.x)

// ---- linearise: key[i] = sum_d coords[d][i] * stride[d] --------------------------------------
template <typename I>
__global__ void linearize_kernel(int ndim, int64_t nnz, const I *__restrict__ coords, int64_t row_stride, DimPack dp,
                                 int64_t *__restrict__ keys) {
    B2S_GRID_STRIDE(i, nnz) {
        int64_t k = 0;
        for (int d = 0; d < ndim; ++d) k += (int64_t)coords[(int64_t)d * row_stride + i] * dp.stride[d];
        keys[i] = k;
    }
}

// ---- trace selector of einsum: keep entry i iff coords[d][i] == coords[first[d]][i] for every d ---
struct DiagPack {
    int first[kMaxDims];
};
template <typename I>
__global__ void diag_flags_kernel(int ndim, int64_t nnz, const I *__restrict__ coords, int64_t row_stride, DiagPack dg,
                                  uint8_t *__restrict__ flags) {
    B2S_GRID_STRIDE(i, nnz) {
        bool keep = true;
        for (int d = 0; d < ndim; ++d) {
            const int f = dg.first[d];
            if (f != d) keep &= coords[(int64_t)d * row_stride + i] == coords[(int64_t)f * row_stride + i];
        }
        flags[i] = keep ? 1 : 0;
    }
}

// ---- unravel: coords[d][i] = (key / stride[d]) % extent[d] ----------------------------------
template <typename I>
__global__ void unravel_kernel(int ndim, int64_t nnz, const int64_t *__restrict__ keys, DimPack dp,
                               I *__restrict__ coords, int64_t row_stride) {
    B2S_GRID_STRIDE(i, nnz) {
        uint64_t k = (uint64_t)keys[i];
        for (int d = ndim - 1; d >= 0; --d) {
            uint64_t q, r;
            dp.fext[d].divmod(k, q, r);
            coords[(int64_t)d * row_stride + i] = (I)r;
            k = q;
        }
    }
}

// ---- basic indexing x[ints / slices]: per entry, test every axis against its (start, step, count) range and emit the
// ---- key over the result shape (integer-indexed axes carry out_stride 0) -----------------------------------------
struct SlicePack {
    FastDiv fext[kMaxDims];     // input extents
    FastDiv fstep[kMaxDims];    // |step|
    int64_t start[kMaxDims];
    int64_t count[kMaxDims];    // len(range(start, stop, step)); 1 for an integer index
    int64_t ostride[kMaxDims];  // stride of the axis in the result (0 when dropped)
    int neg[kMaxDims];          // step < 0
};
__global__ void slice_keys_kernel(int ndim, int64_t nnz, const int64_t *__restrict__ keys, SlicePack sp,
                                  uint8_t *__restrict__ flags, int64_t *__restrict__ okeys) {
    B2S_GRID_STRIDE(i, nnz) {
        uint64_t k = (uint64_t)keys[i];
        int64_t ok = 0;
        bool keep = true;
        for (int d = ndim - 1; d >= 0; --d) {
            uint64_t q, c;
            sp.fext[d].divmod(k, q, c);
            k = q;
            const int64_t off = sp.neg[d] ? sp.start[d] - (int64_t)c : (int64_t)c - sp.start[d];
            uint64_t j = 0, r = 0;
            if (off >= 0) sp.fstep[d].divmod((uint64_t)off, j, r);
            keep &= off >= 0 && r == 0 && (int64_t)j < sp.count[d];
            ok += (int64_t)j * sp.ostride[d];
        }
        flags[i] = keep ? 1 : 0;
        okeys[i] = ok;
    }
}

// ---- sortedness / duplicate flags over a key array ------------------------------------------
__global__ void keys_flags_kernel(const int64_t *__restrict__ keys, int64_t n, int32_t *__restrict__ flags) {
    bool unsorted = false, dup = false;
    B2S_GRID_STRIDE(i, n - 1) {
        const int64_t a = keys[i], b = keys[i + 1];
        unsorted |= (b < a);
        dup |= (b == a);
    }
    if (__any_sync(0xffffffffu, unsorted) && (threadIdx.x & 31) == 0) atomicOr(flags + 0, 1);
    if (__any_sync(0xffffffffu, dup) && (threadIdx.x & 31) == 0) atomicOr(flags + 1, 1);
}

__global__ void iota_kernel(int64_t *__restrict__ out, int64_t n) {
    B2S_GRID_STRIDE(i, n) out[i] = i;
}

template <typename T>
__global__ void gather_kernel(const T *__restrict__ in, const int64_t *__restrict__ perm, int64_t n,
                              T *__restrict__ out) {
    B2S_GRID_STRIDE(i, n) out[i] = in[perm[i]];
}

// head flag: first element of every run of equal keys
__global__ void flag_heads_kernel(const int64_t *__restrict__ keys, int64_t n, uint8_t *__restrict__ flags) {
    B2S_GRID_STRIDE(i, n) flags[i] = (i == 0) || (keys[i] != keys[i - 1]);
}

// prune flag: bitwise difference from the fill value (`equivalent`, _utils.py:448-452)
template <typename U>
__global__ void flag_not_fill_kernel(const U *__restrict__ data, int64_t n, U fill, uint8_t *__restrict__ flags) {
    B2S_GRID_STRIDE(i, n) flags[i] = data[i] != fill;
}

template <typename T>
__global__ void compact_kernel(const T *__restrict__ in, const uint8_t *__restrict__ flags,
                               const int64_t *__restrict__ pos, int64_t n, T *__restrict__ out) {
    B2S_GRID_STRIDE(i, n) if (flags[i]) out[pos[i]] = in[i];
}

// 2-D strided compaction of coordinate rows: out[d][pos[i]] = in[d][i]
template <typename T>
__global__ void compact_rows_kernel(int nrows, const T *__restrict__ in, int64_t in_stride,
                                    const uint8_t *__restrict__ flags, const int64_t *__restrict__ pos, int64_t n,
                                    T *__restrict__ out, int64_t out_stride) {
    B2S_GRID_STRIDE(i, n) {
        if (flags[i]) {
            const int64_t p = pos[i];
            for (int d = 0; d < nrows; ++d) out[(int64_t)d * out_stride + p] = in[(int64_t)d * in_stride + i];
        }
    }
}

// sum of each run of duplicates, in stored order (one thread per run head; runs are short)
template <typename T>
__global__ void segment_sum_kernel(const T *__restrict__ data, const uint8_t *__restrict__ heads,
                                   const int64_t *__restrict__ pos, int64_t n, T *__restrict__ out) {
    B2S_GRID_STRIDE(i, n) {
        if (heads[i]) {
            T s = data[i];
            for (int64_t j = i + 1; j < n && !heads[j]; ++j) s = add_rn(s, data[j]);
            out[pos[i]] = s;
        }
    }
}

// indptr[r] = first position whose row id >= r (rows sorted ascending); r in [0, nrows]
template <typename I, typename O>
__global__ void indptr_from_sorted_kernel(const I *__restrict__ rows, int64_t n, int64_t nrows,
                                          O *__restrict__ indptr) {
    B2S_GRID_STRIDE(r, nrows + 1) {
        int64_t lo = 0, hi = n;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if ((int64_t)rows[mid] < r) lo = mid + 1;
            else hi = mid;
        }
        indptr[r] = (O)lo;
    }
}

// keys (2-D linear over (nrows, ncols)) sorted -> indices = key % ncols, plus indptr via search on key / ncols
template <typename O>
__global__ void split_keys_kernel(const int64_t *__restrict__ keys, int64_t n, int64_t ncols, O *__restrict__ rows,
                                  O *__restrict__ cols) {
    B2S_GRID_STRIDE(i, n) {
        const int64_t k = keys[i];
        const int64_t r = k / ncols;
        if (rows) rows[i] = (O)r;
        cols[i] = (O)(k - r * ncols);
    }
}

template <typename O>
__global__ void indptr_from_keys_kernel(const int64_t *__restrict__ keys, int64_t n, int64_t nrows, int64_t ncols,
                                        O *__restrict__ indptr) {
    B2S_GRID_STRIDE(r, nrows + 1) {
        const int64_t target = r * ncols;  // first key of row r
        int64_t lo = 0, hi = n;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (keys[mid] < target) lo = mid + 1;
            else hi = mid;
        }
        indptr[r] = (O)lo;
    }
}

// rows[p] = r for p in [indptr[r], indptr[r+1])
template <typename I, typename O>
__global__ void rows_from_indptr_kernel(const I *__restrict__ indptr, int64_t nrows, O *__restrict__ rows) {
    // one warp per row: coalesced fill of the row's range
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < nrows; r += nwarps) {
        const int64_t s = (int64_t)indptr[r], e = (int64_t)indptr[r + 1];
        for (int64_t p = s + lane; p < e; p += 32) rows[p] = (O)r;
    }
}

template <typename T>
__global__ void scatter_kernel(const T *__restrict__ data, const int64_t *__restrict__ keys, int64_t n,
                               T *__restrict__ out) {
    B2S_GRID_STRIDE(i, n) out[keys[i]] = data[i];
}

template <typename T>
__global__ void fill_kernel(T *__restrict__ out, int64_t n, T v) {
    B2S_GRID_STRIDE(i, n) out[i] = v;
}

template <typename S, typename D>
__global__ void cast_kernel(const S *__restrict__ in, D *__restrict__ out, int64_t n) {
    B2S_GRID_STRIDE(i, n) out[i] = (D)in[i];
}

// tiled dense transpose: out[c][r] = in[r][c]
template <typename T>
__global__ void transpose_kernel(const T *__restrict__ in, int64_t rows, int64_t cols, int64_t ld_in,
                                 T *__restrict__ out, int64_t ld_out) {
    __shared__ T tile[32][33];
    const int64_t bx = (int64_t)blockIdx.x * 32, by = (int64_t)blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int64_t r = by + j, c = bx + threadIdx.x;
        if (r < rows && c < cols) tile[j][threadIdx.x] = in[r * ld_in + c];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int64_t c = bx + j, r = by + threadIdx.x;
        if (r < rows && c < cols) out[c * ld_out + r] = tile[threadIdx.x][j];
    }
}

template <typename T>
__global__ void any_nan_kernel(const T *__restrict__ x, int64_t n, int32_t *__restrict__ flag) {
    bool f = false;
    B2S_GRID_STRIDE(i, n) f |= (x[i] != x[i]);
    if (__any_sync(0xffffffffu, f) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

static int make_dims(int ndim, const int64_t *v, int64_t *dst) {
    for (int d = 0; d < ndim; ++d) dst[d] = v[d];
    return 0;
}

}  // namespace b2s

using namespace b2s;

extern "C" {

int b2s_coo_linearize(int idx_bytes, int ndim, int64_t nnz, const void *coords_dev, int64_t row_stride,
                      const int64_t *strides_host, int64_t *keys_out_dev, void *stream) {
    B2S_REQUIRE(ndim >= 0 && ndim <= kMaxDims, B2S_ERR_UNSUPPORTED, "linearize: ndim %d > %d", ndim, kMaxDims);
    B2S_REQUIRE(idx_bytes == 4 || idx_bytes == 8, B2S_ERR_INVALID, "linearize: idx_bytes");
    if (nnz == 0) return B2S_OK;
    DimPack dp{};
    make_dims(ndim, strides_host, dp.stride);
    cudaStream_t s = (cudaStream_t)stream;
    if (idx_bytes == 4)
        linearize_kernel<int32_t><<<grid_for(nnz), 256, 0, s>>>(ndim, nnz, (const int32_t *)coords_dev, row_stride, dp,
                                                               keys_out_dev);
    else
        linearize_kernel<int64_t><<<grid_for(nnz), 256, 0, s>>>(ndim, nnz, (const int64_t *)coords_dev, row_stride, dp,
                                                               keys_out_dev);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

int b2s_coo_diag_flags(int idx_bytes, int ndim, int64_t nnz, const void *coords_dev, int64_t row_stride,
                        const int32_t *first_host, uint8_t *flags_out_dev, void *stream) {
    B2S_REQUIRE(ndim >= 0 && ndim <= kMaxDims, B2S_ERR_UNSUPPORTED, "diag_flags: ndim %d > %d", ndim, kMaxDims);
    B2S_REQUIRE(idx_bytes == 4 || idx_bytes == 8, B2S_ERR_INVALID, "diag_flags: idx_bytes");
    if (nnz == 0) return B2S_OK;
    DiagPack dg{};
    for (int d = 0; d < ndim; ++d) {
        B2S_REQUIRE(first_host[d] >= 0 && first_host[d] <= d, B2S_ERR_INVALID, "diag_flags: first[%d] = %d", d,
                    first_host[d]);
        dg.first[d] = first_host[d];
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (idx_bytes == 4)
        diag_flags_kernel<int32_t><<<grid_for(nnz), 256, 0, s>>>(ndim, nnz, (const int32_t *)coords_dev, row_stride, dg,
                                                                flags_out_dev);
    else
        diag_flags_kernel<int64_t><<<grid_for(nnz), 256, 0, s>>>(ndim, nnz, (const int64_t *)coords_dev, row_stride, dg,
                                                                flags_out_dev);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

int b2s_coo_slice_keys(int ndim, int64_t nnz, const int64_t *keys_dev, const int64_t *shape_host,
                       const int64_t *start_host, const int64_t *step_host, const int64_t *count_host,
                       const int64_t *out_stride_host, uint8_t *flags_out_dev, int64_t *keys_out_dev, void *stream) {
    B2S_REQUIRE(ndim >= 1 && ndim <= kMaxDims, B2S_ERR_UNSUPPORTED, "slice_keys: ndim %d not in 1..%d", ndim, kMaxDims);
    if (nnz == 0) return B2S_OK;
    SlicePack sp{};
    for (int d = 0; d < ndim; ++d) {
        B2S_REQUIRE(step_host[d] != 0, B2S_ERR_INVALID, "slice_keys: step[%d] == 0", d);
        sp.fext[d] = make_fastdiv((uint64_t)shape_host[d]);
        sp.neg[d] = step_host[d] < 0;
        sp.fstep[d] = make_fastdiv((uint64_t)(step_host[d] < 0 ? -step_host[d] : step_host[d]));
        sp.start[d] = start_host[d];
        sp.count[d] = count_host[d];
        sp.ostride[d] = out_stride_host[d];
    }
    slice_keys_kernel<<<grid_for(nnz), 256, 0, (cudaStream_t)stream>>>(ndim, nnz, keys_dev, sp, flags_out_dev,
                                                                      keys_out_dev);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

int b2s_coo_unravel(int idx_bytes, int ndim, int64_t nnz, const int64_t *keys_dev, const int64_t *shape_host,
                    void *coords_out_dev, int64_t row_stride, void *stream) {
    B2S_REQUIRE(ndim >= 0 && ndim <= kMaxDims, B2S_ERR_UNSUPPORTED, "unravel: ndim %d > %d", ndim, kMaxDims);
    B2S_REQUIRE(idx_bytes == 4 || idx_bytes == 8, B2S_ERR_INVALID, "unravel: idx_bytes");
    if (nnz == 0 || ndim == 0) return B2S_OK;
    DimPack dp{};
    for (int d = 0; d < ndim; ++d) dp.fext[d] = make_fastdiv((uint64_t)shape_host[d]);
    cudaStream_t s = (cudaStream_t)stream;
    if (idx_bytes == 4)
        unravel_kernel<int32_t><<<grid_for(nnz), 256, 0, s>>>(ndim, nnz, keys_dev, dp, (int32_t *)coords_out_dev,
                                                             row_stride);
    else
        unravel_kernel<int64_t><<<grid_for(nnz), 256, 0, s>>>(ndim, nnz, keys_dev, dp, (int64_t *)coords_out_dev,
                                                             row_stride);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

int b2s_keys_flags(const int64_t *keys_dev, int64_t n, int *unsorted_host, int *has_dups_host, void *stream) {
    *unsorted_host = 0;
    *has_dups_host = 0;
    if (n < 2) return B2S_OK;
    cudaStream_t s = (cudaStream_t)stream;
    int32_t *flags = nullptr;
    int rc = scratch_alloc((void **)&flags, 8, s);
    if (rc) return rc;
    B2S_CUDA(cudaMemsetAsync(flags, 0, 8, s));
    keys_flags_kernel<<<grid_for(n), 256, 0, s>>>(keys_dev, n, flags);
    B2S_CHECK_LAUNCH();
    int32_t h[2];
    B2S_CUDA(cudaMemcpyAsync(h, flags, 8, cudaMemcpyDeviceToHost, s));
    B2S_CUDA(cudaStreamSynchronize(s));
    scratch_free(flags, s);
    *unsorted_host = h[0];
    *has_dups_host = h[1];
    return B2S_OK;
}

int b2s_sort_keys(const int64_t *keys_in_dev, int64_t n, int key_bits, int64_t *keys_out_dev, int64_t *perm_out_dev,
                  void *stream) {
    if (n == 0) return B2S_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (key_bits <= 0 || key_bits > 64) key_bits = 64;
    const bool big = n >= 2147483647LL;  // 64-bit offsets inside CUB only when they are needed
    int64_t *iota = nullptr;
    int rc = scratch_alloc((void **)&iota, (size_t)n * 8, s);
    if (rc) return rc;
    iota_kernel<<<grid_for(n), 256, 0, s>>>(iota, n);
    B2S_CHECK_LAUNCH();
    size_t tmp_bytes = 0;
    // keys are non-negative linear indices: sort them as unsigned over the low key_bits (stable LSD radix)
    void *tmp = nullptr;
    for (int pass = 0; pass < 2; ++pass) {  // size query, then the sort
        if (big)
            B2S_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, (const uint64_t *)keys_in_dev,
                                                     (uint64_t *)keys_out_dev, iota, perm_out_dev, (int64_t)n, 0,
                                                     key_bits, s));
        else
            B2S_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, (const uint64_t *)keys_in_dev,
                                                     (uint64_t *)keys_out_dev, iota, perm_out_dev, (int)n, 0, key_bits,
                                                     s));
        if (pass == 0) {
            rc = scratch_alloc(&tmp, tmp_bytes, s);
            if (rc) return rc;
        }
    }
    count_launch(2 + (key_bits + 7) / 8);  // histogram + onesweep passes (CUB kernels inside this .so)
    scratch_free(tmp, s);
    scratch_free(iota, s);
    return B2S_OK;
}

int b2s_gather(int elem_bytes, const void *in_dev, const int64_t *perm_dev, int64_t n, void *out_dev, void *stream) {
    if (n == 0) return B2S_OK;
    cudaStream_t s = (cudaStream_t)stream;
    switch (elem_bytes) {
        case 1: gather_kernel<uint8_t><<<grid_for(n), 256, 0, s>>>((const uint8_t *)in_dev, perm_dev, n, (uint8_t *)out_dev); break;
        case 2: gather_kernel<uint16_t><<<grid_for(n), 256, 0, s>>>((const uint16_t *)in_dev, perm_dev, n, (uint16_t *)out_dev); break;
        case 4: gather_kernel<uint32_t><<<grid_for(n), 256, 0, s>>>((const uint32_t *)in_dev, perm_dev, n, (uint32_t *)out_dev); break;
        case 16: gather_kernel<U128><<<grid_for(n), 256, 0, s>>>((const U128 *)in_dev, perm_dev, n, (U128 *)out_dev); break;
        case 8: gather_kernel<uint64_t><<<grid_for(n), 256, 0, s>>>((const uint64_t *)in_dev, perm_dev, n, (uint64_t *)out_dev); break;
        default: set_error("gather: elem_bytes %d", elem_bytes); return B2S_ERR_UNSUPPORTED;
    }
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

int b2s_flag_heads(const int64_t *keys_dev, int64_t n, uint8_t *flags_out_dev, void *stream) {
    if (n == 0) return B2S_OK;
    flag_heads_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(keys_dev, n, flags_out_dev);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

int b2s_flag_not_fill(int elem_bytes, const void *data_dev, int64_t n, const void *fill_host, uint8_t *flags_out_dev,
                      void *stream) {
    if (n == 0) return B2S_OK;
    cudaStream_t s = (cudaStream_t)stream;
    switch (elem_bytes) {
        case 1: { uint8_t f; memcpy(&f, fill_host, 1); flag_not_fill_kernel<uint8_t><<<grid_for(n), 256, 0, s>>>((const uint8_t *)data_dev, n, f, flags_out_dev); break; }
        case 2: { uint16_t f; memcpy(&f, fill_host, 2); flag_not_fill_kernel<uint16_t><<<grid_for(n), 256, 0, s>>>((const uint16_t *)data_dev, n, f, flags_out_dev); break; }
        case 4: { uint32_t f; memcpy(&f, fill_host, 4); flag_not_fill_kernel<uint32_t><<<grid_for(n), 256, 0, s>>>((const uint32_t *)data_dev, n, f, flags_out_dev); break; }
        case 16: { U128 f; memcpy(&f, fill_host, 16); flag_not_fill_kernel<U128><<<grid_for(n), 256, 0, s>>>((const U128 *)data_dev, n, f, flags_out_dev); break; }
        case 8: { uint64_t f; memcpy(&f, fill_host, 8); flag_not_fill_kernel<uint64_t><<<grid_for(n), 256, 0, s>>>((const uint64_t *)data_dev, n, f, flags_out_dev); break; }
        default: set_error("flag_not_fill: elem_bytes %d", elem_bytes); return B2S_ERR_UNSUPPORTED;
    }
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

}  // extern "C"

// cub::DeviceScan::ExclusiveSum with 32-bit offsets when n fits (the common case), 64-bit ones otherwise
template <typename InT>
static int exclusive_sum_any(const InT *in, int64_t *out, int64_t n, cudaStream_t s) {
    size_t tmp_bytes = 0;
    void *tmp = nullptr;
    const bool big = n >= 2147483647LL;
    for (int pass = 0; pass < 2; ++pass) {  // size query, then the scan
        if (big) B2S_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, in, out, (int64_t)n, s));
        else B2S_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, in, out, (int)n, s));
        if (pass == 0) {
            const int rc = scratch_alloc(&tmp, tmp_bytes, s);
            if (rc) return rc;
        }
    }
    count_launch(2);
    return scratch_free(tmp, s);
}

extern "C" {

int b2s_scan_flags(const uint8_t *flags_dev, int64_t n, int64_t *pos_out_dev, int64_t *total_host, void *stream) {
    *total_host = 0;
    if (n == 0) return B2S_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int rc = exclusive_sum_any<uint8_t>(flags_dev, pos_out_dev, n, s);
    if (rc) return rc;
    int64_t last_pos = 0;
    uint8_t last_flag = 0;
    B2S_CUDA(cudaMemcpyAsync(&last_pos, pos_out_dev + (n - 1), 8, cudaMemcpyDeviceToHost, s));
    B2S_CUDA(cudaMemcpyAsync(&last_flag, flags_dev + (n - 1), 1, cudaMemcpyDeviceToHost, s));
    B2S_CUDA(cudaStreamSynchronize(s));
    *total_host = last_pos + (last_flag ? 1 : 0);
    return B2S_OK;
}

int b2s_exclusive_scan_i64(const int64_t *in_dev, int64_t n, int64_t *out_dev, void *stream) {
    if (n == 0) return B2S_OK;
    return exclusive_sum_any<int64_t>(in_dev, out_dev, n, (cudaStream_t)stream);
}

int b2s_compact(int elem_bytes, const void *in_dev, const uint8_t *flags_dev, const int64_t *pos_dev, int64_t n,
                void *out_dev, void *stream) {
    if (n == 0) return B2S_OK;
    cudaStream_t s = (cudaStream_t)stream;
    switch (elem_bytes) {
        case 1: compact_kernel<uint8_t><<<grid_for(n), 256, 0, s>>>((const uint8_t *)in_dev, flags_dev, pos_dev, n, (uint8_t *)out_dev); break;
        case 2: compact_kernel<uint16_t><<<grid_for(n), 256, 0, s>>>((const uint16_t *)in_dev, flags_dev, pos_dev, n, (uint16_t *)out_dev); break;
        case 4: compact_kernel<uint32_t><<<grid_for(n), 256, 0, s>>>((const uint32_t *)in_dev, flags_dev, pos_dev, n, (uint32_t *)out_dev); break;
        case 16: compact_kernel<U128><<<grid_for(n), 256, 0, s>>>((const U128 *)in_dev, flags_dev, pos_dev, n, (U128 *)out_dev); break;
        case 8: compact_kernel<uint64_t><<<grid_for(n), 256, 0, s>>>((const uint64_t *)in_dev, flags_dev, pos_dev, n, (uint64_t *)out_dev); break;
        default: set_error("compact: elem_bytes %d", elem_bytes); return B2S_ERR_UNSUPPORTED;
    }
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

int b2s_compact_rows(int elem_bytes, int nrows, const void *in_dev, int64_t in_stride, const uint8_t *flags_dev,
                     const int64_t *pos_dev, int64_t n, void *out_dev, int64_t out_stride, void *stream) {
    if (n == 0 || nrows == 0) return B2S_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (elem_bytes == 2)
        compact_rows_kernel<uint16_t><<<grid_for(n), 256, 0, s>>>(nrows, (const uint16_t *)in_dev, in_stride, flags_dev, pos_dev, n, (uint16_t *)out_dev, out_stride);
    else if (elem_bytes == 4)
        compact_rows_kernel<uint32_t><<<grid_for(n), 256, 0, s>>>(nrows, (const uint32_t *)in_dev, in_stride, flags_dev, pos_dev, n, (uint32_t *)out_dev, out_stride);
    else if (elem_bytes == 8)
        compact_rows_kernel<uint64_t><<<grid_for(n), 256, 0, s>>>(nrows, (const uint64_t *)in_dev, in_stride, flags_dev, pos_dev, n, (uint64_t *)out_dev, out_stride);
    else {
        set_error("compact_rows: elem_bytes %d", elem_bytes);
        return B2S_ERR_UNSUPPORTED;
    }
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

int b2s_segment_sum(int dtype, const void *data_dev, const uint8_t *heads_dev, const int64_t *pos_dev, int64_t n,
                    void *out_dev, void *stream) {
    if (n == 0) return B2S_OK;
    cudaStream_t s = (cudaStream_t)stream;
    switch (dtype) {
        case B2S_F32: segment_sum_kernel<float><<<grid_for(n), 256, 0, s>>>((const float *)data_dev, heads_dev, pos_dev, n, (float *)out_dev); break;
        case B2S_F64: segment_sum_kernel<double><<<grid_for(n), 256, 0, s>>>((const double *)data_dev, heads_dev, pos_dev, n, (double *)out_dev); break;
        case B2S_I32: segment_sum_kernel<int32_t><<<grid_for(n), 256, 0, s>>>((const int32_t *)data_dev, heads_dev, pos_dev, n, (int32_t *)out_dev); break;
        case B2S_I64: segment_sum_kernel<int64_t><<<grid_for(n), 256, 0, s>>>((const int64_t *)data_dev, heads_dev, pos_dev, n, (int64_t *)out_dev); break;
        default: set_error("segment_sum: dtype %d", dtype); return B2S_ERR_UNSUPPORTED;
    }
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

int b2s_indptr_from_sorted(int in_idx_bytes, const void *rows_dev, int64_t n, int64_t nrows, int out_idx_bytes,
                           void *indptr_out_dev, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned g = grid_for(nrows + 1);
#define B2S_IP(I, O) indptr_from_sorted_kernel<I, O><<<g, 256, 0, s>>>((const I *)rows_dev, n, nrows, (O *)indptr_out_dev)
    if (in_idx_bytes == 4 && out_idx_bytes == 4) B2S_IP(int32_t, int32_t);
    else if (in_idx_bytes == 4 && out_idx_bytes == 8) B2S_IP(int32_t, int64_t);
    else if (in_idx_bytes == 8 && out_idx_bytes == 4) B2S_IP(int64_t, int32_t);
    else if (in_idx_bytes == 8 && out_idx_bytes == 8) B2S_IP(int64_t, int64_t);
    else {
        set_error("indptr_from_sorted: idx bytes %d/%d", in_idx_bytes, out_idx_bytes);
        return B2S_ERR_INVALID;
    }
#undef B2S_IP
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

int b2s_csr_from_keys(const int64_t *keys_dev, int64_t n, int64_t nrows, int64_t ncols, int out_idx_bytes,
                      void *rows_out_dev_or_null, void *indices_out_dev, void *indptr_out_dev, void *stream) {
    B2S_REQUIRE(out_idx_bytes == 4 || out_idx_bytes == 8, B2S_ERR_INVALID, "csr_from_keys: idx bytes");
    cudaStream_t s = (cudaStream_t)stream;
    if (out_idx_bytes == 4) {
        if (n) split_keys_kernel<int32_t><<<grid_for(n), 256, 0, s>>>(keys_dev, n, ncols, (int32_t *)rows_out_dev_or_null, (int32_t *)indices_out_dev);
        if (indptr_out_dev) indptr_from_keys_kernel<int32_t><<<grid_for(nrows + 1), 256, 0, s>>>(keys_dev, n, nrows, ncols, (int32_t *)indptr_out_dev);
    } else {
        if (n) split_keys_kernel<int64_t><<<grid_for(n), 256, 0, s>>>(keys_dev, n, ncols, (int64_t *)rows_out_dev_or_null, (int64_t *)indices_out_dev);
        if (indptr_out_dev) indptr_from_keys_kernel<int64_t><<<grid_for(nrows + 1), 256, 0, s>>>(keys_dev, n, nrows, ncols, (int64_t *)indptr_out_dev);
    }
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

int b2s_rows_from_indptr(int in_idx_bytes, const void *indptr_dev, int64_t nrows, int out_idx_bytes,
                         void *rows_out_dev, void *stream) {
    if (nrows == 0) return B2S_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned g = grid_for(nrows * 32);
#define B2S_RI(I, O) rows_from_indptr_kernel<I, O><<<g, 256, 0, s>>>((const I *)indptr_dev, nrows, (O *)rows_out_dev)
    if (in_idx_bytes == 4 && out_idx_bytes == 4) B2S_RI(int32_t, int32_t);
    else if (in_idx_bytes == 4 && out_idx_bytes == 8) B2S_RI(int32_t, int64_t);
    else if (in_idx_bytes == 8 && out_idx_bytes == 4) B2S_RI(int64_t, int32_t);
    else if (in_idx_bytes == 8 && out_idx_bytes == 8) B2S_RI(int64_t, int64_t);
    else {
        set_error("rows_from_indptr: idx bytes %d/%d", in_idx_bytes, out_idx_bytes);
        return B2S_ERR_INVALID;
    }
#undef B2S_RI
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

int b2s_cast(int src_dtype, int dst_dtype, const void *in_dev, int64_t n, void *out_dev, void *stream) {
    if (n == 0) return B2S_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned g = grid_for(n);
#define B2S_C(S, D) cast_kernel<S, D><<<g, 256, 0, s>>>((const S *)in_dev, (D *)out_dev, n)
#define B2S_ROW(SCODE, S)                                                   \
    if (src_dtype == SCODE) {                                               \
        switch (dst_dtype) {                                                \
            case B2S_F32: B2S_C(S, float); break;                           \
            case B2S_F64: B2S_C(S, double); break;                          \
            case B2S_I32: B2S_C(S, int32_t); break;                         \
            case B2S_I64: B2S_C(S, int64_t); break;                         \
            case B2S_BOOL: B2S_C(S, bool); break;                           \
            case B2S_I8: B2S_C(S, int8_t); break;                           \
            case B2S_I16: B2S_C(S, int16_t); break;                         \
            case B2S_U8: B2S_C(S, uint8_t); break;                          \
            case B2S_U16: B2S_C(S, uint16_t); break;                        \
            case B2S_U32: B2S_C(S, uint32_t); break;                        \
            case B2S_U64: B2S_C(S, uint64_t); break;                        \
            default: set_error("cast: dst dtype %d", dst_dtype); return B2S_ERR_UNSUPPORTED; \
        }                                                                   \
        B2S_CHECK_LAUNCH();                                                 \
        return B2S_OK;                                                      \
    }
    B2S_ROW(B2S_F32, float)
    B2S_ROW(B2S_F64, double)
    B2S_ROW(B2S_I32, int32_t)
    B2S_ROW(B2S_I64, int64_t)
    B2S_ROW(B2S_BOOL, bool)
    // storage-only integer widths: the cast is their way into (and out of) the compute dtype matrix
    B2S_ROW(B2S_I8, int8_t)
    B2S_ROW(B2S_I16, int16_t)
    B2S_ROW(B2S_U8, uint8_t)
    B2S_ROW(B2S_U16, uint16_t)
    B2S_ROW(B2S_U32, uint32_t)
    B2S_ROW(B2S_U64, uint64_t)
#undef B2S_ROW
#undef B2S_C
    set_error("cast: src dtype %d", src_dtype);
    return B2S_ERR_UNSUPPORTED;
}

int b2s_scatter(int elem_bytes, const void *data_dev, const int64_t *keys_dev, int64_t n, void *out_dev, void *stream) {
    if (n == 0) return B2S_OK;
    cudaStream_t s = (cudaStream_t)stream;
    switch (elem_bytes) {
        case 1: scatter_kernel<uint8_t><<<grid_for(n), 256, 0, s>>>((const uint8_t *)data_dev, keys_dev, n, (uint8_t *)out_dev); break;
        case 2: scatter_kernel<uint16_t><<<grid_for(n), 256, 0, s>>>((const uint16_t *)data_dev, keys_dev, n, (uint16_t *)out_dev); break;
        case 4: scatter_kernel<uint32_t><<<grid_for(n), 256, 0, s>>>((const uint32_t *)data_dev, keys_dev, n, (uint32_t *)out_dev); break;
        case 16: scatter_kernel<U128><<<grid_for(n), 256, 0, s>>>((const U128 *)data_dev, keys_dev, n, (U128 *)out_dev); break;
        case 8: scatter_kernel<uint64_t><<<grid_for(n), 256, 0, s>>>((const uint64_t *)data_dev, keys_dev, n, (uint64_t *)out_dev); break;
        default: set_error("scatter: elem_bytes %d", elem_bytes); return B2S_ERR_UNSUPPORTED;
    }
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

int b2s_fill(int elem_bytes, void *out_dev, int64_t n, const void *value_host, void *stream) {
    if (n == 0) return B2S_OK;
    cudaStream_t s = (cudaStream_t)stream;
    switch (elem_bytes) {
        case 1: { uint8_t v; memcpy(&v, value_host, 1); fill_kernel<uint8_t><<<grid_for(n), 256, 0, s>>>((uint8_t *)out_dev, n, v); break; }
        case 2: { uint16_t v; memcpy(&v, value_host, 2); fill_kernel<uint16_t><<<grid_for(n), 256, 0, s>>>((uint16_t *)out_dev, n, v); break; }
        case 4: { uint32_t v; memcpy(&v, value_host, 4); fill_kernel<uint32_t><<<grid_for(n), 256, 0, s>>>((uint32_t *)out_dev, n, v); break; }
        case 16: { U128 v; memcpy(&v, value_host, 16); fill_kernel<U128><<<grid_for(n), 256, 0, s>>>((U128 *)out_dev, n, v); break; }
        case 8: { uint64_t v; memcpy(&v, value_host, 8); fill_kernel<uint64_t><<<grid_for(n), 256, 0, s>>>((uint64_t *)out_dev, n, v); break; }
        default: set_error("fill: elem_bytes %d", elem_bytes); return B2S_ERR_UNSUPPORTED;
    }
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

int b2s_transpose_dense(int elem_bytes, const void *in_dev, int64_t rows, int64_t cols, int64_t ld_in, void *out_dev,
                        int64_t ld_out, void *stream) {
    if (rows == 0 || cols == 0) return B2S_OK;
    cudaStream_t s = (cudaStream_t)stream;
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
    B2S_REQUIRE((rows + 31) / 32 <= 65535, B2S_ERR_OVERFLOW, "transpose: too many rows (%lld) for one launch",
                (long long)rows);
    dim3 block(32, 8);
    switch (elem_bytes) {
        case 1: transpose_kernel<uint8_t><<<grid, block, 0, s>>>((const uint8_t *)in_dev, rows, cols, ld_in, (uint8_t *)out_dev, ld_out); break;
        case 2: transpose_kernel<uint16_t><<<grid, block, 0, s>>>((const uint16_t *)in_dev, rows, cols, ld_in, (uint16_t *)out_dev, ld_out); break;
        case 4: transpose_kernel<uint32_t><<<grid, block, 0, s>>>((const uint32_t *)in_dev, rows, cols, ld_in, (uint32_t *)out_dev, ld_out); break;
        case 16: transpose_kernel<U128><<<grid, block, 0, s>>>((const U128 *)in_dev, rows, cols, ld_in, (U128 *)out_dev, ld_out); break;
        case 8: transpose_kernel<uint64_t><<<grid, block, 0, s>>>((const uint64_t *)in_dev, rows, cols, ld_in, (uint64_t *)out_dev, ld_out); break;
        default: set_error("transpose: elem_bytes %d", elem_bytes); return B2S_ERR_UNSUPPORTED;
    }
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

int b2s_any_nan(int dtype, const void *data_dev, int64_t n, int *result_host, void *stream) {
    *result_host = 0;
    if (n == 0 || (dtype != B2S_F32 && dtype != B2S_F64)) return B2S_OK;
    cudaStream_t s = (cudaStream_t)stream;
    int32_t *flag = nullptr;
    int rc = scratch_alloc((void **)&flag, 4, s);
    if (rc) return rc;
    B2S_CUDA(cudaMemsetAsync(flag, 0, 4, s));
    if (dtype == B2S_F32) any_nan_kernel<float><<<grid_for(n), 256, 0, s>>>((const float *)data_dev, n, flag);
    else any_nan_kernel<double><<<grid_for(n), 256, 0, s>>>((const double *)data_dev, n, flag);
    B2S_CHECK_LAUNCH();
    int32_t h = 0;
    B2S_CUDA(cudaMemcpyAsync(&h, flag, 4, cudaMemcpyDeviceToHost, s));
    B2S_CUDA(cudaStreamSynchronize(s));
    scratch_free(flag, s);
    *result_host = h;
    return B2S_OK;
}

}  // extern "C"
