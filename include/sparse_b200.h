/*
 * sparse_b200.h -- C ABI of libsparse_b200.so, the B200 (sm_100a) implementation
 * of pydata/sparse's data-parallel hot path.
 *
 * The reference has no C/FFI seam for this path: its "native" layer is a set of
 * numba-JIT closures called only from sparse/numba_backend/_common.py::_dot,
 * _umath.py::_Elemwise and _coo/core.py::_grouped_reduce.  Each entry point below
 * replaces one of those closures; the reference call site it substitutes is cited
 * on every declaration (paths relative to /root/reference/sparse/numba_backend/).
 *
 * Conventions
 *  - extern "C", plain pointers and sizes; no torch / Python / C++ types.
 *  - every function returns 0 on success, a negative b2s_status otherwise;
 *    b2s_last_error() returns a thread-local message for the last failure.
 *  - "_dev" pointers are device (HBM) pointers on the current CUDA device,
 *    "_host" pointers are host pointers.  `stream` is a cudaStream_t passed as
 *    void* (NULL = legacy default stream).  Calls are asynchronous on `stream`
 *    unless stated otherwise; functions that return a data-dependent size
 *    synchronise the stream once.
 *  - inputs are borrowed and never written; outputs are caller-allocated
 *    (the two-phase count -> fill pattern of the reference's own kernels,
 *    e.g. _csr_csr_count_nnz + _dot_csr_csr).
 *  - index arrays are int32 or int64 (`idx_bytes` = 4 or 8); value arrays are
 *    described by b2s_dtype.  Mixed-dtype products are promoted by the caller
 *    (the reference promotes to _dot_dtype(dt1, dt2), _common.py:635-636).
 *  - there is NO CPU fallback anywhere in this library.
 */
#ifndef SPARSE_B200_H_
#define SPARSE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2S_ABI_VERSION 1

typedef enum {
    B2S_OK = 0,
    B2S_ERR_INVALID = -1,     /* bad argument (dtype, alignment, negative size ...) */
    B2S_ERR_CUDA = -2,        /* a CUDA runtime call failed; see b2s_last_error()   */
    B2S_ERR_UNSUPPORTED = -3, /* dtype / op outside the supported matrix            */
    B2S_ERR_OVERFLOW = -4,    /* sizes do not fit the chosen index width            */
    B2S_ERR_NO_DEVICE = -5    /* no CUDA device visible                             */
} b2s_status;

typedef enum { B2S_F32 = 0, B2S_F64 = 1, B2S_I32 = 2, B2S_I64 = 3, B2S_BOOL = 4 } b2s_dtype;

/* ---- runtime ---------------------------------------------------------- */
int b2s_abi_version(void);
const char *b2s_last_error(void);
int b2s_device_count(int *count);
/* name_buf may be NULL. sm = 10*major+minor (100 on B200). */
int b2s_device_info(int device, char *name_buf, size_t name_len, int *sm, int *n_sms, size_t *hbm_bytes,
                    size_t *l2_bytes);
int b2s_set_device(int device);
int b2s_malloc(void **dev_ptr, size_t nbytes);
int b2s_free(void *dev_ptr);
int b2s_host_register(void *host_ptr, size_t nbytes); /* pin caller memory so H2D/D2H run at link rate */
int b2s_host_unregister(void *host_ptr);
int b2s_memcpy_h2d(void *dst_dev, const void *src_host, size_t nbytes, void *stream);
int b2s_memcpy_d2h(void *dst_host, const void *src_dev, size_t nbytes, void *stream);
int b2s_memset(void *dst_dev, int value, size_t nbytes, void *stream);
int b2s_stream_sync(void *stream);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
int64_t b2s_launch_count(void);

/* ---- CSR x dense -> dense (K1) ---------------------------------------- */
/*
 * Replaces _dot_csr_ndarray_type(dt1, dt2)(out_shape, a_data, a_indices, a_indptr, b)
 * (_common.py:720-755, called at _common.py:389 and :440): out[M,N] = CSR(A)[M,K] . B[K,N].
 * Bit-exact contract: every out[i,j] is accumulated in dtype arithmetic, in the
 * stored order of row i, product and sum rounded separately (no FMA), starting
 * from +0 -- the same operation sequence as the reference loop.
 * ldb / ldc are row strides in ELEMENTS (>= N).  Rows without entries are
 * written as zeros (out needs no pre-initialisation).
 */
int b2s_spmm_csr_dense(int dtype, int idx_bytes, int64_t M, int64_t K, int64_t N, const void *a_data_dev,
                       const void *a_indices_dev, const void *a_indptr_dev, const void *b_dev, int64_t ldb,
                       void *out_dev, int64_t ldc, void *stream);

/*
 * Same product through HOST buffers (the reference-facing call: numpy arrays in,
 * numpy array out).  Uploads A and B, runs K1, downloads out; synchronous.
 * Host buffers should be pinned (b2s_host_register) for full link rate.
 * a_indices/a_indptr are int64 on the host (np.intp, as the reference holds
 * them); they are narrowed to int32 on the device when M, K and nnz allow.
 */
int b2s_spmm_csr_dense_host(int dtype, int64_t M, int64_t K, int64_t N, int64_t nnz, const void *a_data_host,
                            const int64_t *a_indices_host, const int64_t *a_indptr_host, const void *b_host,
                            void *out_host);

/* Tuning knob for K1 (0 = default).  variant: 1 = register-staged LDG gather,
 * 2 = 1-D bulk-TMA (cp.async.bulk) gather through a shared-memory ring. */
int b2s_spmm_set_variant(int variant, int unroll);

#ifdef __cplusplus
}
#endif
#endif /* SPARSE_B200_H_ */
