/*
 * oracle/dot_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded restatement of the pydata/sparse numba kernels on
 * the tensordot path.  It is the parity checker for the CUDA kernels in
 * sparse_b200/csrc and the "port" CPU baseline of bench.py.  Nothing in the
 * product package (sparse_b200/) may import, link or call this file; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs do.
 *
 * Parity status: PINNED.  Every function below is checked bit-for-bit against
 * outputs of the reference itself (the tests/golden fixtures, produced by
 * tests/golden/make_golden.py which imports /root/reference in the authoring
 * container) in tests/test_oracle_golden.py.
 *
 * Each function cites the reference lines it restates
 * (paths relative to /root/reference/sparse/numba_backend/).
 *
 * Build: see oracle/Makefile.  -ffp-contract=off is REQUIRED: the reference
 * (numba, no fastmath) rounds the product and the sum separately.
 *
 * Index type on the oracle side is always int64 (np.intp); values are
 * instantiated for f32, f64, i32 and i64.  Mixed-dtype products are promoted
 * by the caller first (the reference promotes av*bv to the common dtype, which
 * is the same thing: _common.py:635-636).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t idx_t;

#if defined(_OPENMP)
#include <omp.h>
#endif

int orc_max_threads(void) {
#if defined(_OPENMP)
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* fix the number of OpenMP threads of the row-parallel loops (bench.py's "all cores" figure must not depend on the
 * launcher's OMP_NUM_THREADS: torchrun exports OMP_NUM_THREADS=1) */
void orc_set_threads(int n) {
#if defined(_OPENMP)
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------ */
/* _csr_csr_count_nnz, _common.py:543-570                              */
/* Number of structural entries of A*B (marker array per output row).  */
/* ------------------------------------------------------------------ */
idx_t orc_csr_csr_count(idx_t n_row, idx_t n_col, const idx_t *a_indices,
                        const idx_t *b_indices, const idx_t *a_indptr,
                        const idx_t *b_indptr) {
    idx_t *seen = (idx_t *)malloc(sizeof(idx_t) * (size_t)(n_col > 0 ? n_col : 1));
    for (idx_t c = 0; c < n_col; ++c) seen[c] = -1;
    idx_t total = 0;
    for (idx_t r = 0; r < n_row; ++r) {
        for (idx_t p = a_indptr[r]; p < a_indptr[r + 1]; ++p) {
            idx_t mid = a_indices[p];
            for (idx_t q = b_indptr[mid]; q < b_indptr[mid + 1]; ++q) {
                idx_t c = b_indices[q];
                if (seen[c] != r) {
                    seen[c] = r;
                    ++total;
                }
            }
        }
    }
    free(seen);
    return total;
}

/* _match_arrays, _umath.py:53-92.  Sort-merge join of two sorted key arrays;
 * emits every (ia, ib) with a[ia] == b[ib] ordered by (ia, ib).  Pass
 * out_a == NULL to count only. */
idx_t orc_match_arrays(const idx_t *a, idx_t na, const idx_t *b, idx_t nb,
                       idx_t *out_a, idx_t *out_b) {
    if (na == 0 || nb == 0) return 0;
    idx_t n = 0, ib = 0, anchor = 0;
    for (idx_t ia = 0; ia < na; ++ia) {
        idx_t key = a[ia];
        if (key == b[anchor]) ib = anchor; /* rewind for duplicate runs in a */
        while (ib < nb && key >= b[ib]) {
            if (key == b[ib]) {
                if (out_a) {
                    out_a[n] = ia;
                    out_b[n] = ib;
                }
                ++n;
                if (b[anchor] < b[ib]) anchor = ib;
            }
            ++ib;
        }
    }
    return n;
}

/* _calc_counts_invidx, _coo/core.py:1601-1628.  Run starts and run lengths of
 * a sequence of contiguous group ids.  Returns the number of runs. */
idx_t orc_counts_invidx(const idx_t *groups, idx_t n, idx_t *inv_idx, idx_t *counts) {
    if (n == 0) return 0;
    idx_t runs = 0;
    inv_idx[0] = 0;
    idx_t cur = groups[0];
    for (idx_t i = 1; i < n; ++i) {
        if (groups[i] != cur) {
            counts[runs] = i - inv_idx[runs];
            ++runs;
            inv_idx[runs] = i;
            cur = groups[i];
        }
    }
    counts[runs] = n - inv_idx[runs];
    return runs + 1;
}

/* uncompress_dimension, _compressed/convert.py:81-87: indptr -> row id per nnz */
void orc_uncompress(const idx_t *indptr, idx_t n_row, idx_t *rows) {
    for (idx_t r = 0; r < n_row; ++r)
        for (idx_t p = indptr[r]; p < indptr[r + 1]; ++p) rows[p] = r;
}

/*
 * Typed kernels.  T = value type, W = the accumulator numba infers for
 * "val = 0; val += v * b" (int64 unified with T: f32->f64, f64->f64,
 * i32->i64, i64->i64), D = np.zeros(n) default dtype (float64) used by
 * _dot_csc_ndarray_sparse's `sums` (_common.py:835).
 */
#define ORC_DEFINE(T, W, SUF)                                                                      \
                                                                                                   \
    /* _dot_csr_ndarray, _common.py:720-755: out[i,:] += a[i,k] * b[k,:], stored order, */         \
    /* product rounded to T, then sum rounded to T.                                     */         \
    void orc_csr_dense_##SUF(idx_t M, idx_t N, const T *a_data, const idx_t *a_indices,            \
                             const idx_t *a_indptr, const T *b, T *out) {                          \
        memset(out, 0, sizeof(T) * (size_t)M * (size_t)N);                                         \
        _Pragma("omp parallel for schedule(dynamic, 256)") for (idx_t i = 0; i < M; ++i) {         \
            T *row = out + (size_t)i * (size_t)N;                                                  \
            for (idx_t p = a_indptr[i]; p < a_indptr[i + 1]; ++p) {                                \
                const T v = a_data[p];                                                             \
                const T *brow = b + (size_t)a_indices[p] * (size_t)N;                              \
                for (idx_t j = 0; j < N; ++j) {                                                    \
                    T prod = v * brow[j];                                                          \
                    row[j] = row[j] + prod;                                                        \
                }                                                                                  \
            }                                                                                      \
        }                                                                                          \
    }                                                                                              \
                                                                                                   \
    /* _dot_csc_ndarray, _common.py:869-904: A given by columns; scatter into out rows. */         \
    void orc_csc_dense_##SUF(idx_t a_rows, idx_t a_cols, idx_t N, const T *a_data,                 \
                             const idx_t *a_indices, const idx_t *a_indptr, const T *b, T *out) {  \
        memset(out, 0, sizeof(T) * (size_t)a_rows * (size_t)N);                                    \
        for (idx_t c = 0; c < a_cols; ++c) {                                                       \
            const T *brow = b + (size_t)c * (size_t)N;                                             \
            for (idx_t p = a_indptr[c]; p < a_indptr[c + 1]; ++p) {                                \
                const T v = a_data[p];                                                             \
                T *row = out + (size_t)a_indices[p] * (size_t)N;                                   \
                for (idx_t j = 0; j < N; ++j) {                                                    \
                    T prod = v * brow[j];                                                          \
                    row[j] = row[j] + prod;                                                        \
                }                                                                                  \
            }                                                                                      \
        }                                                                                          \
    }                                                                                              \
                                                                                                   \
    /* _dot_csr_csr, _common.py:639-717 (and _dot_coo_coo :907-976 with coo_rows != NULL).  */     \
    /* Gustavson with an intrusive linked list: columns of a row come out in REVERSE order  */     \
    /* of first touch.  dense_flip reproduces :709-714 (CSR variant only).                  */     \
    void orc_csr_csr_##SUF(idx_t n_row, idx_t n_col, const T *a_data, const T *b_data,             \
                           const idx_t *a_indices, const idx_t *b_indices, const idx_t *a_indptr,  \
                           const idx_t *b_indptr, T *data, idx_t *indices, idx_t *indptr,          \
                           idx_t *coo_rows, int dense_flip) {                                      \
        idx_t *link = (idx_t *)malloc(sizeof(idx_t) * (size_t)(n_col > 0 ? n_col : 1));            \
        T *acc = (T *)calloc((size_t)(n_col > 0 ? n_col : 1), sizeof(T));                          \
        for (idx_t c = 0; c < n_col; ++c) link[c] = -1;                                            \
        idx_t w = 0;                                                                               \
        if (indptr) indptr[0] = 0;                                                                 \
        for (idx_t r = 0; r < n_row; ++r) {                                                        \
            idx_t head = -2, len = 0;                                                              \
            for (idx_t p = a_indptr[r]; p < a_indptr[r + 1]; ++p) {                                \
                const idx_t mid = a_indices[p];                                                    \
                const T av = a_data[p];                                                            \
                for (idx_t q = b_indptr[mid]; q < b_indptr[mid + 1]; ++q) {                        \
                    const idx_t c = b_indices[q];                                                  \
                    T prod = av * b_data[q];                                                       \
                    acc[c] = acc[c] + prod;                                                        \
                    if (link[c] == -1) {                                                           \
                        link[c] = head;                                                            \
                        head = c;                                                                  \
                        ++len;                                                                     \
                    }                                                                              \
                }                                                                                  \
            }                                                                                      \
            for (idx_t t = 0; t < len; ++t) {                                                      \
                /* the reference's `if next_[head] != -1` (:696) is always true */                 \
                indices[w] = head;                                                                 \
                if (coo_rows) coo_rows[w] = r;                                                     \
                data[w] = acc[head];                                                               \
                ++w;                                                                               \
                idx_t nxt = link[head];                                                            \
                link[head] = -1;                                                                   \
                acc[head] = 0;                                                                     \
                head = nxt;                                                                        \
            }                                                                                      \
            if (indptr) indptr[r + 1] = w;                                                         \
        }                                                                                          \
        if (dense_flip && n_col > 0 && w == n_col * n_row) {                                       \
            for (idx_t r = 0; r < n_row; ++r) {                                                    \
                idx_t lo = r * n_col, hi = lo + n_col - 1;                                         \
                while (lo < hi) {                                                                  \
                    T td = data[lo];                                                               \
                    data[lo] = data[hi];                                                           \
                    data[hi] = td;                                                                 \
                    idx_t ti = indices[lo];                                                        \
                    indices[lo] = indices[hi];                                                     \
                    indices[hi] = ti;                                                              \
                    ++lo;                                                                          \
                    --hi;                                                                          \
                }                                                                                  \
            }                                                                                      \
        }                                                                                          \
        free(link);                                                                                \
        free(acc);                                                                                 \
    }                                                                                              \
                                                                                                   \
    /* _csr_ndarray_count_nnz, _common.py:573-600: entry (i,j) exists iff some stored k  */        \
    /* of row i has b[k,j] != 0 (structural test on B).  Fills indptr, returns nnz.      */        \
    idx_t orc_csr_dense_sparse_count_##SUF(idx_t M, idx_t N, const idx_t *a_indices,               \
                                           const idx_t *a_indptr, const T *b, idx_t *indptr) {     \
        idx_t total = 0;                                                                           \
        indptr[0] = 0;                                                                             \
        for (idx_t i = 0; i < M; ++i) {                                                            \
            for (idx_t j = 0; j < N; ++j) {                                                        \
                for (idx_t p = a_indptr[i]; p < a_indptr[i + 1]; ++p) {                            \
                    if (b[(size_t)a_indices[p] * (size_t)N + (size_t)j] != 0) {                    \
                        ++total;                                                                   \
                        break;                                                                     \
                    }                                                                              \
                }                                                                                  \
            }                                                                                      \
            indptr[i + 1] = total;                                                                 \
        }                                                                                          \
        return total;                                                                              \
    }                                                                                              \
                                                                                                   \
    /* _dot_csr_ndarray_sparse, _common.py:758-804.  The accumulator is numba's unified */         \
    /* type of `val = 0` and `v*b` (W); the product itself is rounded to T first.       */         \
    void orc_csr_dense_sparse_fill_##SUF(idx_t M, idx_t N, const T *a_data,                        \
                                         const idx_t *a_indices, const idx_t *a_indptr,            \
                                         const T *b, T *data, idx_t *indices) {                    \
        idx_t w = 0;                                                                               \
        for (idx_t i = 0; i < M; ++i) {                                                            \
            for (idx_t j = 0; j < N; ++j) {                                                        \
                W val = 0;                                                                         \
                int any = 0;                                                                       \
                for (idx_t p = a_indptr[i]; p < a_indptr[i + 1]; ++p) {                            \
                    const T bv = b[(size_t)a_indices[p] * (size_t)N + (size_t)j];                  \
                    T prod = a_data[p] * bv;                                                       \
                    val = val + (W)prod;                                                           \
                    if (bv != 0) any = 1;                                                          \
                }                                                                                  \
                if (any) {                                                                         \
                    data[w] = (T)val;                                                              \
                    indices[w] = j;                                                                \
                    ++w;                                                                           \
                }                                                                                  \
            }                                                                                      \
        }                                                                                          \
    }                                                                                              \
                                                                                                   \
    /* _csc_ndarray_count_nnz, _common.py:603-632.  A by columns (a_cols of them),      */         \
    /* b is (a_cols x N); output is compressed by COLUMN of the product.                */         \
    idx_t orc_csc_dense_sparse_count_##SUF(idx_t a_rows, idx_t a_cols, idx_t N,                    \
                                           const idx_t *a_indices, const idx_t *a_indptr,          \
                                           const T *b, idx_t *indptr) {                            \
        idx_t *seen = (idx_t *)malloc(sizeof(idx_t) * (size_t)(a_rows > 0 ? a_rows : 1));          \
        for (idx_t r = 0; r < a_rows; ++r) seen[r] = -1;                                           \
        idx_t total = 0;                                                                           \
        for (idx_t j = 0; j < N; ++j) {                                                            \
            for (idx_t c = 0; c < a_cols; ++c) {                                                   \
                if (b[(size_t)c * (size_t)N + (size_t)j] == 0) continue;                           \
                for (idx_t p = a_indptr[c]; p < a_indptr[c + 1]; ++p) {                            \
                    idx_t r = a_indices[p];                                                        \
                    if (seen[r] != j) {                                                            \
                        seen[r] = j;                                                               \
                        ++total;                                                                   \
                    }                                                                              \
                }                                                                                  \
            }                                                                                      \
            indptr[j + 1] = total;                                                                 \
        }                                                                                          \
        free(seen);                                                                                \
        return total;                                                                              \
    }                                                                                              \
                                                                                                   \
    /* _dot_csc_ndarray_sparse, _common.py:807-866.  `sums` is float64 whatever T is    */         \
    /* (:835); rows come out in reverse first-touch order; entries whose sum == 0 are   */         \
    /* skipped (so fewer than the counted nnz may be written: returns the number).      */         \
    idx_t orc_csc_dense_sparse_fill_##SUF(idx_t a_rows, idx_t a_cols, idx_t N, const T *a_data,    \
                                          const idx_t *a_indices, const idx_t *a_indptr,           \
                                          const T *b, T *data, idx_t *indices) {                   \
        double *acc = (double *)calloc((size_t)(a_rows > 0 ? a_rows : 1), sizeof(double));         \
        idx_t *link = (idx_t *)malloc(sizeof(idx_t) * (size_t)(a_rows > 0 ? a_rows : 1));          \
        for (idx_t r = 0; r < a_rows; ++r) link[r] = -1;                                           \
        idx_t w = 0;                                                                               \
        for (idx_t j = 0; j < N; ++j) {                                                            \
            idx_t head = -2, len = 0;                                                              \
            for (idx_t c = 0; c < a_cols; ++c) {                                                   \
                const T u = b[(size_t)c * (size_t)N + (size_t)j];                                  \
                if (u == 0) continue;                                                              \
                for (idx_t p = a_indptr[c]; p < a_indptr[c + 1]; ++p) {                            \
                    const idx_t r = a_indices[p];                                                  \
                    T prod = u * a_data[p];                                                        \
                    acc[r] = acc[r] + (double)prod;                                                \
                    if (link[r] == -1) {                                                           \
                        link[r] = head;                                                            \
                        head = r;                                                                  \
                        ++len;                                                                     \
                    }                                                                              \
                }                                                                                  \
            }                                                                                      \
            for (idx_t t = 0; t < len; ++t) {                                                      \
                if (acc[head] != 0) {                                                              \
                    indices[w] = head;                                                             \
                    data[w] = (T)acc[head];                                                        \
                    ++w;                                                                           \
                }                                                                                  \
                idx_t nxt = link[head];                                                            \
                link[head] = -1;                                                                   \
                acc[head] = 0;                                                                     \
                head = nxt;                                                                        \
            }                                                                                      \
        }                                                                                          \
        free(acc);                                                                                 \
        free(link);                                                                                \
        return w;                                                                                  \
    }                                                                                              \
                                                                                                   \
    /* _dot_coo_ndarray (dense out), _common.py:979-1014.  b2t is b transposed,         */         \
    /* shape (N, K) with row stride ldb (the reference passes the b.T *view*).          */         \
    void orc_coo_dense_##SUF(idx_t nnz, const idx_t *rows, const idx_t *cols, const T *a_data,     \
                             const T *b2t, idx_t ldb_row, idx_t ldb_col, idx_t M, idx_t N,         \
                             T *out) {                                                             \
        memset(out, 0, sizeof(T) * (size_t)M * (size_t)N);                                         \
        idx_t p = 0;                                                                               \
        while (p < nnz) {                                                                          \
            const idx_t r = rows[p];                                                               \
            const idx_t start = p;                                                                 \
            for (idx_t j = 0; j < N; ++j) {                                                        \
                p = start;                                                                         \
                while (p < nnz && rows[p] == r) {                                                  \
                    T prod = a_data[p] * b2t[j * ldb_row + cols[p] * ldb_col];                     \
                    out[(size_t)r * (size_t)N + (size_t)j] += prod;                                \
                    ++p;                                                                           \
                }                                                                                  \
            }                                                                                      \
            if (N == 0) {                                                                          \
                while (p < nnz && rows[p] == r) ++p;                                               \
            }                                                                                      \
        }                                                                                          \
    }                                                                                              \
                                                                                                   \
    /* _dot_coo_ndarray (sparse out), _common.py:1017-1072: accumulator typed T          */        \
    /* (locals data_curr), entries with sum == 0 skipped.  Returns count; pass           */        \
    /* out_data == NULL to count only.                                                   */        \
    idx_t orc_coo_dense_sparse_##SUF(idx_t nnz, const idx_t *rows, const idx_t *cols,              \
                                     const T *a_data, const T *b2t, idx_t ldb_row, idx_t ldb_col,  \
                                     idx_t N, idx_t *out_rows, idx_t *out_cols, T *out_data) {     \
        idx_t w = 0, p = 0;                                                                        \
        while (p < nnz) {                                                                          \
            const idx_t r = rows[p];                                                               \
            idx_t q = p;                                                                           \
            for (idx_t j = 0; j < N; ++j) {                                                        \
                q = p;                                                                             \
                T s = 0;                                                                           \
                while (q < nnz && rows[q] == r) {                                                  \
                    T prod = a_data[q] * b2t[j * ldb_row + cols[q] * ldb_col];                     \
                    s = s + prod;                                                                  \
                    ++q;                                                                           \
                }                                                                                  \
                if (s != 0) {                                                                      \
                    if (out_data) {                                                                \
                        out_rows[w] = r;                                                           \
                        out_cols[w] = j;                                                           \
                        out_data[w] = s;                                                           \
                    }                                                                              \
                    ++w;                                                                           \
                }                                                                                  \
            }                                                                                      \
            if (N == 0) {                                                                          \
                while (q < nnz && rows[q] == r) ++q;                                               \
            }                                                                                      \
            p = q;                                                                                 \
        }                                                                                          \
        return w;                                                                                  \
    }                                                                                              \
                                                                                                   \
    /* _dot_ndarray_coo (dense out), _common.py:1075-1103: out[i, col] += a[i,row]*v     */        \
    /* for every stored (row, col, v) of b, in stored order.                             */        \
    void orc_dense_coo_##SUF(idx_t M, idx_t K, idx_t N, const T *a, idx_t nnz,                     \
                             const idx_t *b_rows, const idx_t *b_cols, const T *b_data, T *out) {  \
        memset(out, 0, sizeof(T) * (size_t)M * (size_t)N);                                         \
        for (idx_t i = 0; i < M; ++i) {                                                            \
            for (idx_t p = 0; p < nnz; ++p) {                                                      \
                T prod = a[(size_t)i * (size_t)K + (size_t)b_rows[p]] * b_data[p];                 \
                out[(size_t)i * (size_t)N + (size_t)b_cols[p]] += prod;                            \
            }                                                                                      \
        }                                                                                          \
    }                                                                                              \
                                                                                                   \
    /* _dot_ndarray_coo (sparse out), _common.py:1106-1158.  t_first/t_second are the    */        \
    /* coords of b.T (sorted by b's column): t_first = output column, t_second = k.      */        \
    /* Restated with its quirk: the running column starts at 0 and a flush happens only  */        \
    /* on a column change or at the end of the scan, and only if the sum != 0.           */        \
    idx_t orc_dense_coo_sparse_##SUF(idx_t M, idx_t K, const T *a, idx_t nnz,                      \
                                     const idx_t *t_first, const idx_t *t_second,                  \
                                     const T *t_data, idx_t *out_rows, idx_t *out_cols,            \
                                     T *out_data) {                                                \
        idx_t w = 0;                                                                               \
        for (idx_t i = 0; i < M; ++i) {                                                            \
            T s = 0;                                                                               \
            idx_t cur = 0;                                                                         \
            for (idx_t p = 0; p < nnz; ++p) {                                                      \
                if (t_first[p] != cur) {                                                           \
                    if (s != 0) {                                                                  \
                        if (out_data) {                                                            \
                            out_rows[w] = i;                                                       \
                            out_cols[w] = cur;                                                     \
                            out_data[w] = s;                                                       \
                        }                                                                          \
                        ++w;                                                                       \
                        s = 0;                                                                     \
                    }                                                                              \
                    cur = t_first[p];                                                              \
                }                                                                                  \
                T prod = a[(size_t)i * (size_t)K + (size_t)t_second[p]] * t_data[p];               \
                s = s + prod;                                                                      \
            }                                                                                      \
            if (s != 0) {                                                                          \
                if (out_data) {                                                                    \
                    out_rows[w] = i;                                                               \
                    out_cols[w] = cur;                                                             \
                    out_data[w] = s;                                                               \
                }                                                                                  \
                ++w;                                                                               \
            }                                                                                      \
        }                                                                                          \
        return w;                                                                                  \
    }

ORC_DEFINE(float, double, f32)
ORC_DEFINE(double, double, f64)
ORC_DEFINE(int32_t, int64_t, i32)
ORC_DEFINE(int64_t, int64_t, i64)
