// spgemm.cu -- K4: CSR x CSR -> CSR (and COO x COO -> COO), Gustavson row products with an
// on-chip hash accumulator, reproducing the reference's results bit-for-bit:
//   * every output value is summed in the reference's visiting order (A row entries in stored
//     order, then B row entries in stored order), product and sum rounded separately;
//   * REF order: the columns of an output row come out in REVERSE order of first touch, exactly
//     like the intrusive linked list of _dot_csr_csr (sparse/numba_backend/_common.py:639-717),
//     including the row reversal when the result is completely dense (:709-714);
//   * SORTED order: ascending columns (the canonical COO order that _dot_coo_coo's result gets from
//     the COO constructor, _common.py:907-976 + :462-469), so no global sort is needed afterwards.
//
// Replaces _csr_csr_count_nnz (:543-570), _dot_csr_csr (:639-717), _dot_coo_coo (:907-976).
//
// Structure (one numeric pass, no separate symbolic pass):
//   1. row_products: P_i = sum over A-row entries of the B-row lengths; U_i = min(P_i, n_col) bounds nnz_i.
//   2. rows are binned by P_i: warp-per-row with a shared-memory hash of 128 or 512 slots (P_i <= 64 / 256),
//      CTA-per-row with a global-memory hash for longer rows.
//   3. numeric kernels stage the row's products in visiting order, insert them in that order
//      (__match_any_sync groups equal columns inside a 32-chunk so adds stay sequential), remember each
//      column's first-touch sequence number, and write (col, sum) at offset ub_off[i] + rank, where rank comes
//      from a bitmap over sequence numbers (REF) or a counting rank over columns (SORTED).
//   4. finish: per-row compaction from the upper-bound layout to the final CSR/COO arrays, optionally
//      dropping values bitwise equal to +0 (the prune=True of _common.py:374-379) and reversing rows in the
//      all-dense case.
#include <cub/cub.cuh>
#include <type_traits>

#include "common.cuh"

namespace b2s {

constexpr unsigned FULL = 0xffffffffu;

template <typename I>
__device__ __forceinline__ unsigned hash_col(I k) {
    uint64_t x = (uint64_t)k;
    x *= 0x9E3779B97F4A7C15ull;
    return (unsigned)(x >> 32);
}

template <typename I>
struct Empty {
    static constexpr I value = (I)-1;
};

__device__ __forceinline__ int32_t cas_key(int32_t *p, int32_t cmp, int32_t val) {
    return (int32_t)atomicCAS((int *)p, (int)cmp, (int)val);
}
__device__ __forceinline__ int64_t cas_key(int64_t *p, int64_t cmp, int64_t val) {
    return (int64_t)atomicCAS((unsigned long long *)p, (unsigned long long)cmp, (unsigned long long)val);
}

template <typename T>
__device__ __forceinline__ bool is_pos_zero_bits(T v) {
    if constexpr (sizeof(T) == 4) {
        uint32_t u;
        memcpy(&u, &v, 4);
        return u == 0u;
    } else {
        uint64_t u;
        memcpy(&u, &v, 8);
        return u == 0ull;
    }
}

// Accumulator -> stored value.  W == T: identity.  W == double with T != double is the float64 `sums`
// array of _dot_csc_ndarray_sparse (_common.py:835): entries whose wide sum == 0 are skipped there
// (:852), which we encode as +0 so that the pruning compaction drops them.
template <typename T, typename W>
__device__ __forceinline__ T narrow_sum(W s) {
    if constexpr (std::is_same<T, W>::value) {
        return s;
    } else {
        if (s == W(0)) return T(0);
        return (T)s;
    }
}

// ---------------------------------------------------------------------------------------------
// 1. products per row
// ---------------------------------------------------------------------------------------------
template <typename I>
__global__ void row_products_kernel(int64_t M, int64_t n_col, const I *__restrict__ a_indptr,
                                    const I *__restrict__ a_indices, const I *__restrict__ b_indptr,
                                    int64_t *__restrict__ P, int64_t *__restrict__ U) {
    // 8 lanes per row: A rows are short in the common case
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t row = gid >> 3;
    const int sub = threadIdx.x & 7;
    int64_t acc = 0;
    if (row < M) {
        const int64_t s = (int64_t)a_indptr[row], e = (int64_t)a_indptr[row + 1];
        for (int64_t p = s + sub; p < e; p += 8) {
            const I j = a_indices[p];
            acc += (int64_t)b_indptr[j + 1] - (int64_t)b_indptr[j];
        }
    }
    acc += __shfl_xor_sync(FULL, acc, 1);
    acc += __shfl_xor_sync(FULL, acc, 2);
    acc += __shfl_xor_sync(FULL, acc, 4);
    if (row < M && sub == 0) {
        P[row] = acc;
        U[row] = acc < n_col ? acc : n_col;
    }
}

// classify rows into bins (atomic append; order inside a bin is irrelevant).  lists = 4 arrays of M entries.
__global__ void bin_rows_kernel(int64_t M, const int64_t *__restrict__ P, int64_t t0, int64_t t1, int64_t t2,
                                int64_t *__restrict__ lists, unsigned long long *__restrict__ counts,
                                unsigned long long *__restrict__ maxP) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= M) return;
    const int64_t p = P[row];
    if (p == 0) return;
    const int bin = p <= t0 ? 0 : (p <= t1 ? 1 : (p <= t2 ? 2 : 3));
    lists[(int64_t)bin * M + atomicAdd(counts + bin, 1ull)] = row;
    if (bin == 3) atomicMax(maxP, (unsigned long long)p);
}

// ---------------------------------------------------------------------------------------------
// 3a. warp-per-row numeric kernel, shared-memory hash of H slots (rows with P <= H/2)
// ---------------------------------------------------------------------------------------------
template <typename T, typename W, typename I, int H, int WARPS, bool SORTED>
__global__ void __launch_bounds__(WARPS * 32)
spgemm_warp_kernel(const int64_t *__restrict__ rows, int64_t n_rows, const I *__restrict__ a_indptr,
                   const I *__restrict__ a_indices, const T *__restrict__ a_data, const I *__restrict__ b_indptr,
                   const I *__restrict__ b_indices, const T *__restrict__ b_data,
                   const int64_t *__restrict__ ub_off, int64_t *__restrict__ tmp_idx, T *__restrict__ tmp_val,
                   int64_t *__restrict__ row_nnz, int64_t *__restrict__ row_nz) {
    constexpr int PMAX = H / 2;
    constexpr int WORDS = (PMAX + 31) / 32;
    static_assert(WORDS <= 32, "bitmap must fit one word per lane");
    constexpr I EMPTY = Empty<I>::value;
    __shared__ I st_key[WARPS][PMAX];
    __shared__ T st_val[WARPS][PMAX];
    __shared__ I tb_key[WARPS][H];
    __shared__ int tb_seq[WARPS][H];
    __shared__ W tb_sum[WARPS][H];
    __shared__ int s_off[WARPS][32];
    __shared__ int64_t s_bs[WARPS][32];
    __shared__ T s_av[WARPS][32];
    __shared__ unsigned s_bits[WARPS][32];

    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    const int64_t warps_total = (int64_t)gridDim.x * WARPS;
    for (int64_t ri = (int64_t)blockIdx.x * WARPS + w; ri < n_rows; ri += warps_total) {
        const int64_t row = rows[ri];
        for (int s = lane; s < H; s += 32) tb_key[w][s] = EMPTY;
        s_bits[w][lane] = 0u;
        __syncwarp();

        // ---- stage 1: expand the row's products, in visiting order, into shared memory ----------
        const int64_t as = (int64_t)a_indptr[row], ae = (int64_t)a_indptr[row + 1];
        int P = 0;
        for (int64_t ab = as; ab < ae; ab += 32) {
            const bool live = ab + lane < ae;
            I j = 0;
            T av = T(0);
            int64_t bs = 0;
            int len = 0;
            if (live) {
                j = a_indices[ab + lane];
                av = a_data[ab + lane];
                bs = (int64_t)b_indptr[j];
                len = (int)((int64_t)b_indptr[j + 1] - bs);
            }
            int incl = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += v;
            }
            const int total = __shfl_sync(FULL, incl, 31);
            s_off[w][lane] = incl - len;
            s_bs[w][lane] = bs;
            s_av[w][lane] = av;
            __syncwarp();
            for (int t = lane; t < total; t += 32) {
                int lo = 0;
#pragma unroll
                for (int step = 16; step > 0; step >>= 1)
                    if (s_off[w][lo + step] <= t) lo += step;
                const int q = t - s_off[w][lo];
                const int64_t src = s_bs[w][lo] + q;
                st_key[w][P + t] = b_indices[src];
                st_val[w][P + t] = mul_rn(s_av[w][lo], b_data[src]);
            }
            P += total;
            __syncwarp();
        }

        // ---- stage 2: insert in visiting order; equal columns inside a chunk stay sequential -----
        int distinct = 0;
        for (int c = 0; c < P; c += 32) {
            const int t = c + lane;
            const bool active = t < P;
            const unsigned amask = __ballot_sync(FULL, active);
            bool isnew = false;
            if (active) {
                const I k = st_key[w][t];
                const T p = st_val[w][t];
                const unsigned grp = __match_any_sync(amask, k);
                const int leader = __ffs(grp) - 1;
                const int gsz = __popc(grp);
                const bool lead = lane == leader;
                int slot = 0;
                if (lead) {
                    unsigned h = hash_col<I>(k) & (H - 1);
                    while (true) {
                        const I cur = tb_key[w][h];
                        if (cur == k) break;
                        if (cur == EMPTY) {
                            const I old = cas_key(&tb_key[w][h], EMPTY, k);
                            if (old == EMPTY) {
                                isnew = true;
                                break;
                            }
                            if (old == k) break;
                        }
                        h = (h + 1) & (H - 1);
                    }
                    slot = (int)h;
                    if (isnew) {
                        tb_seq[w][slot] = t;
                        tb_sum[w][slot] = W(0);
                        atomicOr(&s_bits[w][t >> 5], 1u << (t & 31));
                    }
                }
                const int maxg = __reduce_max_sync(amask, gsz);
                W s = lead ? tb_sum[w][slot] : W(0);
                for (int r = 0; r < maxg; ++r) {
                    const bool take = lead && r < gsz;
                    const int src = take ? (int)__fns(grp, 0, r + 1) : lane;
                    const T v = __shfl_sync(amask, p, src);
                    if (take) s = add_rn(s, (W)v);
                }
                if (lead) tb_sum[w][slot] = s;
            }
            distinct += __popc(__ballot_sync(FULL, isnew));
            __syncwarp();
        }

        // ---- stage 3: order the row and write it at its upper-bound offset -----------------------
        const int64_t ub = ub_off[row];
        int nz = 0;
        if constexpr (!SORTED) {
            // rank = number of first touches with a larger sequence number (reverse first-touch order)
            const unsigned word = s_bits[w][lane];
            int suf = (lane < WORDS) ? __popc(word) : 0;  // inclusive suffix count over words >= lane
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_down_sync(FULL, suf, o);
                if (lane + o < 32) suf += v;
            }
            s_off[w][lane] = suf;
            __syncwarp();
            for (int s = lane; s < H; s += 32) {
                const I k = tb_key[w][s];
                if (k != EMPTY) {
                    const int t = tb_seq[w][s];
                    const int wd = t >> 5;
                    const int above = (wd + 1 < 32 ? s_off[w][wd + 1] : 0) + __popc((s_bits[w][wd] >> (t & 31)) >> 1);
                    const T v = narrow_sum<T, W>(tb_sum[w][s]);
                    tmp_idx[ub + above] = (int64_t)k;
                    tmp_val[ub + above] = v;
                    nz += is_pos_zero_bits(v) ? 0 : 1;
                }
            }
        } else {
            // compact the table into the (now free) staging area, then counting rank by column
            int base = 0;
            for (int s0 = 0; s0 < H; s0 += 32) {
                const int s = s0 + lane;
                const I k = tb_key[w][s];
                const bool occ = k != EMPTY;
                const unsigned m = __ballot_sync(FULL, occ);
                if (occ) {
                    const int dst = base + __popc(m & ((1u << lane) - 1));
                    st_key[w][dst] = k;
                    st_val[w][dst] = narrow_sum<T, W>(tb_sum[w][s]);
                }
                base += __popc(m);
            }
            __syncwarp();
            for (int e = lane; e < distinct; e += 32) {
                const I k = st_key[w][e];
                int rank = 0;
                for (int f = 0; f < distinct; ++f) rank += (st_key[w][f] < k) ? 1 : 0;
                const T v = st_val[w][e];
                tmp_idx[ub + rank] = (int64_t)k;
                tmp_val[ub + rank] = v;
                nz += is_pos_zero_bits(v) ? 0 : 1;
            }
        }
        nz = __reduce_add_sync(FULL, nz);
        if (lane == 0) {
            row_nnz[row] = distinct;
            row_nz[row] = nz;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// 3b. CTA-per-row numeric kernel for long rows: hash table and sequence bitmap in global scratch.
// Exact order across warps by COLUMN OWNERSHIP: warp w only handles columns with hash % WARPS == w and
// scans every staged tile in visiting order, so each column's adds stay sequential.
// ---------------------------------------------------------------------------------------------
template <typename T, typename W, typename I, int WARPS, int TILE>
__global__ void __launch_bounds__(WARPS * 32)
spgemm_block_kernel(const int64_t *__restrict__ rows, int64_t n_rows, int64_t n_col,
                    const I *__restrict__ a_indptr, const I *__restrict__ a_indices, const T *__restrict__ a_data,
                    const I *__restrict__ b_indptr, const I *__restrict__ b_indices, const T *__restrict__ b_data,
                    const int64_t *__restrict__ Pv, const int64_t *__restrict__ ub_off,
                    int64_t *__restrict__ tmp_idx, T *__restrict__ tmp_val, int64_t *__restrict__ row_nnz,
                    int64_t *__restrict__ row_nz, unsigned char *__restrict__ scratch, size_t per_cta,
                    int64_t Hmax, int64_t Pmax) {
    constexpr int THREADS = WARPS * 32;
    constexpr I EMPTY = Empty<I>::value;
    __shared__ I st_key[TILE];
    __shared__ T st_val[TILE];
    __shared__ int64_t s_off[THREADS + 1];
    __shared__ int64_t s_bs[THREADS];
    __shared__ T s_av[THREADS];
    __shared__ int64_t s_red[WARPS + 1];
    __shared__ int64_t s_carry;

    // carve this CTA's scratch
    unsigned char *base = scratch + (size_t)blockIdx.x * per_cta;
    const int64_t nwords_max = (Pmax + 31) / 32 + 2;
    I *tb_key = reinterpret_cast<I *>(base);
    int64_t *tb_seq = reinterpret_cast<int64_t *>(base + (((size_t)Hmax * sizeof(I) + 15) & ~(size_t)15));
    W *tb_sum = reinterpret_cast<W *>(reinterpret_cast<unsigned char *>(tb_seq) + (size_t)Hmax * 8);
    unsigned *bits = reinterpret_cast<unsigned *>(reinterpret_cast<unsigned char *>(tb_sum) + (size_t)Hmax * 8);
    int64_t *wsuf = reinterpret_cast<int64_t *>(reinterpret_cast<unsigned char *>(bits) +
                                                 (((size_t)nwords_max * 4 + 15) & ~(size_t)15));

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int w = tid >> 5;

    for (int64_t ri = blockIdx.x; ri < n_rows; ri += gridDim.x) {
        const int64_t row = rows[ri];
        const int64_t P = Pv[row];
        const int64_t cap = P < n_col ? P : n_col;
        int64_t H = 64;
        while (H < 2 * cap) H <<= 1;  // H <= Hmax by construction on the host
        const int64_t nwords = (P + 31) / 32;
        for (int64_t s = tid; s < H; s += THREADS) tb_key[s] = EMPTY;
        for (int64_t s = tid; s < nwords + 1; s += THREADS) bits[s] = 0u;
        __syncthreads();

        const int64_t as = (int64_t)a_indptr[row], ae = (int64_t)a_indptr[row + 1];
        int64_t t_base = 0;  // sequence number of the first product of the current A chunk
        int64_t my_new = 0;
        for (int64_t ab = as; ab < ae; ab += THREADS) {
            // per-thread A entry -> B row extent; block exclusive scan of the lengths
            const bool live = ab + tid < ae;
            int64_t bs = 0, len = 0;
            T av = T(0);
            if (live) {
                const I j = a_indices[ab + tid];
                av = a_data[ab + tid];
                bs = (int64_t)b_indptr[j];
                len = (int64_t)b_indptr[j + 1] - bs;
            }
            int64_t incl = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int64_t v = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) s_red[w] = incl;
            __syncthreads();
            int64_t woff = 0;
            for (int i = 0; i < w; ++i) woff += s_red[i];
            int64_t chunk_total = 0;
            for (int i = 0; i < WARPS; ++i) chunk_total += s_red[i];
            s_off[tid] = woff + incl - len;
            s_bs[tid] = bs;
            s_av[tid] = av;
            if (tid == 0) s_off[THREADS] = chunk_total;
            __syncthreads();

            for (int64_t tile0 = 0; tile0 < chunk_total; tile0 += TILE) {
                const int64_t tn = (chunk_total - tile0) < TILE ? (chunk_total - tile0) : TILE;
                // stage the tile (parallel gather, visiting order preserved by position)
                for (int64_t x = tid; x < tn; x += THREADS) {
                    const int64_t t = tile0 + x;
                    int lo = 0;
#pragma unroll
                    for (int step = THREADS / 2; step > 0; step >>= 1)
                        if (s_off[lo + step] <= t) lo += step;
                    const int64_t src = s_bs[lo] + (t - s_off[lo]);
                    st_key[x] = b_indices[src];
                    st_val[x] = mul_rn(s_av[lo], b_data[src]);
                }
                __syncthreads();
                // every warp scans the tile in order and handles the columns it owns
                for (int64_t c = 0; c < tn; c += 32) {
                    const int64_t x = c + lane;
                    I k = EMPTY;
                    T p = T(0);
                    bool mine = false;
                    if (x < tn) {
                        k = st_key[x];
                        p = st_val[x];
                        mine = ((hash_col<I>(k) >> 20) % WARPS) == (unsigned)w;
                    }
                    const unsigned amask = __ballot_sync(FULL, mine);
                    if (mine) {
                        const unsigned grp = __match_any_sync(amask, k);
                        const int leader = __ffs(grp) - 1;
                        const int gsz = __popc(grp);
                        const bool lead = lane == leader;
                        int64_t slot = 0;
                        bool isnew = false;
                        if (lead) {
                            int64_t h = (int64_t)(hash_col<I>(k) & (unsigned)(H - 1));
                            while (true) {
                                const I cur = *((volatile I *)&tb_key[h]);
                                if (cur == k) break;
                                if (cur == EMPTY) {
                                    const I old = cas_key(&tb_key[h], EMPTY, k);
                                    if (old == EMPTY) {
                                        isnew = true;
                                        break;
                                    }
                                    if (old == k) break;
                                }
                                h = (h + 1) & (H - 1);
                            }
                            slot = h;
                            if (isnew) {
                                const int64_t t = t_base + tile0 + x;
                                tb_seq[slot] = t;
                                tb_sum[slot] = W(0);
                                atomicOr(&bits[t >> 5], 1u << (t & 31));
                                ++my_new;
                            }
                        }
                        const int maxg = __reduce_max_sync(amask, gsz);
                        W s = lead ? tb_sum[slot] : W(0);
                        for (int r = 0; r < maxg; ++r) {
                            const bool take = lead && r < gsz;
                            const int src = take ? (int)__fns(grp, 0, r + 1) : lane;
                            const T v = __shfl_sync(amask, p, src);
                            if (take) s = add_rn(s, (W)v);
                        }
                        if (lead) tb_sum[slot] = s;
                    }
                    __syncwarp();
                }
                __syncthreads();
            }
            t_base += chunk_total;
            __syncthreads();
        }

        // distinct count of the row
        int64_t cnt = my_new;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(FULL, cnt, o);
        if (lane == 0) s_red[w] = cnt;
        __threadfence_block();
        __syncthreads();
        int64_t distinct = 0;
        for (int i = 0; i < WARPS; ++i) distinct += s_red[i];

        // suffix popcounts over the sequence bitmap: wsuf[i] = number of first touches in words >= i
        if (tid == 0) s_carry = 0;
        __syncthreads();
        for (int64_t hi = nwords; hi > 0; hi -= THREADS) {
            const int64_t idx = hi - 1 - tid;  // thread 0 takes the highest word of this tile
            int64_t c = (idx >= 0) ? (int64_t)__popc(bits[idx]) : 0;
            int64_t incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int64_t v = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) s_off[w] = incl;
            __syncthreads();
            int64_t woff = 0;
            for (int i = 0; i < w; ++i) woff += s_off[i];
            const int64_t carry = s_carry;
            if (idx >= 0) wsuf[idx] = carry + woff + incl;
            __syncthreads();
            if (tid == THREADS - 1) s_carry = carry + woff + incl;
            __syncthreads();
        }
        if (tid == 0) wsuf[nwords] = 0;
        __threadfence_block();
        __syncthreads();

        const int64_t ub = ub_off[row];
        int64_t nz = 0;
        for (int64_t s = tid; s < H; s += THREADS) {
            const I k = tb_key[s];
            if (k != EMPTY) {
                const int64_t t = tb_seq[s];
                const int64_t wd = t >> 5;
                const int64_t above = wsuf[wd + 1] + __popc((bits[wd] >> (t & 31)) >> 1);
                const T v = narrow_sum<T, W>(tb_sum[s]);
                tmp_idx[ub + above] = (int64_t)k;
                tmp_val[ub + above] = v;
                nz += is_pos_zero_bits(v) ? 0 : 1;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nz += __shfl_xor_sync(FULL, nz, o);
        __syncthreads();
        if (lane == 0) s_red[w] = nz;
        __syncthreads();
        if (tid == 0) {
            int64_t tot = 0;
            for (int i = 0; i < WARPS; ++i) tot += s_red[i];
            row_nnz[row] = distinct;
            row_nz[row] = tot;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// 4. finish: compaction from the upper-bound layout (+ optional prune, row reversal, COO rows,
//    per-row ascending sort for long rows in SORTED mode)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void spgemm_finish_kernel(int64_t M, const int64_t *__restrict__ ub_off,
                                     const int64_t *__restrict__ row_nnz, const int64_t *__restrict__ out_ptr,
                                     const int64_t *__restrict__ tmp_idx, const T *__restrict__ tmp_val, int prune,
                                     int reverse, int64_t *__restrict__ out_idx, int64_t *__restrict__ out_rows,
                                     T *__restrict__ out_val) {
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp; row < M; row += nwarps) {
        const int64_t n = row_nnz[row];
        if (n == 0) continue;
        const int64_t src = ub_off[row];
        int64_t dst = out_ptr[row];
        for (int64_t c = 0; c < n; c += 32) {
            const int64_t p = c + lane;
            bool keep = false;
            int64_t k = 0;
            T v = T(0);
            if (p < n) {
                const int64_t q = reverse ? (n - 1 - p) : p;
                k = tmp_idx[src + q];
                v = tmp_val[src + q];
                keep = !prune || !is_pos_zero_bits(v);
            }
            const unsigned m = __ballot_sync(FULL, keep);
            if (keep) {
                const int64_t o = dst + __popc(m & ((1u << lane) - 1));
                out_idx[o] = k;
                out_val[o] = v;
                if (out_rows) out_rows[o] = row;
            }
            dst += __popc(m);
        }
    }
}

__global__ void zero_i64_kernel(int64_t *p, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        p[i] = 0;
}

// per-row ascending sort of (idx, val) segments for rows that went through the block kernel in SORTED mode
// (rare, long rows): one CTA per row, odd-even transposition over global memory is too slow, so use a
// bitonic-free approach: rank by counting inside the CTA (O(n^2 / threads)); rows here have n <= n_col.
template <typename T>
__global__ void sort_long_rows_kernel(const int64_t *__restrict__ rows, int64_t n_rows,
                                      const int64_t *__restrict__ ub_off, const int64_t *__restrict__ row_nnz,
                                      int64_t *__restrict__ tmp_idx, T *__restrict__ tmp_val,
                                      int64_t *__restrict__ sc_idx, T *__restrict__ sc_val) {
    for (int64_t ri = blockIdx.x; ri < n_rows; ri += gridDim.x) {
        const int64_t row = rows[ri];
        const int64_t n = row_nnz[row];
        const int64_t off = ub_off[row];
        for (int64_t e = threadIdx.x; e < n; e += blockDim.x) {
            const int64_t k = tmp_idx[off + e];
            int64_t rank = 0;
            for (int64_t f = 0; f < n; ++f) rank += tmp_idx[off + f] < k ? 1 : 0;
            sc_idx[off + rank] = k;
            sc_val[off + rank] = tmp_val[off + e];
        }
        __syncthreads();
        for (int64_t e = threadIdx.x; e < n; e += blockDim.x) {
            tmp_idx[off + e] = sc_idx[off + e];
            tmp_val[off + e] = sc_val[off + e];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// plan object kept between begin and finish
// ---------------------------------------------------------------------------------------------
struct SpgemmPlan {
    int dtype;
    int sorted;
    int64_t M, n_col;
    int64_t nnz_struct, nnz_pruned, ub_total;
    int64_t *P, *U, *ub_off, *row_nnz, *row_nz;  // [M] (+1 for offsets)
    int64_t *tmp_idx;
    void *tmp_val;
    cudaStream_t stream;
};

static int64_t g_t0 = 64, g_t1 = 128, g_t2 = 256;  // bin thresholds on products per row (tests may lower them)

static int exclusive_scan_i64(const int64_t *in, int64_t *out, int64_t n, cudaStream_t s) {
    if (n == 0) return B2S_OK;
    size_t tmp_bytes = 0;
    B2S_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, out, (int)n, s));
    void *tmp = nullptr;
    int rc = scratch_alloc(&tmp, tmp_bytes, s);
    if (rc) return rc;
    B2S_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, in, out, (int)n, s));
    count_launch(2);
    return scratch_free(tmp, s);
}

template <typename T, typename W, typename I>
static int spgemm_numeric(SpgemmPlan *pl, const void *a_indptr, const void *a_indices, const void *a_data,
                          const void *b_indptr, const void *b_indices, const void *b_data) {
    cudaStream_t s = pl->stream;
    const int64_t M = pl->M;
    const I *ap = (const I *)a_indptr, *ai = (const I *)a_indices, *bp = (const I *)b_indptr,
            *bi = (const I *)b_indices;
    const T *ad = (const T *)a_data, *bd = (const T *)b_data;
    int rc;
    // 1. products per row and upper bounds
    {
        const int64_t threads = M * 8;
        row_products_kernel<I><<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(M, pl->n_col, ap, ai, bp, pl->P, pl->U);
        B2S_CHECK_LAUNCH();
    }
    if ((rc = exclusive_scan_i64(pl->U, pl->ub_off, M + 1, s))) return rc;  // U[M] = 0 sentinel
    B2S_CUDA(cudaMemcpyAsync(&pl->ub_total, pl->ub_off + M, 8, cudaMemcpyDeviceToHost, s));
    // 2. bins
    int64_t *lists = nullptr;
    unsigned long long *counts = nullptr;
    if ((rc = scratch_alloc((void **)&lists, (size_t)M * 4 * 8, s))) return rc;
    if ((rc = scratch_alloc((void **)&counts, 5 * 8, s))) return rc;
    B2S_CUDA(cudaMemsetAsync(counts, 0, 40, s));
    bin_rows_kernel<<<(unsigned)((M + 255) / 256), 256, 0, s>>>(M, pl->P, g_t0, g_t1, g_t2, lists, counts, counts + 4);
    B2S_CHECK_LAUNCH();
    unsigned long long hc[5];
    B2S_CUDA(cudaMemcpyAsync(hc, counts, 40, cudaMemcpyDeviceToHost, s));
    B2S_CUDA(cudaStreamSynchronize(s));
    // upper-bound output buffers
    if ((rc = scratch_alloc((void **)&pl->tmp_idx, (size_t)pl->ub_total * 8, s))) return rc;
    if ((rc = scratch_alloc((void **)&pl->tmp_val, (size_t)pl->ub_total * sizeof(T), s))) return rc;
    zero_i64_kernel<<<(unsigned)((M + 255) / 256 > 2048 ? 2048 : (M + 255) / 256), 256, 0, s>>>(pl->row_nnz, M + 1);
    zero_i64_kernel<<<(unsigned)((M + 255) / 256 > 2048 ? 2048 : (M + 255) / 256), 256, 0, s>>>(pl->row_nz, M + 1);
    count_launch(2);
    const int sms = num_sms();
    // 3a. warp-per-row bins: persistent grids sized by the occupancy the shared-memory footprint allows
#define B2S_WARP_BIN(BIN, H, WARPS)                                                                                  \
    if (hc[BIN]) {                                                                                                   \
        auto kern_s = spgemm_warp_kernel<T, W, I, H, WARPS, true>;                                                   \
        auto kern_r = spgemm_warp_kernel<T, W, I, H, WARPS, false>;                                                  \
        int occ = 1;                                                                                                 \
        if (pl->sorted) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern_s, WARPS * 32, 0);                  \
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern_r, WARPS * 32, 0);                             \
        if (occ < 1) occ = 1;                                                                                        \
        int64_t blocks = ((int64_t)hc[BIN] + WARPS - 1) / WARPS;                                                     \
        if (blocks > (int64_t)sms * occ) blocks = (int64_t)sms * occ;                                                \
        const int64_t *lst = lists + (int64_t)BIN * M;                                                               \
        if (pl->sorted)                                                                                              \
            kern_s<<<(unsigned)blocks, WARPS * 32, 0, s>>>(lst, (int64_t)hc[BIN], ap, ai, ad, bp, bi, bd, pl->ub_off, \
                                                           pl->tmp_idx, (T *)pl->tmp_val, pl->row_nnz, pl->row_nz);   \
        else                                                                                                         \
            kern_r<<<(unsigned)blocks, WARPS * 32, 0, s>>>(lst, (int64_t)hc[BIN], ap, ai, ad, bp, bi, bd, pl->ub_off, \
                                                           pl->tmp_idx, (T *)pl->tmp_val, pl->row_nnz, pl->row_nz);   \
        B2S_CHECK_LAUNCH();                                                                                          \
    }
    constexpr bool kWide = (sizeof(W) + sizeof(I)) > 8;  // keeps static shared memory under 48 KB
    B2S_WARP_BIN(0, 128, 8)
    if constexpr (kWide) {
        B2S_WARP_BIN(1, 256, 4)
        B2S_WARP_BIN(2, 512, 2)
    } else {
        B2S_WARP_BIN(1, 256, 8)
        B2S_WARP_BIN(2, 512, 4)
    }
#undef B2S_WARP_BIN
    // 3b. long rows
    if (hc[3]) {
        constexpr int WARPS = 8;
        constexpr int TILE = 2048;
        const int64_t Pmax = (int64_t)hc[4];
        const int64_t cap = Pmax < pl->n_col ? Pmax : pl->n_col;
        int64_t Hmax = 64;
        while (Hmax < 2 * cap) Hmax <<= 1;
        const size_t nwords_max = (size_t)((Pmax + 31) / 32 + 2);
        size_t per_cta = (((size_t)Hmax * sizeof(I) + 15) & ~(size_t)15) + (size_t)Hmax * 8 + (size_t)Hmax * 8 +
                         ((nwords_max * 4 + 15) & ~(size_t)15) + nwords_max * 8 + 64;
        per_cta = (per_cta + 255) & ~(size_t)255;
        int64_t ctas = (int64_t)hc[3] < (int64_t)sms * 2 ? (int64_t)hc[3] : (int64_t)sms * 2;
        const size_t budget = (size_t)8 << 30;
        while (ctas > 1 && (size_t)ctas * per_cta > budget) ctas /= 2;
        unsigned char *scratch = nullptr;
        if ((rc = scratch_alloc((void **)&scratch, (size_t)ctas * per_cta, s))) return rc;
        spgemm_block_kernel<T, W, I, WARPS, TILE><<<(unsigned)ctas, WARPS * 32, 0, s>>>(
            lists + 3 * M, (int64_t)hc[3], pl->n_col, ap, ai, ad, bp, bi, bd, pl->P, pl->ub_off, pl->tmp_idx,
            (T *)pl->tmp_val, pl->row_nnz, pl->row_nz, scratch, per_cta, Hmax, Pmax);
        B2S_CHECK_LAUNCH();
        if (pl->sorted) {
            int64_t *sc_idx = nullptr;
            T *sc_val = nullptr;
            if ((rc = scratch_alloc((void **)&sc_idx, (size_t)pl->ub_total * 8, s))) return rc;
            if ((rc = scratch_alloc((void **)&sc_val, (size_t)pl->ub_total * sizeof(T), s))) return rc;
            sort_long_rows_kernel<T><<<(unsigned)ctas, 256, 0, s>>>(lists + 3 * M, (int64_t)hc[3], pl->ub_off, pl->row_nnz,
                                                                   pl->tmp_idx, (T *)pl->tmp_val, sc_idx, sc_val);
            B2S_CHECK_LAUNCH();
            scratch_free(sc_idx, s);
            scratch_free(sc_val, s);
        }
        scratch_free(scratch, s);
    }
    scratch_free(lists, s);
    scratch_free(counts, s);
    // totals: structural nnz and pruned nnz
    int64_t *red = nullptr;
    if ((rc = scratch_alloc((void **)&red, 16, s))) return rc;
    {
        size_t tb = 0;
        B2S_CUDA(cub::DeviceReduce::Sum(nullptr, tb, pl->row_nnz, red, (int)M, s));
        void *tmp = nullptr;
        if ((rc = scratch_alloc(&tmp, tb, s))) return rc;
        B2S_CUDA(cub::DeviceReduce::Sum(tmp, tb, pl->row_nnz, red, (int)M, s));
        B2S_CUDA(cub::DeviceReduce::Sum(tmp, tb, pl->row_nz, red + 1, (int)M, s));
        count_launch(2);
        scratch_free(tmp, s);
    }
    int64_t h[2];
    B2S_CUDA(cudaMemcpyAsync(h, red, 16, cudaMemcpyDeviceToHost, s));
    B2S_CUDA(cudaStreamSynchronize(s));
    scratch_free(red, s);
    pl->nnz_struct = h[0];
    pl->nnz_pruned = h[1];
    return B2S_OK;
}

template <typename T>
static int spgemm_finish_t(SpgemmPlan *pl, int prune, int64_t *indptr_out, int64_t *indices_out, int64_t *rows_out,
                           void *data_out) {
    cudaStream_t s = pl->stream;
    const int64_t M = pl->M;
    int rc;
    const int64_t *cnt = prune ? pl->row_nz : pl->row_nnz;
    // exclusive scan of the per-row counts -> indptr (M+1 entries; cnt[M] == 0 sentinel)
    int64_t *ptr = indptr_out;
    int64_t *own = nullptr;
    if (!ptr) {
        if ((rc = scratch_alloc((void **)&own, (size_t)(M + 1) * 8, s))) return rc;
        ptr = own;
    }
    if ((rc = exclusive_scan_i64(cnt, ptr, M + 1, s))) return rc;
    const int reverse = (!pl->sorted && pl->n_col > 0 && pl->nnz_struct == pl->M * pl->n_col) ? 1 : 0;
    if (pl->nnz_struct > 0) {
        int64_t blocks = (M * 32 + 255) / 256;
        if (blocks > (int64_t)num_sms() * 16) blocks = (int64_t)num_sms() * 16;
        spgemm_finish_kernel<T><<<(unsigned)blocks, 256, 0, s>>>(M, pl->ub_off, pl->row_nnz, ptr, pl->tmp_idx,
                                                                (const T *)pl->tmp_val, prune, reverse, indices_out,
                                                                rows_out, (T *)data_out);
        B2S_CHECK_LAUNCH();
    }
    if (own) scratch_free(own, s);
    return B2S_OK;
}

static void plan_free(SpgemmPlan *pl) {
    cudaStream_t s = pl->stream;
    scratch_free(pl->P, s);
    scratch_free(pl->U, s);
    scratch_free(pl->ub_off, s);
    scratch_free(pl->row_nnz, s);
    scratch_free(pl->row_nz, s);
    scratch_free(pl->tmp_idx, s);
    scratch_free(pl->tmp_val, s);
    delete pl;
}

}  // namespace b2s

using namespace b2s;

extern "C" {

int b2s_spgemm_set_thresholds(int64_t t0, int64_t t1) {
    g_t0 = (t0 >= 1 && t0 <= 64) ? t0 : 64;
    g_t2 = (t1 >= g_t0 && t1 <= 256) ? t1 : 256;
    g_t1 = g_t2 < 128 ? g_t2 : (g_t0 > 128 ? g_t0 : 128);
    if (g_t1 > g_t2) g_t1 = g_t2;
    return B2S_OK;
}

int b2s_spgemm_begin(int dtype, int idx_bytes, int64_t M, int64_t K, int64_t n_col, const void *a_indptr_dev,
                     const void *a_indices_dev, const void *a_data_dev, const void *b_indptr_dev,
                     const void *b_indices_dev, const void *b_data_dev, int sorted_order, int wide_accumulate,
                     void **plan_out,
                     int64_t *nnz_struct_out, int64_t *nnz_pruned_out, void *stream) {
    B2S_REQUIRE(plan_out && nnz_struct_out && nnz_pruned_out, B2S_ERR_INVALID, "spgemm_begin: NULL output");
    B2S_REQUIRE(M >= 0 && K >= 0 && n_col >= 0, B2S_ERR_INVALID, "spgemm_begin: negative dimension");
    B2S_REQUIRE(idx_bytes == 4 || idx_bytes == 8, B2S_ERR_INVALID, "spgemm_begin: idx_bytes");
    B2S_REQUIRE(M + 1 < 2147483647LL, B2S_ERR_OVERFLOW, "spgemm_begin: M too large for the device scan");
    cudaStream_t s = (cudaStream_t)stream;
    SpgemmPlan *pl = new SpgemmPlan();
    memset(pl, 0, sizeof(*pl));
    pl->dtype = dtype;
    pl->sorted = sorted_order ? 1 : 0;
    const bool wide = wide_accumulate != 0;
    pl->M = M;
    pl->n_col = n_col;
    pl->stream = s;
    int rc = B2S_OK;
    const size_t mb = (size_t)(M + 1) * 8;
    if ((rc = scratch_alloc((void **)&pl->P, mb, s)) || (rc = scratch_alloc((void **)&pl->U, mb, s)) ||
        (rc = scratch_alloc((void **)&pl->ub_off, mb, s)) || (rc = scratch_alloc((void **)&pl->row_nnz, mb, s)) ||
        (rc = scratch_alloc((void **)&pl->row_nz, mb, s))) {
        plan_free(pl);
        return rc;
    }
    cudaMemsetAsync(pl->P, 0, mb, s);
    cudaMemsetAsync(pl->U, 0, mb, s);
    if (M > 0) {
#define B2S_NUM(T, I)                                                                                         \
    rc = wide ? spgemm_numeric<T, double, I>(pl, a_indptr_dev, a_indices_dev, a_data_dev, b_indptr_dev,        \
                                             b_indices_dev, b_data_dev)                                       \
              : spgemm_numeric<T, T, I>(pl, a_indptr_dev, a_indices_dev, a_data_dev, b_indptr_dev,             \
                                        b_indices_dev, b_data_dev)
        if (idx_bytes == 4) {
            switch (dtype) {
                case B2S_F32: B2S_NUM(float, int32_t); break;
                case B2S_F64: B2S_NUM(double, int32_t); break;
                case B2S_I32: B2S_NUM(int32_t, int32_t); break;
                case B2S_I64: B2S_NUM(int64_t, int32_t); break;
                default: set_error("spgemm: dtype %d", dtype); rc = B2S_ERR_UNSUPPORTED;
            }
        } else {
            switch (dtype) {
                case B2S_F32: B2S_NUM(float, int64_t); break;
                case B2S_F64: B2S_NUM(double, int64_t); break;
                case B2S_I32: B2S_NUM(int32_t, int64_t); break;
                case B2S_I64: B2S_NUM(int64_t, int64_t); break;
                default: set_error("spgemm: dtype %d", dtype); rc = B2S_ERR_UNSUPPORTED;
            }
        }
#undef B2S_NUM
    }
    if (rc != B2S_OK) {
        plan_free(pl);
        return rc;
    }
    *plan_out = pl;
    *nnz_struct_out = pl->nnz_struct;
    *nnz_pruned_out = pl->nnz_pruned;
    return B2S_OK;
}

int b2s_spgemm_finish(void *plan, int prune, int64_t *indptr_out_dev, int64_t *indices_out_dev,
                      int64_t *rows_out_dev, void *data_out_dev) {
    B2S_REQUIRE(plan != nullptr, B2S_ERR_INVALID, "spgemm_finish: NULL plan");
    SpgemmPlan *pl = (SpgemmPlan *)plan;
    int rc;
    switch (pl->dtype) {
        case B2S_F32: rc = spgemm_finish_t<float>(pl, prune, indptr_out_dev, indices_out_dev, rows_out_dev, data_out_dev); break;
        case B2S_F64: rc = spgemm_finish_t<double>(pl, prune, indptr_out_dev, indices_out_dev, rows_out_dev, data_out_dev); break;
        case B2S_I32: rc = spgemm_finish_t<int32_t>(pl, prune, indptr_out_dev, indices_out_dev, rows_out_dev, data_out_dev); break;
        case B2S_I64: rc = spgemm_finish_t<int64_t>(pl, prune, indptr_out_dev, indices_out_dev, rows_out_dev, data_out_dev); break;
        default: rc = B2S_ERR_UNSUPPORTED;
    }
    plan_free(pl);
    return rc;
}

int b2s_spgemm_abort(void *plan) {
    if (plan) plan_free((SpgemmPlan *)plan);
    return B2S_OK;
}

}  // extern "C"
