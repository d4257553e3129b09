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
// Structure (ONE pass over the products, final arrays written directly -- no symbolic pass, no upper-bound layout,
// no compaction pass, no scan kernels):
//   1. products: P_i = sum over A-row entries of the B-row lengths (8 lanes per row), U_i = min(P_i, n_col);
//      sum U_i bounds the output size (the caller allocates that much and trims), rows with P_i > 256 are listed.
//   2. (rare) long rows: CTA-per-row kernel with a global-memory hash, results parked in a side buffer.
//   3. ordered kernel: rows are taken in ROW ORDER, WARPS consecutive rows per tile, tiles in ticket order.  A warp
//      builds its row in a shared-memory hash table sized for the row (32 .. 512 slots, load <= 1/2): the products are
//      generated 32 at a time in the reference's visiting order, equal columns inside a chunk are grouped by
//      __match_any_sync so their adds stay sequential, slots are claimed WITHOUT atomics (store, __syncwarp, re-read:
//      the table belongs to one warp), and the n-th first touch records its slot in ord[n] -- so the row comes out
//      in REVERSE first-touch order (REF; exactly the reference's intrusive linked list) or ranked by column
//      (SORTED) straight from shared memory.  The tile's row counts are scanned, the tile's offset comes from a
//      DECOUPLED LOOK-BACK over the preceding tiles' (status | count) words, and every warp writes its row -- pruned of
//      +0 if asked (the prune=True of _common.py:374-379) -- coalesced at its final position, together with indptr.
//   4. (tiny matrices only) the row reversal of a completely dense result (:709-714), in place.
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
// 1. products per row, output bound, list of long rows
// ---------------------------------------------------------------------------------------------
// counters: [0] sum of U over all rows, [1] number of long rows, [2] sum of U over long rows, [3] max P of a long row
template <typename I>
__global__ void __launch_bounds__(256)
spgemm_products_kernel(int64_t M, int64_t n_col, int64_t pmax_short, const I *__restrict__ a_indptr,
                       const I *__restrict__ a_indices, const I *__restrict__ b_indptr, int64_t *__restrict__ P,
                       int64_t *__restrict__ long_rows, int64_t *__restrict__ side_off,
                       unsigned long long *__restrict__ counters) {
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
    int64_t u = 0;
    if (row < M && sub == 0) {
        P[row] = acc;
        u = acc < n_col ? acc : n_col;
        if (acc > pmax_short) {
            const unsigned long long at = atomicAdd(counters + 1, 1ull);
            long_rows[at] = row;
            side_off[row] = (int64_t)atomicAdd(counters + 2, (unsigned long long)u);
            atomicMax(counters + 3, (unsigned long long)acc);
        }
    }
    // one atomic per CTA for the output bound
    __shared__ int64_t s_part[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) u += __shfl_xor_sync(FULL, u, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = u;
    __syncthreads();
    if (threadIdx.x == 0) {
        int64_t t = 0;
        for (int i = 0; i < 8; ++i) t += s_part[i];
        if (t) atomicAdd(counters + 0, (unsigned long long)t);
    }
}

// ---------------------------------------------------------------------------------------------
// 3. ordered single-pass numeric kernel (see the header comment)
// ---------------------------------------------------------------------------------------------
#define SG_AGG (1ull << 62)
#define SG_PREFIX (2ull << 62)
#define SG_VAL ((1ull << 62) - 1)

template <typename T, typename W, typename I, int HMAX, int WARPS>
struct OrderedSmem {
    static constexpr int PMAX = HMAX / 2;
    // per-warp carve-up (bytes), 16-byte aligned pieces
    static constexpr size_t key_b = ((size_t)HMAX * sizeof(I) + 15) & ~(size_t)15;
    static constexpr size_t sum_b = ((size_t)HMAX * sizeof(W) + 15) & ~(size_t)15;
    static constexpr size_t ord_b = ((size_t)PMAX * 2 + 15) & ~(size_t)15;
    static constexpr size_t off_b = 36 * 4;                 // exclusive product offsets of the A chunk (+ total)
    static constexpr size_t bs_b = 32 * 8;                  // B row starts
    static constexpr size_t av_b = ((size_t)32 * sizeof(T) + 15) & ~(size_t)15;
    static constexpr size_t per_warp = key_b + sum_b + ord_b + off_b + bs_b + av_b;
    static constexpr size_t total = per_warp * WARPS;
};

template <typename T, typename W, typename I, int HMAX, int WARPS, bool SORTED>
__global__ void __launch_bounds__(WARPS * 32)
spgemm_ordered_kernel(int64_t M, const I *__restrict__ a_indptr, const I *__restrict__ a_indices,
                      const T *__restrict__ a_data, const I *__restrict__ b_indptr, const I *__restrict__ b_indices,
                      const T *__restrict__ b_data, const int64_t *__restrict__ Pv,
                      const int64_t *__restrict__ side_off, const int64_t *__restrict__ side_idx,
                      const T *__restrict__ side_val, const int64_t *__restrict__ long_nnz,
                      const int64_t *__restrict__ long_nz, int prune, uint64_t *__restrict__ status,
                      unsigned int *__restrict__ ticket, int64_t *__restrict__ out_ptr, int64_t *__restrict__ out_idx,
                      int64_t *__restrict__ out_rows, T *__restrict__ out_val, unsigned long long *__restrict__ totals) {
    using L = OrderedSmem<T, W, I, HMAX, WARPS>;
    constexpr int PMAX = L::PMAX;
    constexpr I EMPTY = Empty<I>::value;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int64_t s_excl[WARPS];
    __shared__ int64_t s_prefix;
    __shared__ int64_t s_tile;

    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    unsigned char *mine = smem_raw + (size_t)w * L::per_warp;
    I *key = reinterpret_cast<I *>(mine);
    W *sum = reinterpret_cast<W *>(mine + L::key_b);
    uint16_t *ord = reinterpret_cast<uint16_t *>(mine + L::key_b + L::sum_b);
    int *s_off = reinterpret_cast<int *>(mine + L::key_b + L::sum_b + L::ord_b);
    int64_t *s_bs = reinterpret_cast<int64_t *>(mine + L::key_b + L::sum_b + L::ord_b + L::off_b);
    T *s_av = reinterpret_cast<T *>(mine + L::key_b + L::sum_b + L::ord_b + L::off_b + L::bs_b);

    const int64_t n_tiles = (M + WARPS - 1) / WARPS;
    unsigned long long my_struct = 0;  // structural entries of the rows this warp handled (lane 0 counts)

    while (true) {
        __syncthreads();  // everybody is done with the previous tile's shared state
        if (threadIdx.x == 0) s_tile = (int64_t)atomicAdd(ticket, 1u);
        __syncthreads();
        const int64_t tile = s_tile;
        if (tile >= n_tiles) break;
        const int64_t row = tile * WARPS + w;
        const int64_t P = row < M ? Pv[row] : 0;
        const bool is_long = P > PMAX;
        int distinct = 0;  // structural entries of a short row
        int64_t cnt = 0;   // entries this row contributes to the output

        if (P > 0 && !is_long) {
            int H = 32;
            while (H < 2 * (int)P) H <<= 1;
            const unsigned mask = (unsigned)(H - 1);
            for (int s = lane; s < H; s += 32) key[s] = EMPTY;
            __syncwarp();
            const int64_t as = (int64_t)a_indptr[row], ae = (int64_t)a_indptr[row + 1];
            for (int64_t ab = as; ab < ae; ab += 32) {
                // ---- the A chunk: B row extents and their exclusive offsets -------------------------------
                const bool live = ab + lane < ae;
                T av = T(0);
                int64_t bs = 0;
                int len = 0;
                if (live) {
                    const I j = a_indices[ab + lane];
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
                s_off[lane] = incl - len;
                s_bs[lane] = bs;
                s_av[lane] = av;
                __syncwarp();
                // ---- its products, 32 at a time, in visiting order ----------------------------------------
                for (int t0 = 0; t0 < total; t0 += 32) {
                    const int t = t0 + lane;
                    const bool active = t < total;
                    I k = EMPTY;
                    T p = T(0);
                    if (active) {
                        int lo = 0;
#pragma unroll
                        for (int step = 16; step > 0; step >>= 1)
                            if (s_off[lo + step] <= t) lo += step;
                        const int64_t src = s_bs[lo] + (t - s_off[lo]);
                        k = b_indices[src];
                        p = mul_rn(s_av[lo], b_data[src]);
                    }
                    const unsigned amask = __ballot_sync(FULL, active);
                    unsigned grp = 0;
                    if (active) grp = __match_any_sync(amask, k);
                    const bool lead = active && (lane == __ffs(grp) - 1);
                    const int gsz = __popc(grp);
                    // claim / find the slot -- no atomics: the table belongs to this warp; a lane that sees an empty
                    // slot stores its key, everybody synchronises, and the lane whose key is there owns the slot
                    unsigned h = hash_col<I>(k) & mask;
                    bool pending = lead, isnew = false;
                    while (__any_sync(FULL, pending)) {
                        I cur = EMPTY;
                        if (pending) cur = key[h];
                        const bool empty = pending && cur == EMPTY;
                        __syncwarp();
                        if (empty) key[h] = k;
                        __syncwarp();
                        if (empty) {
                            cur = key[h];
                            isnew = cur == k;
                        }
                        if (pending) {
                            if (cur == k) pending = false;
                            else h = (h + 1) & mask;
                        }
                    }
                    // the n-th first touch of the row remembers its slot: ord[n] = slot
                    const unsigned newmask = __ballot_sync(FULL, isnew);
                    if (isnew) ord[distinct + __popc(newmask & lt)] = (uint16_t)h;
                    distinct += __popc(newmask);
                    // accumulate: adds to one column stay in visiting order (lane order inside the chunk)
                    if (!__any_sync(FULL, lead && gsz > 1)) {
                        if (lead) {
                            const W s0 = isnew ? W(0) : sum[h];
                            sum[h] = add_rn(s0, (W)p);
                        }
                    } else {
                        const int maxg = __reduce_max_sync(FULL, lead ? gsz : 0);
                        W s0 = (lead && !isnew) ? sum[h] : W(0);
                        for (int r = 0; r < maxg; ++r) {
                            const bool take = lead && r < gsz;
                            const int src = take ? (int)__fns(grp, 0, r + 1) : lane;
                            const T v = __shfl_sync(FULL, p, src);
                            if (take) s0 = add_rn(s0, (W)v);
                        }
                        if (lead) sum[h] = s0;
                    }
                    __syncwarp();
                }
                __syncwarp();
            }
            cnt = distinct;
            if (prune) {  // entries that survive prune=True: everything but +0
                int nz = 0;
                for (int i = lane; i < distinct; i += 32) {
                    const T v = narrow_sum<T, W>(sum[ord[i]]);
                    nz += is_pos_zero_bits(v) ? 0 : 1;
                }
                cnt = __reduce_add_sync(FULL, nz);
            }
            if (lane == 0) my_struct += (unsigned long long)distinct;
        } else if (is_long) {
            cnt = prune ? long_nz[row] : long_nnz[row];
            if (lane == 0) my_struct += (unsigned long long)long_nnz[row];
        }

        // ---- the tile's offset: scan of the WARPS row counts + decoupled look-back over the earlier tiles ----
        if (lane == 0) s_excl[w] = cnt;  // holds the count until warp 0 scans it
        __syncthreads();
        if (w == 0) {
            int64_t c = lane < WARPS ? s_excl[lane] : 0;
            int64_t incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int64_t v = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += v;
            }
            const int64_t tile_total = __shfl_sync(FULL, incl, 31);
            if (lane < WARPS) s_excl[lane] = incl - c;
            if (lane == 0)
                *(volatile uint64_t *)&status[tile] = (tile == 0 ? SG_PREFIX : SG_AGG) | (uint64_t)tile_total;
            int64_t excl = 0;
            if (tile > 0) {
                int64_t look = tile - 1;  // lane 0 looks at the nearest predecessor
                while (true) {
                    const int64_t idx = look - lane;
                    uint64_t sv = idx >= 0 ? *(volatile uint64_t *)&status[idx] : SG_PREFIX;
                    while (__any_sync(FULL, (sv >> 62) == 0)) {
                        if ((sv >> 62) == 0) sv = *(volatile uint64_t *)&status[idx];
                    }
                    const unsigned pm = __ballot_sync(FULL, (sv >> 62) == 2);
                    const int first = pm ? __ffs(pm) - 1 : 31;  // nearest tile that already knows its prefix
                    int64_t v = lane <= first ? (int64_t)(sv & SG_VAL) : 0;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
                    excl += v;
                    if (pm) break;
                    look -= 32;
                }
                if (lane == 0) *(volatile uint64_t *)&status[tile] = SG_PREFIX | (uint64_t)(excl + tile_total);
            }
            if (lane == 0) {
                s_prefix = excl;
                if (tile == n_tiles - 1) {
                    out_ptr[M] = excl + tile_total;
                    totals[1] = (unsigned long long)(excl + tile_total);
                }
            }
        }
        __syncthreads();
        if (row >= M) continue;
        const int64_t off = s_prefix + s_excl[w];
        if (lane == 0) out_ptr[row] = off;

        // ---- write the row at its final position ----------------------------------------------------------------
        if (is_long) {
            const int64_t src = side_off[row];
            const int64_t n = long_nnz[row];
            int64_t base = 0;
            for (int64_t c0 = 0; c0 < n; c0 += 32) {
                const int64_t q = c0 + lane;
                int64_t k = 0;
                T v = T(0);
                bool keep = false;
                if (q < n) {
                    k = side_idx[src + q];
                    v = side_val[src + q];
                    keep = !prune || !is_pos_zero_bits(v);
                }
                const unsigned m = __ballot_sync(FULL, keep);
                if (keep) {
                    const int64_t o = off + base + __popc(m & lt);
                    out_idx[o] = k;
                    out_val[o] = v;
                    if (out_rows) out_rows[o] = row;
                }
                base += __popc(m);
            }
        } else if (distinct > 0) {
            if constexpr (!SORTED) {
                int base = 0;
                for (int i0 = 0; i0 < distinct; i0 += 32) {
                    const int i = i0 + lane;
                    I k = 0;
                    T v = T(0);
                    bool keep = false;
                    if (i < distinct) {
                        const int slot = ord[distinct - 1 - i];  // reverse first-touch order
                        k = key[slot];
                        v = narrow_sum<T, W>(sum[slot]);
                        keep = !prune || !is_pos_zero_bits(v);
                    }
                    const unsigned m = __ballot_sync(FULL, keep);
                    if (keep) {
                        const int64_t o = off + base + __popc(m & lt);
                        out_idx[o] = (int64_t)k;
                        out_val[o] = v;
                        if (out_rows) out_rows[o] = row;
                    }
                    base += __popc(m);
                }
            } else {
                // ascending columns: rank by counting among the kept entries (pruned ones are struck out of ord)
                if (prune) {
                    for (int i = lane; i < distinct; i += 32) {
                        const T v = narrow_sum<T, W>(sum[ord[i]]);
                        if (is_pos_zero_bits(v)) ord[i] = 0xFFFFu;
                    }
                    __syncwarp();
                }
                for (int e = lane; e < distinct; e += 32) {
                    const unsigned slot = ord[e];
                    if (slot == 0xFFFFu) continue;
                    const I k = key[slot];
                    int rank = 0;
                    for (int f = 0; f < distinct; ++f) {
                        const unsigned sf = ord[f];
                        rank += (sf != 0xFFFFu && key[sf] < k) ? 1 : 0;
                    }
                    const int64_t o = off + rank;
                    out_idx[o] = (int64_t)k;
                    out_val[o] = narrow_sum<T, W>(sum[slot]);
                    if (out_rows) out_rows[o] = row;
                }
            }
        }
    }
    if (lane == 0 && my_struct) atomicAdd(totals + 0, my_struct);
}

// completely dense result: the reference re-reverses every row (_common.py:709-714).  In place, warp per row.
template <typename T>
__global__ void spgemm_reverse_rows_kernel(int64_t M, const int64_t *__restrict__ ptr, int64_t *__restrict__ idx,
                                           T *__restrict__ val) {
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp; row < M; row += nwarps) {
        const int64_t s = ptr[row], n = ptr[row + 1] - s;
        for (int64_t i = lane; i < n / 2; i += 32) {
            const int64_t a = s + i, b = s + n - 1 - i;
            const int64_t ka = idx[a], kb = idx[b];
            const T va = val[a], vb = val[b];
            idx[a] = kb;
            idx[b] = ka;
            val[a] = vb;
            val[b] = va;
        }
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
// plan object kept between begin and run
// ---------------------------------------------------------------------------------------------
struct SpgemmPlan {
    int dtype, idx_bytes, sorted, wide;
    int64_t M, n_col;
    int64_t ub_total, n_long, side_total, long_pmax;
    const void *a_indptr, *a_indices, *a_data, *b_indptr, *b_indices, *b_data;
    int64_t *P, *long_rows, *side_off;          // [M]
    unsigned long long *counters;               // [4] (spgemm_products_kernel) + [2] totals of the ordered kernel
    cudaStream_t stream;
};

// rows with more products than this go through the CTA-per-row kernel (tests may lower it to exercise that path)
static int64_t g_pmax_short = 256;

template <typename T, typename W, typename I>
static int spgemm_products(SpgemmPlan *pl) {
    cudaStream_t s = pl->stream;
    const int64_t M = pl->M;
    B2S_CUDA(cudaMemsetAsync(pl->counters, 0, 6 * 8, s));
    const int64_t threads = M * 8;
    spgemm_products_kernel<I><<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(
        M, pl->n_col, g_pmax_short, (const I *)pl->a_indptr, (const I *)pl->a_indices, (const I *)pl->b_indptr, pl->P,
        pl->long_rows, pl->side_off, pl->counters);
    B2S_CHECK_LAUNCH();
    unsigned long long hc[4];
    B2S_CUDA(cudaMemcpyAsync(hc, pl->counters, 32, cudaMemcpyDeviceToHost, s));
    B2S_CUDA(cudaStreamSynchronize(s));
    pl->ub_total = (int64_t)hc[0];
    pl->n_long = (int64_t)hc[1];
    pl->side_total = (int64_t)hc[2];
    pl->long_pmax = (int64_t)hc[3];
    return B2S_OK;
}

template <typename T, typename W, typename I, int HMAX, int WARPS>
static int launch_ordered(SpgemmPlan *pl, int prune, const int64_t *side_idx, const T *side_val,
                          const int64_t *long_nnz, const int64_t *long_nz, uint64_t *status, unsigned int *ticket,
                          int64_t *out_ptr, int64_t *out_idx, int64_t *out_rows, T *out_val) {
    using L = OrderedSmem<T, W, I, HMAX, WARPS>;
    cudaStream_t s = pl->stream;
    auto kern_r = spgemm_ordered_kernel<T, W, I, HMAX, WARPS, false>;
    auto kern_s = spgemm_ordered_kernel<T, W, I, HMAX, WARPS, true>;
    auto kern = pl->sorted ? kern_s : kern_r;
    B2S_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::total));
    int occ = 1;
    B2S_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, L::total));
    if (occ < 1) occ = 1;
    const int64_t n_tiles = (pl->M + WARPS - 1) / WARPS;
    int64_t blocks = (int64_t)num_sms() * occ;
    if (blocks > n_tiles) blocks = n_tiles;
    kern<<<(unsigned)blocks, WARPS * 32, L::total, s>>>(
        pl->M, (const I *)pl->a_indptr, (const I *)pl->a_indices, (const T *)pl->a_data, (const I *)pl->b_indptr,
        (const I *)pl->b_indices, (const T *)pl->b_data, pl->P, pl->side_off, side_idx, side_val, long_nnz, long_nz,
        prune, status, ticket, out_ptr, out_idx, out_rows, out_val, pl->counters + 4);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

template <typename T, typename W, typename I>
static int spgemm_run_t(SpgemmPlan *pl, int prune, int64_t *indptr_out, int64_t *indices_out, int64_t *rows_out,
                        void *data_out, int64_t *nnz_struct_out, int64_t *nnz_out) {
    cudaStream_t s = pl->stream;
    const int64_t M = pl->M;
    const I *ap = (const I *)pl->a_indptr, *ai = (const I *)pl->a_indices, *bp = (const I *)pl->b_indptr,
            *bi = (const I *)pl->b_indices;
    const T *ad = (const T *)pl->a_data, *bd = (const T *)pl->b_data;
    int rc;
    // 2. long rows (rare): CTA per row, global-memory hash, parked in a side buffer in their final order
    int64_t *side_idx = nullptr, *long_nnz = nullptr, *long_nz = nullptr;
    T *side_val = nullptr;
    if (pl->n_long) {
        constexpr int WARPS = 8;
        constexpr int TILE = 2048;
        if ((rc = scratch_alloc((void **)&side_idx, (size_t)pl->side_total * 8, s))) return rc;
        if ((rc = scratch_alloc((void **)&side_val, (size_t)pl->side_total * sizeof(T), s))) return rc;
        if ((rc = scratch_alloc((void **)&long_nnz, (size_t)(M + 1) * 8, s))) return rc;  // indexed by row; only the
        if ((rc = scratch_alloc((void **)&long_nz, (size_t)(M + 1) * 8, s))) return rc;   // long rows' entries are used
        const int sms = num_sms();
        const int64_t Pmax = pl->long_pmax;
        const int64_t cap = Pmax < pl->n_col ? Pmax : pl->n_col;
        int64_t Hmax = 64;
        while (Hmax < 2 * cap) Hmax <<= 1;
        const size_t nwords_max = (size_t)((Pmax + 31) / 32 + 2);
        size_t per_cta = (((size_t)Hmax * sizeof(I) + 15) & ~(size_t)15) + (size_t)Hmax * 8 + (size_t)Hmax * 8 +
                         ((nwords_max * 4 + 15) & ~(size_t)15) + nwords_max * 8 + 64;
        per_cta = (per_cta + 255) & ~(size_t)255;
        int64_t ctas = pl->n_long < (int64_t)sms * 2 ? pl->n_long : (int64_t)sms * 2;
        const size_t budget = (size_t)8 << 30;
        while (ctas > 1 && (size_t)ctas * per_cta > budget) ctas /= 2;
        unsigned char *scratch = nullptr;
        if ((rc = scratch_alloc((void **)&scratch, (size_t)ctas * per_cta, s))) return rc;
        spgemm_block_kernel<T, W, I, WARPS, TILE><<<(unsigned)ctas, WARPS * 32, 0, s>>>(
            pl->long_rows, pl->n_long, pl->n_col, ap, ai, ad, bp, bi, bd, pl->P, pl->side_off, side_idx, side_val,
            long_nnz, long_nz, scratch, per_cta, Hmax, Pmax);
        B2S_CHECK_LAUNCH();
        if (pl->sorted) {
            int64_t *sc_idx = nullptr;
            T *sc_val = nullptr;
            if ((rc = scratch_alloc((void **)&sc_idx, (size_t)pl->side_total * 8, s))) return rc;
            if ((rc = scratch_alloc((void **)&sc_val, (size_t)pl->side_total * sizeof(T), s))) return rc;
            sort_long_rows_kernel<T><<<(unsigned)ctas, 256, 0, s>>>(pl->long_rows, pl->n_long, pl->side_off, long_nnz,
                                                                   side_idx, side_val, sc_idx, sc_val);
            B2S_CHECK_LAUNCH();
            scratch_free(sc_idx, s);
            scratch_free(sc_val, s);
        }
        scratch_free(scratch, s);
    }
    // 3. ordered single pass: final indptr / indices / data (/ COO rows)
    int64_t *own_ptr = nullptr;
    int64_t *ptr = indptr_out;
    if (!ptr) {
        if ((rc = scratch_alloc((void **)&own_ptr, (size_t)(M + 1) * 8, s))) return rc;
        ptr = own_ptr;
    }
    constexpr int WARPS_O = 8;
    const int64_t n_tiles = (M + WARPS_O - 1) / WARPS_O;
    unsigned char *look = nullptr;
    if ((rc = scratch_alloc((void **)&look, 16 + (size_t)n_tiles * 8, s))) return rc;
    B2S_CUDA(cudaMemsetAsync(look, 0, 16 + (size_t)n_tiles * 8, s));  // ticket and every status word start at 0
    rc = launch_ordered<T, W, I, 512, WARPS_O>(pl, prune, side_idx, side_val, long_nnz, long_nz,
                                               (uint64_t *)(look + 16), (unsigned int *)look, ptr, indices_out, rows_out,
                                               (T *)data_out);
    if (rc) return rc;
    unsigned long long tot[2];
    B2S_CUDA(cudaMemcpyAsync(tot, pl->counters + 4, 16, cudaMemcpyDeviceToHost, s));
    B2S_CUDA(cudaStreamSynchronize(s));
    scratch_free(look, s);
    if (side_idx) scratch_free(side_idx, s);
    if (side_val) scratch_free(side_val, s);
    if (long_nnz) scratch_free(long_nnz, s);
    if (long_nz) scratch_free(long_nz, s);
    *nnz_struct_out = (int64_t)tot[0];
    *nnz_out = (int64_t)tot[1];
    // 4. the completely dense result comes out with every row reversed (_common.py:709-714)
    if (!pl->sorted && pl->n_col > 0 && (int64_t)tot[0] == M * pl->n_col && tot[1] > 0) {
        int64_t blocks = (M * 32 + 255) / 256;
        if (blocks > (int64_t)num_sms() * 16) blocks = (int64_t)num_sms() * 16;
        spgemm_reverse_rows_kernel<T><<<(unsigned)blocks, 256, 0, s>>>(M, ptr, indices_out, (T *)data_out);
        B2S_CHECK_LAUNCH();
    }
    if (own_ptr) scratch_free(own_ptr, s);
    return B2S_OK;
}

static void plan_free(SpgemmPlan *pl) {
    cudaStream_t s = pl->stream;
    scratch_free(pl->P, s);
    scratch_free(pl->long_rows, s);
    scratch_free(pl->side_off, s);
    scratch_free(pl->counters, s);
    delete pl;
}

}  // namespace b2s

using namespace b2s;

extern "C" {

int b2s_spgemm_set_thresholds(int64_t t0, int64_t t1) {
    // historical two-threshold form: the larger one is the product count above which a row takes the CTA-per-row path
    (void)t0;
    g_pmax_short = (t1 >= 1 && t1 <= 256) ? t1 : 256;
    return B2S_OK;
}

int b2s_spgemm_begin(int dtype, int idx_bytes, int64_t M, int64_t K, int64_t n_col, const void *a_indptr_dev,
                     const void *a_indices_dev, const void *a_data_dev, const void *b_indptr_dev,
                     const void *b_indices_dev, const void *b_data_dev, int sorted_order, int wide_accumulate,
                     void **plan_out, int64_t *capacity_out, void *stream) {
    B2S_REQUIRE(plan_out && capacity_out, B2S_ERR_INVALID, "spgemm_begin: NULL output");
    B2S_REQUIRE(M >= 0 && K >= 0 && n_col >= 0, B2S_ERR_INVALID, "spgemm_begin: negative dimension");
    B2S_REQUIRE(idx_bytes == 4 || idx_bytes == 8, B2S_ERR_INVALID, "spgemm_begin: idx_bytes");
    B2S_REQUIRE(dtype == B2S_F32 || dtype == B2S_F64 || dtype == B2S_I32 || dtype == B2S_I64, B2S_ERR_UNSUPPORTED,
                "spgemm: dtype %d", dtype);
    cudaStream_t s = (cudaStream_t)stream;
    SpgemmPlan *pl = new SpgemmPlan();
    memset(pl, 0, sizeof(*pl));
    pl->dtype = dtype;
    pl->idx_bytes = idx_bytes;
    pl->sorted = sorted_order ? 1 : 0;
    pl->wide = wide_accumulate ? 1 : 0;
    pl->M = M;
    pl->n_col = n_col;
    pl->stream = s;
    pl->a_indptr = a_indptr_dev;
    pl->a_indices = a_indices_dev;
    pl->a_data = a_data_dev;
    pl->b_indptr = b_indptr_dev;
    pl->b_indices = b_indices_dev;
    pl->b_data = b_data_dev;
    int rc = B2S_OK;
    const size_t mb = (size_t)(M + 1) * 8;
    if ((rc = scratch_alloc((void **)&pl->P, mb, s)) || (rc = scratch_alloc((void **)&pl->long_rows, mb, s)) ||
        (rc = scratch_alloc((void **)&pl->side_off, mb, s)) || (rc = scratch_alloc((void **)&pl->counters, 64, s))) {
        plan_free(pl);
        return rc;
    }
    if (M > 0) {
        rc = idx_bytes == 4 ? spgemm_products<float, float, int32_t>(pl) : spgemm_products<float, float, int64_t>(pl);
    }
    if (rc != B2S_OK) {
        plan_free(pl);
        return rc;
    }
    *plan_out = pl;
    *capacity_out = pl->ub_total;
    return B2S_OK;
}

int b2s_spgemm_run(void *plan, int prune, int64_t *indptr_out_dev, int64_t *indices_out_dev, int64_t *rows_out_dev,
                   void *data_out_dev, int64_t *nnz_struct_out, int64_t *nnz_out) {
    B2S_REQUIRE(plan != nullptr && nnz_struct_out && nnz_out, B2S_ERR_INVALID, "spgemm_run: NULL argument");
    SpgemmPlan *pl = (SpgemmPlan *)plan;
    int rc = B2S_OK;
    *nnz_struct_out = 0;
    *nnz_out = 0;
    if (pl->M > 0) {
#define B2S_RUN(T, I)                                                                                             \
    rc = pl->wide ? spgemm_run_t<T, double, I>(pl, prune, indptr_out_dev, indices_out_dev, rows_out_dev,           \
                                               data_out_dev, nnz_struct_out, nnz_out)                             \
                  : spgemm_run_t<T, T, I>(pl, prune, indptr_out_dev, indices_out_dev, rows_out_dev, data_out_dev,  \
                                          nnz_struct_out, nnz_out)
        if (pl->idx_bytes == 4) {
            switch (pl->dtype) {
                case B2S_F32: B2S_RUN(float, int32_t); break;
                case B2S_F64: B2S_RUN(double, int32_t); break;
                case B2S_I32: B2S_RUN(int32_t, int32_t); break;
                default: B2S_RUN(int64_t, int32_t); break;
            }
        } else {
            switch (pl->dtype) {
                case B2S_F32: B2S_RUN(float, int64_t); break;
                case B2S_F64: B2S_RUN(double, int64_t); break;
                case B2S_I32: B2S_RUN(int32_t, int64_t); break;
                default: B2S_RUN(int64_t, int64_t); break;
            }
        }
#undef B2S_RUN
    } else if (indptr_out_dev) {
        cudaMemsetAsync(indptr_out_dev, 0, 8, pl->stream);
    }
    plan_free(pl);
    return rc;
}

int b2s_spgemm_abort(void *plan) {
    if (plan) plan_free((SpgemmPlan *)plan);
    return B2S_OK;
}

}  // extern "C"
