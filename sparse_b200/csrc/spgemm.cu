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
//   1. products: P_i = sum over A-row entries of the B-row lengths (8 lanes per row), U_i = min(P_i, n_col) bounds
//      nnz_i; an exclusive scan of U gives every row its place in an upper-bound layout; rows with P_i > 256 are listed.
//   2. (rare) long rows: CTA-per-row kernel with a global-memory hash.
//   3. row kernel: one warp = one row at a time (rows dealt out four at a time by a ticket), software-pipelined three
//      rows deep: the ticket, the A row and the B row extents of the following rows are loaded while the current row
//      is multiplied, and all B entries of a row are fetched in one round trip before the inserts start.  The row is
//      built in a shared-memory hash table sized for it (64 .. 512 slots, load <= 1/2): products are inserted 32 at a
//      time in the reference's visiting order, equal columns inside a chunk are grouped by __match_any_sync so their
//      adds stay sequential, slots are claimed WITHOUT atomics (store, __syncwarp, re-read: the table belongs to one
//      warp), and the n-th first touch records its slot in ord[n] -- so the row is written in REVERSE first-touch
//      order (REF; exactly the reference's intrusive linked list) or ranked by column (SORTED), coalesced, at its
//      upper-bound offset.
//   4. finish: scan of the row counts -> indptr; per-row compaction from the upper-bound layout to the final CSR/COO
//      arrays, optionally dropping values bitwise equal to +0 (the prune=True of _common.py:374-379) and reversing
//      rows in the all-dense case (:709-714).
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
// cheaper variant for the shared-memory tables (<= 2^12 slots): 32-bit multiplicative (Fibonacci) hashing, upper bits
template <typename I>
__device__ __forceinline__ unsigned hash_small(I k) {
    unsigned x = (unsigned)k;
    if constexpr (sizeof(I) == 8) x ^= (unsigned)((uint64_t)k >> 32) * 0x85EBCA6Bu;
    return (x * 0x9E3779B1u) >> 20;
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
// counters: [0] sum of U over all rows, [1] number of long rows, [3] max P of a long row
template <typename I>
__global__ void __launch_bounds__(256)
spgemm_products_kernel(int64_t M, int64_t n_col, int64_t pmax_short, const I *__restrict__ a_indptr,
                       const I *__restrict__ a_indices, const I *__restrict__ b_indptr, int64_t *__restrict__ P,
                       int64_t *__restrict__ U, int64_t *__restrict__ long_rows,
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
        U[row] = u;
        if (acc > pmax_short) {
            const unsigned long long at = atomicAdd(counters + 1, 1ull);
            long_rows[at] = row;
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
// 3. warp-per-row numeric kernel (see the header comment)
// ---------------------------------------------------------------------------------------------
template <typename T, typename W, typename I, int HMAX, int WARPS>
struct OrderedSmem {
    static constexpr int PMAX = HMAX / 2;
    // per-warp carve-up (bytes), 16-byte aligned pieces: the hash table (key, sum) and the first-touch order
    static constexpr size_t key_b = ((size_t)HMAX * sizeof(I) + 15) & ~(size_t)15;
    static constexpr size_t sum_b = ((size_t)HMAX * sizeof(W) + 15) & ~(size_t)15;
    static constexpr size_t ord_b = ((size_t)PMAX * 2 + 15) & ~(size_t)15;
    static constexpr size_t per_warp = key_b + sum_b + ord_b;
    static constexpr size_t total = per_warp * WARPS;
    // CTAs per SM the shared-memory footprint allows (227 KB per SM, 1 KB reserved per CTA), at most 8
    static constexpr int ctas_by_smem = (int)((227 * 1024) / (total + 1024));
    static constexpr int min_ctas = ctas_by_smem < 1 ? 1 : (ctas_by_smem > 8 ? 8 : ctas_by_smem);
};

// One warp = one row at a time, rows dealt out four at a time by a global ticket.  Three rows ahead of the one being
// multiplied the warp already holds the row number, two rows ahead the A row (indices, values), one row ahead the B row
// extents: the dependent loads a_indptr -> a_indices -> b_indptr of a row are issued one row-time before their results
// are used.
template <typename T, typename W, typename I, int HMAX, int WARPS, bool SORTED>
__global__ void __launch_bounds__(WARPS * 32, (OrderedSmem<T, W, I, HMAX, WARPS>::min_ctas))
spgemm_rows_kernel(int64_t M, const I *__restrict__ a_indptr, const I *__restrict__ a_indices,
                   const T *__restrict__ a_data, const I *__restrict__ b_indptr, const I *__restrict__ b_indices,
                   const T *__restrict__ b_data, const int64_t *__restrict__ Pv, const int64_t *__restrict__ ub_off,
                   int64_t pmax_short, unsigned int *__restrict__ ticket, I *__restrict__ tmp_idx,
                   T *__restrict__ tmp_val, int64_t *__restrict__ row_nnz, int64_t *__restrict__ row_nz,
                   unsigned long long *__restrict__ totals /* [0] structural entries, [1] entries != +0 */) {
    using L = OrderedSmem<T, W, I, HMAX, WARPS>;
    constexpr int PMAX = L::PMAX;
    constexpr I EMPTY = Empty<I>::value;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    unsigned char *mine = smem_raw + (size_t)w * L::per_warp;
    I *key = reinterpret_cast<I *>(mine);
    W *sum = reinterpret_cast<W *>(mine + L::key_b);
    uint16_t *ord = reinterpret_cast<uint16_t *>(mine + L::key_b + L::sum_b);

    unsigned long long my_struct = 0, my_nz = 0;  // totals over the rows this warp handled
    int distinct = 0;                             // structural entries of the row in the table

    // ---- one chunk of <= 32 products (visiting order = lane order) goes into the table ----------------------
    auto insert = [&](I k, T p, bool active, unsigned mask) {
        const unsigned amask = __ballot_sync(FULL, active);
        unsigned grp = 0;
        if (active) grp = __match_any_sync(amask, k);
        const bool lead = active && (lane == __ffs(grp) - 1);
        const int gsz = __popc(grp);
        // claim / find the slot -- no atomics: the table belongs to this warp; a lane that sees an empty slot stores
        // its key, everybody synchronises, and the lane whose key is there owns the slot
        unsigned h = hash_small<I>(k) & mask;
        bool pending = lead, isnew = false;
        while (__any_sync(FULL, pending)) {
            I cur = EMPTY;
            if (pending) cur = key[h];
            const bool empty = pending && cur == EMPTY;
            __syncwarp();
            // Every lane that found this slot empty stores its key with ONE warp-wide store instruction.  When several
            // lanes target the same slot, the hardware serialises their stores and exactly one key ends up there (CUDA
            // programming guide, "if a non-atomic instruction executed by a warp writes to the same location ... for more
            // than one of the threads of the warp, ... one of the writes is guaranteed to succeed"); the re-read below
            // tells every lane whether it was the one.  compute-sanitizer racecheck reports this store as a WAW hazard
            // (profiles/r02_compute_sanitizer_racecheck.log): it is the election itself.  Electing the writer with
            // __match_any_sync instead costs 25 % of the whole product (2.13 -> 2.66 ms at C5, measured).
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
    };

    // ---- the products of one A chunk (<= 32 entries; lane l holds entry l: B row start, length, A value) ------------
    // exclusive offsets of the entries' product ranges (lanes past the row's entries hold `total`)
    auto scan_lens = [&](int len, int &excl, int &total) {
        int incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += v;
        }
        total = __shfl_sync(FULL, incl, 31);
        excl = incl - len;
    };
    // issue the loads of product chunk c (column, B value) and pick its A value; NOTHING here waits for the loads: the
    // product is formed by the consumer one chunk-time (or one row-time) later
    auto fetch = [&](int c, int excl, int total, I bs, T av, I &k, T &bv, T &a0) {
        k = EMPTY;
        bv = T(0);
        const int t = c * 32 + lane;
        int lo = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const int e = __shfl_sync(FULL, excl, lo + step);
            if (e <= t) lo += step;
        }
        const I b0 = __shfl_sync(FULL, bs, lo);
        a0 = __shfl_sync(FULL, av, lo);
        const int e0 = __shfl_sync(FULL, excl, lo);
        if (t < total) {
            const int64_t src = (int64_t)b0 + (t - e0);
            k = b_indices[src];
            bv = b_data[src];
        }
    };
    // software pipeline over the chunks of 32 products: the loads of chunk c + 1 are in flight while chunk c is
    // inserted; chunk 0 arrives already fetched (kn, bn, an)
    auto run_chunks = [&](int excl, int total, I bs, T av, I kn, T bn, T an, unsigned mask) {
#pragma unroll 1
        for (int c = 0; c * 32 < total; ++c) {
            const I k = kn;
            const T p = mul_rn(an, bn);
            if ((c + 1) * 32 < total) fetch(c + 1, excl, total, bs, av, kn, bn, an);
            insert(k, p, c * 32 + lane < total, mask);
        }
    };

    // ---- row pipeline ---------------------------------------------------------------------------------------------
    // (row numbers fit an int: M < 2^31 is checked on the host; offsets into A are as wide as its index type)
    using off_t = typename std::conditional<sizeof(I) == 4, int, int64_t>::type;
    constexpr int ROWS_PER_TICKET = 4;
    bool exhausted = false;
    int gen_next = 0, gen_end = 0;
    const int Mi = (int)M;
    auto next_row = [&]() -> int {
        if (gen_next == gen_end) {
            if (exhausted) return -1;
            unsigned t = 0;
            if (lane == 0) t = atomicAdd(ticket, 1u);
            t = __shfl_sync(FULL, t, 0);
            if (t >= (unsigned)((Mi + ROWS_PER_TICKET - 1) / ROWS_PER_TICKET)) {
                exhausted = true;
                return -1;
            }
            gen_next = (int)t * ROWS_PER_TICKET;
            gen_end = gen_next + ROWS_PER_TICKET < Mi ? gen_next + ROWS_PER_TICKET : Mi;
        }
        return gen_next++;
    };
    // stage 3 (ticket + row extent), stage 2 (+ A entries), stage 1 (+ B extents); row < 0 = empty slot
    int row3 = -1, row2 = -1, row1 = -1;
    off_t as3 = 0, ae3 = 0, as2 = 0, ae2 = 0, as1 = 0, ae1 = 0;
    I j2 = 0;
    T av2 = T(0);
    I bs1 = 0, be1 = 0;  // B row extent of this lane's A entry (raw loads: nothing is computed from them here)
    T av1 = T(0);
    // chunk 0 of the NEXT row, fetched before the current row is written out
    bool pre_ok = false;
    int pre_excl = 0, pre_total = 0;
    I pre_k = EMPTY;
    T pre_b = T(0), pre_a = T(0);

    while (true) {
        // rotate the pipeline: the loads issued here are consumed one iteration later
        const int row = row1;
        const off_t as = as1, ae = ae1;
        const I bs = bs1;
        const int len = (int)(be1 - bs1);
        const T av = av1;
        row1 = row2, as1 = as2, ae1 = ae2, av1 = av2;
        bs1 = 0, be1 = 0;
        if (row1 >= 0 && as1 + lane < ae1) {
            bs1 = b_indptr[j2];
            be1 = b_indptr[j2 + 1];
        }
        row2 = row3, as2 = as3, ae2 = ae3;
        j2 = 0, av2 = T(0);
        if (row2 >= 0 && as2 + lane < ae2) {
            j2 = a_indices[as2 + lane];
            av2 = a_data[as2 + lane];
        }
        row3 = next_row();
        if (row3 >= 0) {
            as3 = (off_t)a_indptr[row3];
            ae3 = (off_t)a_indptr[row3 + 1];
        }
        if (row < 0) {
            if (row1 < 0 && row2 < 0 && row3 < 0) break;
            continue;
        }

        // ---- multiply row `row` -------------------------------------------------------------------------------------
        const bool one_chunk = ae - as <= 32;
        int excl = 0, total = 0;
        I kn = EMPTY;
        T bn = T(0), an = T(0);
        int64_t P;
        if (one_chunk) {
            if (pre_ok) {  // scanned and fetched while the previous row was being written
                excl = pre_excl, total = pre_total, kn = pre_k, bn = pre_b, an = pre_a;
            } else {
                scan_lens(len, excl, total);
                if (total > 0 && total <= pmax_short) fetch(0, excl, total, bs, av, kn, bn, an);
            }
            P = total;
        } else {
            P = Pv[row];
        }
        pre_ok = false;
        const bool skip_row = P == 0 || P > pmax_short;  // long rows belong to the CTA-per-row kernel (and its counts)
        if (P == 0 && lane == 0) row_nnz[row] = 0, row_nz[row] = 0;
        distinct = 0;
        if (!skip_row) {
            int H = 64;
            while (H < 4 * (int)P && H < HMAX) H <<= 1;  // load <= 1/4 when the table has room, <= 1/2 always
            const unsigned mask = (unsigned)(H - 1);
            for (int s = lane; s < H; s += 32) key[s] = EMPTY;
            __syncwarp();
            if (one_chunk) {
                run_chunks(excl, total, bs, av, kn, bn, an, mask);
            } else {
                for (off_t ab = as; ab < ae; ab += 32) {
                    I cbs = 0;
                    int clen = 0;
                    T cav = T(0);
                    if (ab + lane < ae) {
                        const I j = a_indices[ab + lane];
                        cav = a_data[ab + lane];
                        cbs = b_indptr[j];
                        clen = (int)(b_indptr[j + 1] - cbs);
                    }
                    scan_lens(clen, excl, total);
                    if (total > 0) {
                        fetch(0, excl, total, cbs, cav, kn, bn, an);
                        run_chunks(excl, total, cbs, cav, kn, bn, an, mask);
                    }
                }
            }
        }
        // the next row's first chunk: its B extents landed one row-time ago -- scan them and issue the loads now, so
        // they are in flight while this row is written out
        if (row1 >= 0 && ae1 - as1 <= 32) {
            scan_lens((int)(be1 - bs1), pre_excl, pre_total);
            if (pre_total > 0 && pre_total <= pmax_short) fetch(0, pre_excl, pre_total, bs1, av1, pre_k, pre_b, pre_a);
            pre_ok = true;
        }
        if (skip_row) continue;
        // ---- write the row at its upper-bound offset, in its final order ---------------------------------------------
        const int64_t ub = ub_off[row];
        int nz = 0;
        if constexpr (!SORTED) {
            for (int i = lane; i < distinct; i += 32) {
                const int slot = ord[distinct - 1 - i];  // reverse first-touch order
                const T v = narrow_sum<T, W>(sum[slot]);
                tmp_idx[ub + i] = key[slot];
                tmp_val[ub + i] = v;
                nz += is_pos_zero_bits(v) ? 0 : 1;
            }
        } else {
            for (int e = lane; e < distinct; e += 32) {  // ascending columns: rank by counting
                const int slot = ord[e];
                const I k = key[slot];
                int rank = 0;
                for (int f = 0; f < distinct; ++f) rank += (key[ord[f]] < k) ? 1 : 0;
                const T v = narrow_sum<T, W>(sum[slot]);
                tmp_idx[ub + rank] = k;
                tmp_val[ub + rank] = v;
                nz += is_pos_zero_bits(v) ? 0 : 1;
            }
        }
        nz = __reduce_add_sync(FULL, nz);
        if (lane == 0) {
            row_nnz[row] = distinct;
            row_nz[row] = nz;
        }
        my_struct += (unsigned long long)distinct;
        my_nz += (unsigned long long)nz;
        __syncwarp();  // the table is reused by the next row
    }
    if (lane == 0 && my_struct) {
        atomicAdd(totals + 0, my_struct);
        atomicAdd(totals + 1, my_nz);
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
                    I *__restrict__ tmp_idx, T *__restrict__ tmp_val, int64_t *__restrict__ row_nnz,
                    int64_t *__restrict__ row_nz, unsigned char *__restrict__ scratch, size_t per_cta,
                    int64_t Hmax, int64_t Pmax, unsigned long long *__restrict__ totals) {
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
                tmp_idx[ub + above] = k;
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
            atomicAdd(totals + 0, (unsigned long long)distinct);
            atomicAdd(totals + 1, (unsigned long long)tot);
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// 4. finish: compaction from the upper-bound layout (+ optional prune, row reversal, COO rows)
// ---------------------------------------------------------------------------------------------
template <typename T, typename I>
__global__ void spgemm_finish_kernel(int64_t M, const int64_t *__restrict__ ub_off,
                                     const int64_t *__restrict__ row_nnz, const int64_t *__restrict__ out_ptr,
                                     const I *__restrict__ tmp_idx, const T *__restrict__ tmp_val, int prune,
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
        // 128 entries at a time: all loads of the group are issued before the first store (4 x the bytes in flight
        // of a chunk-by-chunk copy; the kernel is a pure stream compaction)
        for (int64_t c = 0; c < n; c += 128) {
            int64_t k[4];
            T v[4];
            bool keep[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t p = c + u * 32 + lane;
                keep[u] = false;
                k[u] = 0;
                v[u] = T(0);
                if (p < n) {
                    const int64_t q = reverse ? (n - 1 - p) : p;
                    k[u] = (int64_t)tmp_idx[src + q];
                    v[u] = tmp_val[src + q];
                    keep[u] = true;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (c + u * 32 >= n) break;
                const bool kp = keep[u] && (!prune || !is_pos_zero_bits(v[u]));
                const unsigned m = __ballot_sync(FULL, kp);
                if (kp) {
                    const int64_t o = dst + __popc(m & ((1u << lane) - 1));
                    out_idx[o] = k[u];
                    out_val[o] = v[u];
                    if (out_rows) out_rows[o] = row;
                }
                dst += __popc(m);
            }
        }
    }
}

// per-row ascending sort of (idx, val) segments for rows that went through the block kernel in SORTED mode
// (rare, long rows): one CTA per row, odd-even transposition over global memory is too slow, so use a
// bitonic-free approach: rank by counting inside the CTA (O(n^2 / threads)); rows here have n <= n_col.
template <typename T, typename I>
__global__ void sort_long_rows_kernel(const int64_t *__restrict__ rows, int64_t n_rows,
                                      const int64_t *__restrict__ ub_off, const int64_t *__restrict__ row_nnz,
                                      I *__restrict__ tmp_idx, T *__restrict__ tmp_val,
                                      I *__restrict__ sc_idx, T *__restrict__ sc_val) {
    for (int64_t ri = blockIdx.x; ri < n_rows; ri += gridDim.x) {
        const int64_t row = rows[ri];
        const int64_t n = row_nnz[row];
        const int64_t off = ub_off[row];
        for (int64_t e = threadIdx.x; e < n; e += blockDim.x) {
            const I k = tmp_idx[off + e];
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
    int64_t *P, *U, *ub_off, *row_nnz, *row_nz, *long_rows;  // [M] (+1 for offsets)
    unsigned long long *counters;                            // [4] products kernel, [4..5] totals, [6] ticket
    void *tmp_idx;  // column indices of the upper-bound layout, in the operands' index width
    int idx_bytes;
    void *tmp_val;
    cudaStream_t stream;
};

// rows with more products than this go through the CTA-per-row kernel (tests may lower it to exercise that path)
static int64_t g_pmax_short = 256;

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
    // 1. products per row, upper bounds, long-row list; the scan places every row in the upper-bound layout
    B2S_CUDA(cudaMemsetAsync(pl->counters, 0, 8 * 8, s));
    {
        const int64_t threads = M * 8;
        spgemm_products_kernel<I><<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(
            M, pl->n_col, g_pmax_short, ap, ai, bp, pl->P, pl->U, pl->long_rows, pl->counters);
        B2S_CHECK_LAUNCH();
    }
    if ((rc = exclusive_scan_i64(pl->U, pl->ub_off, M + 1, s))) return rc;  // U[M] = 0 sentinel
    unsigned long long hc[4];
    B2S_CUDA(cudaMemcpyAsync(hc, pl->counters, 32, cudaMemcpyDeviceToHost, s));
    B2S_CUDA(cudaStreamSynchronize(s));
    pl->ub_total = (int64_t)hc[0];
    const int64_t n_long = (int64_t)hc[1];
    if ((rc = scratch_alloc((void **)&pl->tmp_idx, (size_t)pl->ub_total * sizeof(I), s))) return rc;
    if ((rc = scratch_alloc((void **)&pl->tmp_val, (size_t)pl->ub_total * sizeof(T), s))) return rc;
    const int sms = num_sms();
    unsigned long long *totals = pl->counters + 4;
    // 2. long rows (rare): CTA per row, global-memory hash
    if (n_long) {
        constexpr int WARPS = 8;
        constexpr int TILE = 2048;
        const int64_t Pmax = (int64_t)hc[3];
        const int64_t cap = Pmax < pl->n_col ? Pmax : pl->n_col;
        int64_t Hmax = 64;
        while (Hmax < 2 * cap) Hmax <<= 1;
        const size_t nwords_max = (size_t)((Pmax + 31) / 32 + 2);
        size_t per_cta = (((size_t)Hmax * sizeof(I) + 15) & ~(size_t)15) + (size_t)Hmax * 8 + (size_t)Hmax * 8 +
                         ((nwords_max * 4 + 15) & ~(size_t)15) + nwords_max * 8 + 64;
        per_cta = (per_cta + 255) & ~(size_t)255;
        int64_t ctas = n_long < (int64_t)sms * 2 ? n_long : (int64_t)sms * 2;
        const size_t budget = (size_t)8 << 30;
        while (ctas > 1 && (size_t)ctas * per_cta > budget) ctas /= 2;
        unsigned char *scratch = nullptr;
        if ((rc = scratch_alloc((void **)&scratch, (size_t)ctas * per_cta, s))) return rc;
        spgemm_block_kernel<T, W, I, WARPS, TILE><<<(unsigned)ctas, WARPS * 32, 0, s>>>(
            pl->long_rows, n_long, pl->n_col, ap, ai, ad, bp, bi, bd, pl->P, pl->ub_off, (I *)pl->tmp_idx,
            (T *)pl->tmp_val, pl->row_nnz, pl->row_nz, scratch, per_cta, Hmax, Pmax, totals);
        B2S_CHECK_LAUNCH();
        if (pl->sorted) {
            I *sc_idx = nullptr;
            T *sc_val = nullptr;
            if ((rc = scratch_alloc((void **)&sc_idx, (size_t)pl->ub_total * sizeof(I), s))) return rc;
            if ((rc = scratch_alloc((void **)&sc_val, (size_t)pl->ub_total * sizeof(T), s))) return rc;
            sort_long_rows_kernel<T, I><<<(unsigned)ctas, 256, 0, s>>>(pl->long_rows, n_long, pl->ub_off, pl->row_nnz,
                                                                      (I *)pl->tmp_idx, (T *)pl->tmp_val, sc_idx, sc_val);
            B2S_CHECK_LAUNCH();
            scratch_free(sc_idx, s);
            scratch_free(sc_val, s);
        }
        scratch_free(scratch, s);
    }
    // 3. every other row: warp per row, persistent grid sized by the occupancy the shared-memory footprint allows
    {
        constexpr int WARPS = 4;
        using L = OrderedSmem<T, W, I, 512, WARPS>;
        auto kern_r = spgemm_rows_kernel<T, W, I, 512, WARPS, false>;
        auto kern_s = spgemm_rows_kernel<T, W, I, 512, WARPS, true>;
        auto kern = pl->sorted ? kern_s : kern_r;
        B2S_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::total));
        int occ = 1;
        B2S_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WARPS * 32, L::total));
        if (occ < 1) occ = 1;
        int64_t blocks = (int64_t)sms * occ;
        const int64_t need = (M + WARPS - 1) / WARPS;
        if (blocks > need) blocks = need;
        kern<<<(unsigned)blocks, WARPS * 32, L::total, s>>>(M, ap, ai, ad, bp, bi, bd, pl->P, pl->ub_off,
                                                            g_pmax_short < 256 ? g_pmax_short : (int64_t)256,
                                                            (unsigned int *)(pl->counters + 6), (I *)pl->tmp_idx,
                                                            (T *)pl->tmp_val, pl->row_nnz, pl->row_nz, totals);
        B2S_CHECK_LAUNCH();
    }
    unsigned long long h[2];
    B2S_CUDA(cudaMemcpyAsync(h, totals, 16, cudaMemcpyDeviceToHost, s));
    B2S_CUDA(cudaStreamSynchronize(s));
    pl->nnz_struct = (int64_t)h[0];
    pl->nnz_pruned = (int64_t)h[1];
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
        if (pl->idx_bytes == 4)
            spgemm_finish_kernel<T, int32_t><<<(unsigned)blocks, 256, 0, s>>>(
                M, pl->ub_off, pl->row_nnz, ptr, (const int32_t *)pl->tmp_idx, (const T *)pl->tmp_val, prune, reverse,
                indices_out, rows_out, (T *)data_out);
        else
            spgemm_finish_kernel<T, int64_t><<<(unsigned)blocks, 256, 0, s>>>(
                M, pl->ub_off, pl->row_nnz, ptr, (const int64_t *)pl->tmp_idx, (const T *)pl->tmp_val, prune, reverse,
                indices_out, rows_out, (T *)data_out);
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
    scratch_free(pl->long_rows, s);
    scratch_free(pl->counters, s);
    scratch_free(pl->tmp_idx, s);
    scratch_free(pl->tmp_val, s);
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
    pl->idx_bytes = idx_bytes;
    pl->sorted = sorted_order ? 1 : 0;
    const bool wide = wide_accumulate != 0;
    pl->M = M;
    pl->n_col = n_col;
    pl->stream = s;
    int rc = B2S_OK;
    const size_t mb = (size_t)(M + 1) * 8;
    if ((rc = scratch_alloc((void **)&pl->P, mb, s)) || (rc = scratch_alloc((void **)&pl->U, mb, s)) ||
        (rc = scratch_alloc((void **)&pl->ub_off, mb, s)) || (rc = scratch_alloc((void **)&pl->row_nnz, mb, s)) ||
        (rc = scratch_alloc((void **)&pl->row_nz, mb, s)) || (rc = scratch_alloc((void **)&pl->long_rows, mb, s)) ||
        (rc = scratch_alloc((void **)&pl->counters, 64, s))) {
        plan_free(pl);
        return rc;
    }
    cudaMemsetAsync(pl->U + M, 0, 8, s);        // sentinel of the scans
    cudaMemsetAsync(pl->row_nnz + M, 0, 8, s);
    cudaMemsetAsync(pl->row_nz + M, 0, 8, s);
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
