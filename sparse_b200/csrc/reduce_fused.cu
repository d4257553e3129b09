// reduce_fused.cu -- K7 production path: hand-written segmented-scan reduction over sorted COO keys.
//
// Replaces _calc_counts_invidx + ufunc.reduceat + the fill-value correction + the result's coordinate build
// (sparse/numba_backend/_coo/core.py:1601-1661, 693-723; _sparse_array.py:405-422) with two streaming passes:
//
//   pass 1  per 2048-element tile: head flags (group id = key / n_cols changes), a segmented reduction inside
//           every thread's 8 consecutive elements and a block-level segmented scan -> tile summary
//           (number of heads, "contains a head", value and count of the run that is still open at the tile end);
//   scan    one CTA turns the tile summaries into per-tile head offsets and carry-in runs (segmented scan);
//   pass 2  recomputes the tile, seeds the block scan with the carry-in and lets the thread that owns the LAST element
//           of every run write its group id, its value with the fill-value contribution already applied
//           (add: + fill*n_fill, multiply: * fill**n_fill, others: op(v, fill) when the group is incomplete)
//           and the unravelled coordinates of the kept axes.  It also counts results equal to the result fill value
//           so the (rare) prune compaction only runs when needed.
//
// Traffic: 2 x n x (8 + sizeof(T)) bytes read, n_groups x (8 x (1 + ndim_out) + sizeof(T)) written.
// Association order differs from NumPy's reduceat (itself unspecified) -> tolerance parity for float add/multiply.
#include <type_traits>

#include "common.cuh"
#include "fastdiv.cuh"

namespace b2s {

enum RedOp2 { RF_ADD = 0, RF_MUL = 1, RF_MAX = 2, RF_MIN = 3, RF_AND = 4, RF_OR = 5, RF_BAND = 6, RF_BOR = 7, RF_BXOR = 8,
              RF_FMAX = 9, RF_FMIN = 10 };  // NaN-ignoring (np.fmax / np.fmin: nanmax / nanmin)

constexpr int RD_THREADS = 256;
constexpr int RD_ITEMS = 8;
constexpr int RD_TILE = RD_THREADS * RD_ITEMS;
constexpr int RD_MAXDIM = 16;

template <typename T>
__device__ __forceinline__ T red_apply(int op, T a, T b) {
    if constexpr (std::is_floating_point<T>::value) {
        switch (op) {
            case RF_ADD: return add_rn(a, b);
            case RF_MUL: return mul_rn(a, b);
            case RF_MAX: return (a != a) ? a : ((b != b) ? b : (a >= b ? a : b));  // NaN propagates (np.maximum)
            case RF_MIN: return (a != a) ? a : ((b != b) ? b : (a <= b ? a : b));
            case RF_FMAX: return (a >= b || b != b) ? a : b;  // the other operand when one is NaN
            case RF_FMIN: return (a <= b || b != b) ? a : b;
            default: return a;
        }
    } else {
        switch (op) {
            case RF_ADD: return (T)(a + b);
            case RF_MUL: return (T)(a * b);
            case RF_MAX: case RF_FMAX: return a >= b ? a : b;
            case RF_MIN: case RF_FMIN: return a <= b ? a : b;
            case RF_AND: return (T)((a != T(0)) && (b != T(0)));
            case RF_OR: return (T)((a != T(0)) || (b != T(0)));
            case RF_BAND: return (T)(a & b);
            case RF_BOR: return (T)(a | b);
            case RF_BXOR: return (T)(a ^ b);
            default: return a;
        }
    }
}

// fill-value contribution of SparseArray.reduce (_sparse_array.py:405-422)
template <typename T>
__device__ __forceinline__ T fill_fix(int op, T v, int64_t count, int64_t ncols, T fill) {
    const int64_t nf = ncols - count;
    if (op == RF_ADD) {
        T contrib = T(0);
        if (nf != 0) {
            if constexpr (std::is_floating_point<T>::value) contrib = mul_rn(fill, (T)nf);
            else contrib = (T)(fill * (T)nf);
        }
        return red_apply<T>(RF_ADD, v, contrib);
    }
    if (op == RF_MUL) {
        T contrib = T(1);
        if (nf != 0) {
            if constexpr (std::is_floating_point<T>::value) contrib = (T)pow((double)fill, (double)nf);
            else {
                T b = fill, r = T(1);
                int64_t e = nf;
                while (e > 0) {
                    if (e & 1) r = (T)(r * b);
                    b = (T)(b * b);
                    e >>= 1;
                }
                contrib = r;
            }
        }
        return red_apply<T>(RF_MUL, v, contrib);
    }
    return nf != 0 ? red_apply<T>(op, v, fill) : v;
}

template <typename T>
struct Run {  // state of the run that is open at some position
    int flag;     // 1 if a head lies at or before this position (within the scanned range)
    T val;
    int64_t cnt;
};

template <typename T>
__device__ __forceinline__ Run<T> run_combine(int op, const Run<T> &a, const Run<T> &b) {
    // b comes after a
    Run<T> r;
    r.flag = a.flag | b.flag;
    if (b.flag) {
        r.val = b.val;
        r.cnt = b.cnt;
    } else {
        r.val = (a.cnt == 0) ? b.val : ((b.cnt == 0) ? a.val : red_apply<T>(op, a.val, b.val));
        r.cnt = a.cnt + b.cnt;
    }
    return r;
}

template <typename T>
__device__ __forceinline__ Run<T> run_shfl_up(const Run<T> &x, int o) {
    Run<T> r;
    r.flag = __shfl_up_sync(0xffffffffu, x.flag, o);
    r.val = __shfl_up_sync(0xffffffffu, x.val, o);
    r.cnt = __shfl_up_sync(0xffffffffu, x.cnt, o);
    return r;
}

struct RdShape {
    int ndim;
    FastDiv extent[RD_MAXDIM];
};

// One tile.  EMIT = false: write the tile summary.  EMIT = true: write finished runs.
// Loads and stores go through shared memory so that global accesses are fully coalesced: the tile's 2048 (key, value)
// pairs are loaded with unit stride, each thread then reads its 8 consecutive pairs from a padded layout
// (index e + e/8: 2-way instead of 16-way bank conflicts), and in the emit pass the finished runs of the tile -- which
// occupy one contiguous range of output slots -- are staged in the same buffers and written out with unit stride.
constexpr int RD_PAD = RD_TILE + RD_TILE / 8 + 8;
__device__ __forceinline__ int rd_pad(int e) { return e + (e >> 3); }

template <typename T, bool EMIT>
__global__ void __launch_bounds__(RD_THREADS)
reduce_tile_kernel(const int64_t *__restrict__ keys, const T *__restrict__ vals, int64_t n, int64_t ncols,
                   FastDiv fcols, int op,
                   // summaries (pass 1 out / pass 2 in)
                   int64_t *__restrict__ t_heads, int *__restrict__ t_flag, T *__restrict__ t_val,
                   int64_t *__restrict__ t_cnt, const int64_t *__restrict__ head_off, const int *__restrict__ c_flag,
                   const T *__restrict__ c_val, const int64_t *__restrict__ c_cnt,
                   // outputs (pass 2)
                   T fill, int apply_fix, T result_fill, RdShape shp, int64_t *__restrict__ out_gid,
                   int64_t *__restrict__ out_coords, int64_t coords_stride, T *__restrict__ out_val,
                   unsigned long long *__restrict__ n_equal_fill) {
    __shared__ int64_t sk[RD_PAD];
    __shared__ T sv[RD_PAD];
    __shared__ int s_flag[RD_THREADS / 32];
    __shared__ T s_val[RD_THREADS / 32];
    __shared__ int64_t s_cnt[RD_THREADS / 32];
    __shared__ int s_heads[RD_THREADS / 32];
    const int64_t tile = blockIdx.x;
    const int64_t tile_base = tile * RD_TILE;
    const int64_t base = tile_base + (int64_t)threadIdx.x * RD_ITEMS;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;

    // coalesced tile load: slot e holds element tile_base + e (group id, value); slot -1 / RD_TILE = neighbours
#pragma unroll
    for (int i = 0; i < RD_ITEMS; ++i) {
        const int e = threadIdx.x + i * RD_THREADS;
        const int64_t p = tile_base + e;
        int64_t gk = -1;
        T vv = T(0);
        if (p < n) {
            gk = (int64_t)fcols.div((uint64_t)keys[p]);
            vv = vals[p];
        }
        sk[rd_pad(e)] = gk;
        sv[rd_pad(e)] = vv;
    }
    __shared__ int64_t s_edge[2];
    if (threadIdx.x == 0) {
        s_edge[0] = (tile_base > 0) ? (int64_t)fcols.div((uint64_t)keys[tile_base - 1]) : -1;
        s_edge[1] = (tile_base + RD_TILE < n) ? (int64_t)fcols.div((uint64_t)keys[tile_base + RD_TILE]) : -1;
    }
    __syncthreads();

    int64_t g[RD_ITEMS];
    T v[RD_ITEMS];
    bool head[RD_ITEMS];
    const int e0 = threadIdx.x * RD_ITEMS;
    const int64_t gprev = (threadIdx.x == 0) ? s_edge[0] : sk[rd_pad(e0 - 1)];
    const int64_t gnext_thread = (threadIdx.x == RD_THREADS - 1) ? s_edge[1] : sk[rd_pad(e0 + RD_ITEMS)];
#pragma unroll
    for (int i = 0; i < RD_ITEMS; ++i) {
        const int64_t p = base + i;
        g[i] = sk[rd_pad(e0 + i)];
        v[i] = sv[rd_pad(e0 + i)];
        head[i] = (p < n) && ((p == 0) || (g[i] != (i == 0 ? gprev : g[i - 1])));
    }
    // thread summary: run still open at the end of the thread's range
    Run<T> mine;
    mine.flag = 0;
    mine.val = T(0);
    mine.cnt = 0;
    int nheads = 0;
#pragma unroll
    for (int i = 0; i < RD_ITEMS; ++i) {
        if (base + i < n) {
            if (head[i]) {
                mine.flag = 1;
                mine.val = v[i];
                mine.cnt = 1;
                ++nheads;
            } else {
                mine.val = mine.cnt == 0 ? v[i] : red_apply<T>(op, mine.val, v[i]);
                mine.cnt += 1;
            }
        }
    }
    // block-level inclusive segmented scan of the thread summaries (+ plain scan of head counts)
    Run<T> incl = mine;
    int hincl = nheads;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const Run<T> up = run_shfl_up(incl, o);
        const int hu = __shfl_up_sync(0xffffffffu, hincl, o);
        if (lane >= o) {
            incl = run_combine<T>(op, up, incl);
            hincl += hu;
        }
    }
    if (lane == 31) {
        s_flag[w] = incl.flag;
        s_val[w] = incl.val;
        s_cnt[w] = incl.cnt;
        s_heads[w] = hincl;
    }
    __syncthreads();  // (also: every thread has copied its inputs out of sk / sv)
    Run<T> carry;
    carry.flag = 0;
    carry.val = T(0);
    carry.cnt = 0;
    int64_t hbefore = 0;
    if constexpr (EMIT) {
        carry.flag = c_flag[tile];
        carry.val = c_val[tile];
        carry.cnt = c_cnt[tile];
        hbefore = head_off[tile];
    }
    Run<T> tile_total = carry;
    int tile_heads = 0;
#pragma unroll
    for (int q = 0; q < RD_THREADS / 32; ++q) {
        Run<T> wq;
        wq.flag = s_flag[q];
        wq.val = s_val[q];
        wq.cnt = s_cnt[q];
        if (q < w) {
            carry = run_combine<T>(op, carry, wq);
            hbefore += s_heads[q];
        }
        tile_total = run_combine<T>(op, tile_total, wq);
        tile_heads += s_heads[q];
    }
    if constexpr (!EMIT) {
        if (threadIdx.x == 0) {
            t_heads[tile] = tile_heads;
            t_flag[tile] = tile_total.flag;
            t_val[tile] = tile_total.val;
            t_cnt[tile] = tile_total.cnt;
        }
        return;
    } else {
        const int64_t slot0 = head_off[tile] - 1;  // output slot of a run that started before this tile
        // the staging buffers alias the input buffers: mark every slot empty first
        for (int i = threadIdx.x; i < RD_TILE + 1; i += RD_THREADS) sk[i] = -1;
        __syncthreads();
        Run<T> prev = run_shfl_up(incl, 1);
        int hprev = __shfl_up_sync(0xffffffffu, hincl, 1);
        Run<T> st = carry;
        int64_t hcount = hbefore;  // heads strictly before this thread's first element
        if (lane > 0) {
            st = run_combine<T>(op, carry, prev);
            hcount += hprev;
        }
        T rv = st.val;
        int64_t rc = st.cnt;
#pragma unroll
        for (int i = 0; i < RD_ITEMS; ++i) {
            const int64_t p = base + i;
            if (p < n) {
                if (head[i]) {
                    rv = v[i];
                    rc = 1;
                    ++hcount;
                } else {
                    rv = rc == 0 ? v[i] : red_apply<T>(op, rv, v[i]);
                    rc += 1;
                }
                const bool last = (p == n - 1) ||
                                  ((i + 1 < RD_ITEMS) ? (base + i + 1 < n && head[i + 1]) : (g[i] != gnext_thread));
                if (last) {
                    const int loc = (int)(hcount - 1 - slot0);  // 0 .. RD_TILE
                    sk[loc] = g[i];
                    sv[loc] = apply_fix ? fill_fix<T>(op, rv, rc, ncols, fill) : rv;
                }
            }
        }
        __syncthreads();
        int eq = 0;
        for (int l = threadIdx.x; l < RD_TILE + 1; l += RD_THREADS) {
            const int64_t gid = sk[l];
            if (gid >= 0) {
                const int64_t idx = slot0 + l;
                const T outv = sv[l];
                out_val[idx] = outv;
                out_gid[idx] = gid;
                if (out_coords) {
                    uint64_t k = (uint64_t)gid;
                    for (int d = shp.ndim - 1; d >= 0; --d) {
                        uint64_t q, r;
                        shp.extent[d].divmod(k, q, r);
                        out_coords[(int64_t)d * coords_stride + idx] = (int64_t)r;
                        k = q;
                    }
                }
                bool same;
                if constexpr (sizeof(T) == 1) same = (*(const uint8_t *)&outv) == (*(const uint8_t *)&result_fill);
                else if constexpr (sizeof(T) == 4) {
                    uint32_t x, y;
                    memcpy(&x, &outv, 4);
                    memcpy(&y, &result_fill, 4);
                    same = x == y;
                } else {
                    uint64_t x, y;
                    memcpy(&x, &outv, 8);
                    memcpy(&y, &result_fill, 8);
                    same = x == y;
                }
                eq += same ? 1 : 0;
            }
        }
        eq = __reduce_add_sync(0xffffffffu, eq);
        if (lane == 0 && eq) atomicAdd(n_equal_fill, (unsigned long long)eq);
    }
}

// single-CTA segmented scan over the tile summaries -> per-tile head offsets and carry-in runs
template <typename T>
__global__ void __launch_bounds__(1024)
reduce_scan_tiles_kernel(int64_t ntiles, int op, const int64_t *__restrict__ t_heads, const int *__restrict__ t_flag,
                         const T *__restrict__ t_val, const int64_t *__restrict__ t_cnt,
                         int64_t *__restrict__ head_off, int *__restrict__ c_flag, T *__restrict__ c_val,
                         int64_t *__restrict__ c_cnt) {
    __shared__ int s_flag[32];
    __shared__ T s_val[32];
    __shared__ int64_t s_cnt[32];
    __shared__ int64_t s_heads[32];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int64_t per = (ntiles + 1023) / 1024;
    const int64_t lo = (int64_t)tid * per, hi = (lo + per < ntiles) ? lo + per : ntiles;
    Run<T> mine;
    mine.flag = 0;
    mine.val = T(0);
    mine.cnt = 0;
    int64_t hsum = 0;
    for (int64_t t = lo; t < hi; ++t) {
        Run<T> x;
        x.flag = t_flag[t];
        x.val = t_val[t];
        x.cnt = t_cnt[t];
        mine = run_combine<T>(op, mine, x);
        hsum += t_heads[t];
    }
    Run<T> incl = mine;
    int64_t hincl = hsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const Run<T> up = run_shfl_up(incl, o);
        const int64_t hu = __shfl_up_sync(0xffffffffu, hincl, o);
        if (lane >= o) {
            incl = run_combine<T>(op, up, incl);
            hincl += hu;
        }
    }
    if (lane == 31) {
        s_flag[w] = incl.flag;
        s_val[w] = incl.val;
        s_cnt[w] = incl.cnt;
        s_heads[w] = hincl;
    }
    __syncthreads();
    Run<T> carry;
    carry.flag = 0;
    carry.val = T(0);
    carry.cnt = 0;
    int64_t hbefore = 0, htotal = 0;
    for (int q = 0; q < 32; ++q) {
        Run<T> wq;
        wq.flag = s_flag[q];
        wq.val = s_val[q];
        wq.cnt = s_cnt[q];
        if (q < w) {
            carry = run_combine<T>(op, carry, wq);
            hbefore += s_heads[q];
        }
        htotal += s_heads[q];
    }
    Run<T> prev = run_shfl_up(incl, 1);
    int64_t hprev = __shfl_up_sync(0xffffffffu, hincl, 1);
    Run<T> st = carry;
    int64_t hc = hbefore;
    if (lane > 0) {
        st = run_combine<T>(op, carry, prev);
        hc += hprev;
    }
    for (int64_t t = lo; t < hi; ++t) {
        head_off[t] = hc;
        c_flag[t] = st.flag;
        c_val[t] = st.val;
        c_cnt[t] = st.cnt;
        Run<T> x;
        x.flag = t_flag[t];
        x.val = t_val[t];
        x.cnt = t_cnt[t];
        st = run_combine<T>(op, st, x);
        hc += t_heads[t];
    }
    if (tid == 0) head_off[ntiles] = htotal;
}

struct RdPlan {
    int dtype, op;
    int64_t n, ncols, ntiles, ngroups;
    const int64_t *keys;
    const void *vals;
    int64_t *t_heads, *t_cnt, *head_off, *c_cnt;
    int *t_flag, *c_flag;
    void *t_val, *c_val;
    unsigned long long *n_eq;
    cudaStream_t stream;
};

static void rd_free(RdPlan *pl) {
    cudaStream_t s = pl->stream;
    scratch_free(pl->t_heads, s);
    scratch_free(pl->t_cnt, s);
    scratch_free(pl->head_off, s);
    scratch_free(pl->c_cnt, s);
    scratch_free(pl->t_flag, s);
    scratch_free(pl->c_flag, s);
    scratch_free(pl->t_val, s);
    scratch_free(pl->c_val, s);
    scratch_free(pl->n_eq, s);
    delete pl;
}

template <typename T>
static int rd_begin_t(RdPlan *pl) {
    cudaStream_t s = pl->stream;
    const int64_t nt = pl->ntiles;
    RdShape shp{};
    reduce_tile_kernel<T, false><<<(unsigned)nt, RD_THREADS, 0, s>>>(
        pl->keys, (const T *)pl->vals, pl->n, pl->ncols, make_fastdiv((uint64_t)pl->ncols), pl->op, pl->t_heads, pl->t_flag, (T *)pl->t_val, pl->t_cnt,
        nullptr, nullptr, nullptr, nullptr, T(0), 0, T(0), shp, nullptr, nullptr, 0, nullptr, nullptr);
    B2S_CHECK_LAUNCH();
    reduce_scan_tiles_kernel<T><<<1, 1024, 0, s>>>(nt, pl->op, pl->t_heads, pl->t_flag, (const T *)pl->t_val, pl->t_cnt,
                                                   pl->head_off, pl->c_flag, (T *)pl->c_val, pl->c_cnt);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

template <typename T>
static int rd_finish_t(RdPlan *pl, const void *fill_host, int apply_fix, const void *result_fill_host, int ndim,
                       const int64_t *shape_host, int64_t *gid_out, int64_t *coords_out, int64_t coords_stride,
                       void *vals_out) {
    cudaStream_t s = pl->stream;
    RdShape shp{};
    shp.ndim = ndim;
    for (int d = 0; d < ndim; ++d) shp.extent[d] = make_fastdiv((uint64_t)shape_host[d]);
    T fill, rfill;
    memcpy(&fill, fill_host, sizeof(T));
    memcpy(&rfill, result_fill_host, sizeof(T));
    reduce_tile_kernel<T, true><<<(unsigned)pl->ntiles, RD_THREADS, 0, s>>>(
        pl->keys, (const T *)pl->vals, pl->n, pl->ncols, make_fastdiv((uint64_t)pl->ncols), pl->op, nullptr, nullptr,
        nullptr, nullptr, pl->head_off,
        pl->c_flag, (const T *)pl->c_val, pl->c_cnt, fill, apply_fix, rfill, shp, gid_out, coords_out, coords_stride,
        (T *)vals_out, pl->n_eq);
    B2S_CHECK_LAUNCH();
    return B2S_OK;
}

}  // namespace b2s

using namespace b2s;

extern "C" {

/*
 * Segmented reduction of `vals` over runs of equal group id (= key / ncols) of the sorted `keys`.
 * begin(): pass 1 + tile scan, returns the number of groups (one stream sync).
 * finish(): pass 2 writes group ids, values (fill-value contribution applied when apply_fill_fix) and, if
 * coords_out is given, the coordinates of every group id unravelled over shape_host[ndim]; returns the number of
 * results bitwise equal to result_fill (one stream sync) so the caller can skip the prune compaction when it is 0.
 */
int b2s_reduce_begin(int dtype, int op, const int64_t *keys_dev, const void *vals_dev, int64_t n, int64_t ncols,
                     void **plan_out, int64_t *n_groups_out, void *stream) {
    B2S_REQUIRE(plan_out && n_groups_out, B2S_ERR_INVALID, "reduce_begin: NULL output");
    B2S_REQUIRE(n >= 0 && ncols >= 1, B2S_ERR_INVALID, "reduce_begin: bad sizes");
    cudaStream_t s = (cudaStream_t)stream;
    RdPlan *pl = new RdPlan();
    memset(pl, 0, sizeof(*pl));
    pl->dtype = dtype;
    pl->op = op;
    pl->n = n;
    pl->ncols = ncols;
    pl->keys = keys_dev;
    pl->vals = vals_dev;
    pl->stream = s;
    *plan_out = pl;
    *n_groups_out = 0;
    if (n == 0) return B2S_OK;
    const int64_t nt = (n + RD_TILE - 1) / RD_TILE;
    B2S_REQUIRE(nt < 2147483647LL, B2S_ERR_OVERFLOW, "reduce: too many tiles");
    pl->ntiles = nt;
    int rc;
    const size_t es = dtype_size(dtype);
    if ((rc = scratch_alloc((void **)&pl->t_heads, (size_t)nt * 8, s)) ||
        (rc = scratch_alloc((void **)&pl->t_cnt, (size_t)nt * 8, s)) ||
        (rc = scratch_alloc((void **)&pl->head_off, (size_t)(nt + 1) * 8, s)) ||
        (rc = scratch_alloc((void **)&pl->c_cnt, (size_t)nt * 8, s)) ||
        (rc = scratch_alloc((void **)&pl->t_flag, (size_t)nt * 4, s)) ||
        (rc = scratch_alloc((void **)&pl->c_flag, (size_t)nt * 4, s)) ||
        (rc = scratch_alloc((void **)&pl->t_val, (size_t)nt * es, s)) ||
        (rc = scratch_alloc((void **)&pl->c_val, (size_t)nt * es, s)) ||
        (rc = scratch_alloc((void **)&pl->n_eq, 8, s)))
        return rc;
    B2S_CUDA(cudaMemsetAsync(pl->n_eq, 0, 8, s));
    switch (dtype) {
        case B2S_F32: rc = rd_begin_t<float>(pl); break;
        case B2S_F64: rc = rd_begin_t<double>(pl); break;
        case B2S_I32: rc = rd_begin_t<int32_t>(pl); break;
        case B2S_I64: rc = rd_begin_t<int64_t>(pl); break;
        case B2S_BOOL: rc = rd_begin_t<uint8_t>(pl); break;
        default: set_error("reduce_begin: dtype %d", dtype); rc = B2S_ERR_UNSUPPORTED;
    }
    if (rc) return rc;
    B2S_CUDA(cudaMemcpyAsync(&pl->ngroups, pl->head_off + nt, 8, cudaMemcpyDeviceToHost, s));
    B2S_CUDA(cudaStreamSynchronize(s));
    *n_groups_out = pl->ngroups;
    return B2S_OK;
}

int b2s_reduce_finish(void *plan, const void *fill_host, int apply_fill_fix, const void *result_fill_host, int ndim,
                      const int64_t *shape_host, int64_t *gid_out_dev, int64_t *coords_out_or_null_dev,
                      int64_t coords_stride, void *vals_out_dev, int64_t *n_equal_fill_host) {
    B2S_REQUIRE(plan != nullptr, B2S_ERR_INVALID, "reduce_finish: NULL plan");
    RdPlan *pl = (RdPlan *)plan;
    int rc = B2S_OK;
    if (n_equal_fill_host) *n_equal_fill_host = 0;
    if (ndim < 0 || ndim > RD_MAXDIM) {
        set_error("reduce_finish: ndim %d", ndim);
        rc = B2S_ERR_UNSUPPORTED;
    } else if (pl->n > 0 && pl->ngroups > 0) {
        switch (pl->dtype) {
            case B2S_F32: rc = rd_finish_t<float>(pl, fill_host, apply_fill_fix, result_fill_host, ndim, shape_host, gid_out_dev, coords_out_or_null_dev, coords_stride, vals_out_dev); break;
            case B2S_F64: rc = rd_finish_t<double>(pl, fill_host, apply_fill_fix, result_fill_host, ndim, shape_host, gid_out_dev, coords_out_or_null_dev, coords_stride, vals_out_dev); break;
            case B2S_I32: rc = rd_finish_t<int32_t>(pl, fill_host, apply_fill_fix, result_fill_host, ndim, shape_host, gid_out_dev, coords_out_or_null_dev, coords_stride, vals_out_dev); break;
            case B2S_I64: rc = rd_finish_t<int64_t>(pl, fill_host, apply_fill_fix, result_fill_host, ndim, shape_host, gid_out_dev, coords_out_or_null_dev, coords_stride, vals_out_dev); break;
            case B2S_BOOL: rc = rd_finish_t<uint8_t>(pl, fill_host, apply_fill_fix, result_fill_host, ndim, shape_host, gid_out_dev, coords_out_or_null_dev, coords_stride, vals_out_dev); break;
            default: rc = B2S_ERR_UNSUPPORTED;
        }
        if (rc == B2S_OK && n_equal_fill_host) {
            unsigned long long h = 0;
            cudaError_t e = cudaMemcpyAsync(&h, pl->n_eq, 8, cudaMemcpyDeviceToHost, pl->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(pl->stream);
            if (e != cudaSuccess) {
                set_error("reduce_finish: %s", cudaGetErrorString(e));
                rc = B2S_ERR_CUDA;
            }
            *n_equal_fill_host = (int64_t)h;
        }
    }
    rd_free(pl);
    return rc;
}

}  // extern "C"
