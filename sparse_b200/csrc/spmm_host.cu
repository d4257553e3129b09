// spmm_host.cu -- the reference-facing HOST-buffer form of K1: numpy arrays in, numpy array out.
//
// A three-stream pipeline keeps both PCIe directions and the SMs busy: B goes up first, then A streams up in
// nnz-balanced row chunks; as soon as a chunk has landed its indices are narrowed to int32 on the device and K1 runs
// on that row range (bit-identical to the one-shot kernel: rows are independent), and the finished rows of C stream
// back on the third stream.  End-to-end time ~= the H2D time of the operands (the larger PCIe direction) instead of
// H2D + kernel + D2H.  64-bit column indices are narrowed to 32 bits ON THE HOST (thread pool -> pinned staging ring)
// so that only half of their bytes cross the bus.  Host buffers should be pinned (b2s_host_register or pinned allocations); pageable memory still
// works but serialises.
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "common.cuh"
#include "host_pool.h"

using namespace b2s;

namespace {
// Host-side index narrowing: int64 column indices are converted to int32 by a pool of host threads into a ring of
// pinned staging slots while the previous chunks are on the bus, so only 4 of their 8 bytes cross PCIe (at C2: 1.31 GB
// up instead of 1.71 GB).  g_host_threads: -1 = auto, 0 = off (raw int64 upload + device narrowing), n = n threads.
int g_host_threads = -1;
int g_chunks = 16;            // nnz-balanced row chunks of the pipeline
int g_slots = 4;              // pinned staging slots in use (<= kSlots)
constexpr int kSlots = 32;
struct Staging {
    void *buf[kSlots] = {};
    cudaEvent_t ev[kSlots] = {};
    bool busy[kSlots] = {};
    size_t bytes = 0;
    int slots = 0;
};
Staging &staging() {
    static Staging s;
    return s;
}
int staging_reserve(size_t bytes) {
    Staging &S = staging();
    if (S.bytes >= bytes && S.slots >= g_slots) return B2S_OK;
    for (int i = 0; i < kSlots; ++i) {
        if (S.buf[i]) cudaFreeHost(S.buf[i]);
        S.buf[i] = nullptr;
    }
    S.bytes = 0;
    S.slots = 0;
    const size_t want = bytes + bytes / 8;
    for (int i = 0; i < g_slots; ++i) {
        B2S_CUDA(cudaHostAlloc(&S.buf[i], want, cudaHostAllocDefault));
        if (!S.ev[i]) B2S_CUDA(cudaEventCreateWithFlags(&S.ev[i], cudaEventDisableTiming));
        S.busy[i] = false;
    }
    S.bytes = want;
    S.slots = g_slots;
    return B2S_OK;
}
// dst[i] = (int32) src[i] over [0, n): AVX2 (low dwords gathered with one cross-lane permute per 4 elements,
// non-temporal stores so the staging slot is not read before it is written) with a scalar fallback
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
__attribute__((target("avx2"))) void narrow_range_avx2(const int64_t *src, int32_t *dst, int64_t n) {
    int64_t i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 31)) {
        dst[i] = (int32_t)src[i];
        ++i;
    }
    const __m256i idx = _mm256_setr_epi32(0, 2, 4, 6, 0, 2, 4, 6);
    for (; i + 8 <= n; i += 8) {
        const __m256i a = _mm256_loadu_si256((const __m256i *)(src + i));
        const __m256i b = _mm256_loadu_si256((const __m256i *)(src + i + 4));
        const __m256i lo = _mm256_permutevar8x32_epi32(a, idx);  // a's low dwords in both halves
        const __m256i hi = _mm256_permutevar8x32_epi32(b, idx);
        _mm256_stream_si256((__m256i *)(dst + i), _mm256_blend_epi32(lo, hi, 0xF0));
    }
    for (; i < n; ++i) dst[i] = (int32_t)src[i];
    _mm_sfence();
}
bool have_avx2() {
    static const bool v = __builtin_cpu_supports("avx2");
    return v;
}
#else
bool have_avx2() { return false; }
void narrow_range_avx2(const int64_t *, int32_t *, int64_t) {}
#endif
void narrow_range(const int64_t *src, int32_t *dst, int64_t n) {
    if (have_avx2()) {
        narrow_range_avx2(src, dst, n);
        return;
    }
    for (int64_t i = 0; i < n; ++i) dst[i] = (int32_t)src[i];
}

std::mutex &host_pipe_mutex() {
    static std::mutex m;
    return m;
}

HostPool &host_pool() {
    static HostPool *p = nullptr;  // leaked on purpose, see host_pool.h
    static int built_for = -2;
    int want = g_host_threads;
    if (want < 0) {
        const unsigned hc = std::thread::hardware_concurrency();
        want = hc >= 16 ? 8 : (hc >= 4 ? (int)hc / 2 : 1);  // 8 threads narrow 1e8 indices in ~9 ms (measured)
    }
    if (!p || built_for != want) {
        p = new HostPool(want - 1);
        built_for = want;
    }
    return *p;
}

struct HostPipe {
    cudaStream_t in = nullptr, cmp = nullptr, out = nullptr;
    bool ok = false;
};
HostPipe &pipe() {
    static HostPipe p;
    if (!p.ok) {
        if (cudaStreamCreateWithFlags(&p.in, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&p.cmp, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&p.out, cudaStreamNonBlocking) == cudaSuccess)
            p.ok = true;
    }
    return p;
}
}  // namespace

extern "C" int b2s_spmm_host_set_threads(int n) {
    g_host_threads = n;
    return B2S_OK;
}

extern "C" int b2s_spmm_host_set_pipeline(int chunks, int slots) {
    B2S_REQUIRE(chunks >= 1 && chunks <= 1024 && slots >= 1 && slots <= kSlots, B2S_ERR_INVALID,
                "spmm_host_set_pipeline: chunks in 1..1024, slots in 1..%d", kSlots);
    g_chunks = chunks;
    g_slots = slots;
    return B2S_OK;
}

extern "C" int b2s_host_narrow_i64_i32(const int64_t *src_host, int32_t *dst_host, int64_t n) {
    std::lock_guard<std::mutex> guard(host_pipe_mutex());
    host_pool().parallel_for(n, [src_host, dst_host](int64_t b, int64_t e) {
        narrow_range(src_host + b, dst_host + b, e - b);
    });
    return B2S_OK;
}

extern "C" int b2s_spmm_csr_dense_host(int dtype, int idx_bytes, int64_t M, int64_t K, int64_t N, int64_t nnz,
                                       const void *a_data_host, const void *a_indices_host,
                                       const void *a_indptr_host, const void *b_host, void *out_host) {
    const size_t es = dtype_size(dtype);
    B2S_REQUIRE(es != 0 && dtype != B2S_BOOL, B2S_ERR_UNSUPPORTED, "spmm_host: unsupported dtype %d", dtype);
    B2S_REQUIRE(idx_bytes == 4 || idx_bytes == 8, B2S_ERR_INVALID, "spmm_host: idx_bytes must be 4 or 8");
    B2S_REQUIRE(M >= 0 && K >= 0 && N >= 0 && nnz >= 0, B2S_ERR_INVALID, "spmm_host: negative size");
    if (M == 0 || N == 0) return B2S_OK;
    // the three pipeline streams, the pinned staging ring and the host thread pool are process-wide: callers that come
    // in from several threads (ctypes releases the GIL) take turns -- they would share the PCIe link anyway
    std::lock_guard<std::mutex> guard(host_pipe_mutex());
    HostPipe &P = pipe();
    B2S_REQUIRE(P.ok, B2S_ERR_CUDA, "spmm_host: could not create CUDA streams");
    const bool narrow = idx_bytes == 8 && (K < 2147483647LL) && (nnz < 2147483647LL);
    const int dev_ib = (idx_bytes == 4 || narrow) ? 4 : 8;

    auto ptr_at = [&](int64_t r) -> int64_t {
        return idx_bytes == 8 ? ((const int64_t *)a_indptr_host)[r] : (int64_t)((const int32_t *)a_indptr_host)[r];
    };
    // nnz-balanced row chunks (~16, at least 4096 rows or the whole matrix)
    int nchunks = g_chunks;
    if (nnz < (1 << 20) || M < 8192) nchunks = 1;
    std::vector<int64_t> cut(nchunks + 1, 0);
    cut[nchunks] = M;
    for (int c = 1; c < nchunks; ++c) {
        const int64_t target = nnz / nchunks * c;
        int64_t lo = cut[c - 1], hi = M;  // first row whose start offset >= target
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (ptr_at(mid) < target) lo = mid + 1;
            else hi = mid;
        }
        cut[c] = lo;
    }

    // narrow the indices on the host (pinned staging ring) unless switched off or there is nothing to narrow
    const bool host_narrow = narrow && g_host_threads != 0 && nchunks > 1;
    size_t max_chunk = 0;
    for (int c = 0; c < nchunks; ++c) {
        const size_t len = (size_t)(ptr_at(cut[c + 1]) - ptr_at(cut[c]));
        if (len > max_chunk) max_chunk = len;
    }

    void *d_ad = nullptr, *d_ai_raw = nullptr, *d_ap_raw = nullptr, *d_ai = nullptr, *d_ap = nullptr, *d_b = nullptr,
         *d_out = nullptr;
    std::vector<cudaEvent_t> ev_in(nchunks, nullptr), ev_cmp(nchunks, nullptr), ev_out(nchunks, nullptr);
    cudaEvent_t ev_b = nullptr, ev_t0 = nullptr;
    const bool trace = getenv("B2S_HOST_TRACE") != nullptr;  // print the pipeline timeline (debugging aid)
    int rc = B2S_OK;
#define B2S_TRY(x)                   \
    do {                             \
        rc = (x);                    \
        if (rc != B2S_OK) goto done; \
    } while (0)
#define B2S_TRYCUDA(x)                                           \
    do {                                                         \
        cudaError_t _e = (x);                                    \
        if (_e != cudaSuccess) {                                 \
            set_error("%s: %s", #x, cudaGetErrorString(_e));     \
            rc = B2S_ERR_CUDA;                                   \
            goto done;                                           \
        }                                                        \
    } while (0)
    B2S_TRY(scratch_alloc(&d_b, (size_t)K * N * es, P.in));
    B2S_TRY(scratch_alloc(&d_ap_raw, (size_t)(M + 1) * idx_bytes, P.in));
    B2S_TRY(scratch_alloc(&d_ad, (size_t)nnz * es, P.in));
    if (!host_narrow) B2S_TRY(scratch_alloc(&d_ai_raw, (size_t)nnz * idx_bytes, P.in));
    else B2S_TRY(staging_reserve(max_chunk * 4));
    B2S_TRY(scratch_alloc(&d_out, (size_t)M * N * es, P.in));
    if (narrow) {
        B2S_TRY(scratch_alloc(&d_ai, (size_t)nnz * 4, P.in));
        B2S_TRY(scratch_alloc(&d_ap, (size_t)(M + 1) * 4, P.in));
    } else {
        d_ai = d_ai_raw;
        d_ap = d_ap_raw;
    }
    B2S_TRYCUDA(cudaEventCreateWithFlags(&ev_b, trace ? cudaEventDefault : cudaEventDisableTiming));
    for (int c = 0; c < nchunks; ++c) {
        B2S_TRYCUDA(cudaEventCreateWithFlags(&ev_in[c], trace ? cudaEventDefault : cudaEventDisableTiming));
        B2S_TRYCUDA(cudaEventCreateWithFlags(&ev_cmp[c], trace ? cudaEventDefault : cudaEventDisableTiming));
        if (trace) B2S_TRYCUDA(cudaEventCreate(&ev_out[c]));
    }
    if (trace) {
        B2S_TRYCUDA(cudaEventCreate(&ev_t0));
        B2S_TRYCUDA(cudaEventRecord(ev_t0, P.in));
    }
    // B and indptr first: every chunk needs them
    B2S_TRYCUDA(cudaMemcpyAsync(d_ap_raw, a_indptr_host, (size_t)(M + 1) * idx_bytes, cudaMemcpyHostToDevice, P.in));
    B2S_TRYCUDA(cudaMemcpyAsync(d_b, b_host, (size_t)K * N * es, cudaMemcpyHostToDevice, P.in));
    B2S_TRYCUDA(cudaEventRecord(ev_b, P.in));
    B2S_TRYCUDA(cudaStreamWaitEvent(P.cmp, ev_b, 0));
    if (narrow) B2S_TRY(narrow_i64_i32((const int64_t *)d_ap_raw, (int32_t *)d_ap, M + 1, P.cmp));
    for (int c = 0; c < nchunks; ++c) {
        const int64_t r0 = cut[c], r1 = cut[c + 1];
        if (r1 <= r0) continue;
        const int64_t lo = ptr_at(r0), hi = ptr_at(r1);
        if (hi > lo && host_narrow) {
            Staging &S = staging();
            const int slot = c % g_slots;
            if (S.busy[slot]) B2S_TRYCUDA(cudaEventSynchronize(S.ev[slot]));  // its previous upload has left the slot
            const int64_t *src = (const int64_t *)a_indices_host + lo;
            int32_t *dst = (int32_t *)S.buf[slot];
            host_pool().parallel_for(hi - lo, [src, dst](int64_t b, int64_t e) {
                narrow_range(src + b, dst + b, e - b);
            });
            B2S_TRYCUDA(cudaMemcpyAsync((char *)d_ad + (size_t)lo * es, (const char *)a_data_host + (size_t)lo * es,
                                        (size_t)(hi - lo) * es, cudaMemcpyHostToDevice, P.in));
            B2S_TRYCUDA(cudaMemcpyAsync((int32_t *)d_ai + lo, dst, (size_t)(hi - lo) * 4, cudaMemcpyHostToDevice, P.in));
            B2S_TRYCUDA(cudaEventRecord(S.ev[slot], P.in));
            S.busy[slot] = true;
        } else if (hi > lo) {
            B2S_TRYCUDA(cudaMemcpyAsync((char *)d_ad + (size_t)lo * es, (const char *)a_data_host + (size_t)lo * es,
                                        (size_t)(hi - lo) * es, cudaMemcpyHostToDevice, P.in));
            B2S_TRYCUDA(cudaMemcpyAsync((char *)d_ai_raw + (size_t)lo * idx_bytes,
                                        (const char *)a_indices_host + (size_t)lo * idx_bytes,
                                        (size_t)(hi - lo) * idx_bytes, cudaMemcpyHostToDevice, P.in));
        }
        B2S_TRYCUDA(cudaEventRecord(ev_in[c], P.in));
        B2S_TRYCUDA(cudaStreamWaitEvent(P.cmp, ev_in[c], 0));
        if (narrow && !host_narrow && hi > lo)
            B2S_TRY(narrow_i64_i32((const int64_t *)d_ai_raw + lo, (int32_t *)d_ai + lo, hi - lo, P.cmp));
        // rows [r0, r1): indptr holds absolute offsets, so the data/indices base pointers stay unshifted
        B2S_TRY(spmm_csr_dense_impl(dtype, dev_ib, r1 - r0, K, N, d_ad, d_ai, (const char *)d_ap + (size_t)r0 * dev_ib,
                                    d_b, N, (char *)d_out + (size_t)r0 * N * es, N, P.cmp));
        B2S_TRYCUDA(cudaEventRecord(ev_cmp[c], P.cmp));
        B2S_TRYCUDA(cudaStreamWaitEvent(P.out, ev_cmp[c], 0));
        B2S_TRYCUDA(cudaMemcpyAsync((char *)out_host + (size_t)r0 * N * es, (const char *)d_out + (size_t)r0 * N * es,
                                    (size_t)(r1 - r0) * N * es, cudaMemcpyDeviceToHost, P.out));
        if (trace) B2S_TRYCUDA(cudaEventRecord(ev_out[c], P.out));
    }
    B2S_TRYCUDA(cudaStreamSynchronize(P.out));
    B2S_TRYCUDA(cudaStreamSynchronize(P.cmp));
    B2S_TRYCUDA(cudaStreamSynchronize(P.in));
    if (trace) {
        float tb = 0.f;
        cudaEventElapsedTime(&tb, ev_t0, ev_b);
        fprintf(stderr, "[b2s host pipe] B landed %.2f ms\n", tb);
        for (int c = 0; c < nchunks; ++c) {
            float a = 0.f, k = 0.f, o = 0.f;
            if (cut[c + 1] <= cut[c]) continue;
            cudaEventElapsedTime(&a, ev_t0, ev_in[c]);
            cudaEventElapsedTime(&k, ev_t0, ev_cmp[c]);
            cudaEventElapsedTime(&o, ev_t0, ev_out[c]);
            fprintf(stderr, "[b2s host pipe] chunk %2d: in %.2f  kernel done %.2f  out done %.2f ms\n", c, a, k, o);
        }
    }
done:
    if (rc != B2S_OK) {
        cudaStreamSynchronize(P.in);
        cudaStreamSynchronize(P.cmp);
        cudaStreamSynchronize(P.out);
    }
    if (ev_b) cudaEventDestroy(ev_b);
    for (int c = 0; c < nchunks; ++c) {
        if (ev_in[c]) cudaEventDestroy(ev_in[c]);
        if (ev_cmp[c]) cudaEventDestroy(ev_cmp[c]);
        if (ev_out[c]) cudaEventDestroy(ev_out[c]);
    }
    if (ev_t0) cudaEventDestroy(ev_t0);
    scratch_free(d_ad, P.in);
    if (d_ai_raw) scratch_free(d_ai_raw, P.in);
    scratch_free(d_ap_raw, P.in);
    if (narrow) {
        scratch_free(d_ai, P.in);
        scratch_free(d_ap, P.in);
    }
    scratch_free(d_b, P.in);
    scratch_free(d_out, P.in);
#undef B2S_TRY
#undef B2S_TRYCUDA
    return rc;
}
