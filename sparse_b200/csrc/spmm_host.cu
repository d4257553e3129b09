// spmm_host.cu -- the reference-facing HOST-buffer form of K1: numpy arrays in, numpy array out.
//
// A three-stream pipeline keeps both PCIe directions and the SMs busy: B goes up first, then A streams up in
// nnz-balanced row chunks; as soon as a chunk has landed its indices are narrowed to int32 on the device and K1 runs
// on that row range (bit-identical to the one-shot kernel: rows are independent), and the finished rows of C stream
// back on the third stream.  End-to-end time ~= the H2D time of the operands (the larger PCIe direction) instead of
// H2D + kernel + D2H.  Host buffers should be pinned (b2s_host_register or pinned allocations); pageable memory still
// works but serialises.
#include <vector>

#include "common.cuh"

using namespace b2s;

namespace {
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

extern "C" int b2s_spmm_csr_dense_host(int dtype, int idx_bytes, int64_t M, int64_t K, int64_t N, int64_t nnz,
                                       const void *a_data_host, const void *a_indices_host,
                                       const void *a_indptr_host, const void *b_host, void *out_host) {
    const size_t es = dtype_size(dtype);
    B2S_REQUIRE(es != 0 && dtype != B2S_BOOL, B2S_ERR_UNSUPPORTED, "spmm_host: unsupported dtype %d", dtype);
    B2S_REQUIRE(idx_bytes == 4 || idx_bytes == 8, B2S_ERR_INVALID, "spmm_host: idx_bytes must be 4 or 8");
    B2S_REQUIRE(M >= 0 && K >= 0 && N >= 0 && nnz >= 0, B2S_ERR_INVALID, "spmm_host: negative size");
    if (M == 0 || N == 0) return B2S_OK;
    HostPipe &P = pipe();
    B2S_REQUIRE(P.ok, B2S_ERR_CUDA, "spmm_host: could not create CUDA streams");
    const bool narrow = idx_bytes == 8 && (K < 2147483647LL) && (nnz < 2147483647LL);
    const int dev_ib = (idx_bytes == 4 || narrow) ? 4 : 8;

    auto ptr_at = [&](int64_t r) -> int64_t {
        return idx_bytes == 8 ? ((const int64_t *)a_indptr_host)[r] : (int64_t)((const int32_t *)a_indptr_host)[r];
    };
    // nnz-balanced row chunks (~16, at least 4096 rows or the whole matrix)
    int nchunks = 16;
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

    void *d_ad = nullptr, *d_ai_raw = nullptr, *d_ap_raw = nullptr, *d_ai = nullptr, *d_ap = nullptr, *d_b = nullptr,
         *d_out = nullptr;
    std::vector<cudaEvent_t> ev_in(nchunks, nullptr), ev_cmp(nchunks, nullptr);
    cudaEvent_t ev_b = nullptr;
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
    B2S_TRY(scratch_alloc(&d_ai_raw, (size_t)nnz * idx_bytes, P.in));
    B2S_TRY(scratch_alloc(&d_out, (size_t)M * N * es, P.in));
    if (narrow) {
        B2S_TRY(scratch_alloc(&d_ai, (size_t)nnz * 4, P.in));
        B2S_TRY(scratch_alloc(&d_ap, (size_t)(M + 1) * 4, P.in));
    } else {
        d_ai = d_ai_raw;
        d_ap = d_ap_raw;
    }
    B2S_TRYCUDA(cudaEventCreateWithFlags(&ev_b, cudaEventDisableTiming));
    for (int c = 0; c < nchunks; ++c) {
        B2S_TRYCUDA(cudaEventCreateWithFlags(&ev_in[c], cudaEventDisableTiming));
        B2S_TRYCUDA(cudaEventCreateWithFlags(&ev_cmp[c], cudaEventDisableTiming));
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
        if (hi > lo) {
            B2S_TRYCUDA(cudaMemcpyAsync((char *)d_ad + (size_t)lo * es, (const char *)a_data_host + (size_t)lo * es,
                                        (size_t)(hi - lo) * es, cudaMemcpyHostToDevice, P.in));
            B2S_TRYCUDA(cudaMemcpyAsync((char *)d_ai_raw + (size_t)lo * idx_bytes,
                                        (const char *)a_indices_host + (size_t)lo * idx_bytes,
                                        (size_t)(hi - lo) * idx_bytes, cudaMemcpyHostToDevice, P.in));
        }
        B2S_TRYCUDA(cudaEventRecord(ev_in[c], P.in));
        B2S_TRYCUDA(cudaStreamWaitEvent(P.cmp, ev_in[c], 0));
        if (narrow && hi > lo)
            B2S_TRY(narrow_i64_i32((const int64_t *)d_ai_raw + lo, (int32_t *)d_ai + lo, hi - lo, P.cmp));
        // rows [r0, r1): indptr holds absolute offsets, so the data/indices base pointers stay unshifted
        B2S_TRY(spmm_csr_dense_impl(dtype, dev_ib, r1 - r0, K, N, d_ad, d_ai, (const char *)d_ap + (size_t)r0 * dev_ib,
                                    d_b, N, (char *)d_out + (size_t)r0 * N * es, N, P.cmp));
        B2S_TRYCUDA(cudaEventRecord(ev_cmp[c], P.cmp));
        B2S_TRYCUDA(cudaStreamWaitEvent(P.out, ev_cmp[c], 0));
        B2S_TRYCUDA(cudaMemcpyAsync((char *)out_host + (size_t)r0 * N * es, (const char *)d_out + (size_t)r0 * N * es,
                                    (size_t)(r1 - r0) * N * es, cudaMemcpyDeviceToHost, P.out));
    }
    B2S_TRYCUDA(cudaStreamSynchronize(P.out));
    B2S_TRYCUDA(cudaStreamSynchronize(P.cmp));
    B2S_TRYCUDA(cudaStreamSynchronize(P.in));
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
    }
    scratch_free(d_ad, P.in);
    scratch_free(d_ai_raw, P.in);
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
