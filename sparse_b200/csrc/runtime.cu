// runtime.cu -- error state, device queries, memory and stream wrappers of the C ABI.
#include <atomic>
#include <cstdarg>

#include "common.cuh"

namespace b2s {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            cached = kNumSMsB200;
    }
    return cached;
}

// Keep freed scratch in the stream-ordered pool instead of returning it to the OS at every synchronisation
// (the default release threshold is 0, which makes every call re-map its gigabytes of temporaries).
static void tune_pool_once() {
    static bool done = false;
    if (done) return;
    done = true;
    int dev = 0;
    cudaMemPool_t pool;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    (void)cudaGetLastError();
}

int scratch_alloc(void **p, size_t nbytes, cudaStream_t s) {
    tune_pool_once();
    if (nbytes == 0) nbytes = 16;
    B2S_CUDA(cudaMallocAsync(p, nbytes, s));
    return B2S_OK;
}

int scratch_free(void *p, cudaStream_t s) {
    if (p) B2S_CUDA(cudaFreeAsync(p, s));
    return B2S_OK;
}

}  // namespace b2s

using namespace b2s;

extern "C" {

int b2s_abi_version(void) { return B2S_ABI_VERSION; }

const char *b2s_last_error(void) { return g_err; }

int64_t b2s_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int b2s_device_count(int *count) {
    B2S_REQUIRE(count != nullptr, B2S_ERR_INVALID, "b2s_device_count: count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
        (void)cudaGetLastError();
        return B2S_ERR_NO_DEVICE;
    }
    *count = n;
    return B2S_OK;
}

int b2s_device_info(int device, char *name_buf, size_t name_len, int *sm, int *n_sms, size_t *hbm_bytes,
                    size_t *l2_bytes) {
    cudaDeviceProp prop;
    B2S_CUDA(cudaGetDeviceProperties(&prop, device));
    if (name_buf && name_len) {
        strncpy(name_buf, prop.name, name_len - 1);
        name_buf[name_len - 1] = 0;
    }
    if (sm) *sm = prop.major * 10 + prop.minor;
    if (n_sms) *n_sms = prop.multiProcessorCount;
    if (hbm_bytes) *hbm_bytes = prop.totalGlobalMem;
    if (l2_bytes) *l2_bytes = (size_t)prop.l2CacheSize;
    return B2S_OK;
}

int b2s_set_device(int device) {
    B2S_CUDA(cudaSetDevice(device));
    return B2S_OK;
}

int b2s_malloc(void **dev_ptr, size_t nbytes) {
    B2S_REQUIRE(dev_ptr != nullptr, B2S_ERR_INVALID, "b2s_malloc: dev_ptr is NULL");
    B2S_CUDA(cudaMalloc(dev_ptr, nbytes ? nbytes : 16));
    return B2S_OK;
}

int b2s_free(void *dev_ptr) {
    if (dev_ptr) B2S_CUDA(cudaFree(dev_ptr));
    return B2S_OK;
}

int b2s_host_register(void *host_ptr, size_t nbytes) {
    if (nbytes == 0) return B2S_OK;
    B2S_CUDA(cudaHostRegister(host_ptr, nbytes, cudaHostRegisterDefault));
    return B2S_OK;
}

int b2s_host_unregister(void *host_ptr) {
    B2S_CUDA(cudaHostUnregister(host_ptr));
    return B2S_OK;
}

int b2s_memcpy_h2d(void *dst_dev, const void *src_host, size_t nbytes, void *stream) {
    if (nbytes) B2S_CUDA(cudaMemcpyAsync(dst_dev, src_host, nbytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return B2S_OK;
}

int b2s_memcpy_d2h(void *dst_host, const void *src_dev, size_t nbytes, void *stream) {
    if (nbytes) B2S_CUDA(cudaMemcpyAsync(dst_host, src_dev, nbytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return B2S_OK;
}

int b2s_memset(void *dst_dev, int value, size_t nbytes, void *stream) {
    if (nbytes) B2S_CUDA(cudaMemsetAsync(dst_dev, value, nbytes, (cudaStream_t)stream));
    return B2S_OK;
}

int b2s_stream_sync(void *stream) {
    B2S_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return B2S_OK;
}

}  // extern "C"
