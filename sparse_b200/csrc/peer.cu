// peer.cu -- copy-engine exchange of the row-sharded dense operand between the GPUs of one box.
//
// The multi-GPU product C_r = A_r @ B needs every rank to see all of B before its K1 launch.  An SM-based
// collective (NCCL's all-gather kernels) shares SMs and issue slots with the DRAM-bound K1 of the previous step that
// it is meant to hide behind; a copy-engine transfer does not.  Every rank allocates its shard with cudaMalloc,
// exports it through CUDA IPC, maps its peers' shards, and each step PULLS the world-1 remote shards over
// NVLink/NVSwitch with cudaMemcpyAsync (DMA, no kernel) into its own full-size buffer.  Starting the ring at rank+1
// keeps the ranks on different source GPUs at any moment.
#include "common.cuh"

using namespace b2s;

extern "C" {

int b2s_peer_alloc(void **dev_ptr, int64_t nbytes) {
    B2S_REQUIRE(dev_ptr != nullptr && nbytes >= 0, B2S_ERR_INVALID, "b2s_peer_alloc: bad arguments");
    B2S_CUDA(cudaMalloc(dev_ptr, nbytes > 0 ? (size_t)nbytes : 16));
    return B2S_OK;
}

int b2s_peer_free(void *dev_ptr) {
    if (dev_ptr) B2S_CUDA(cudaFree(dev_ptr));
    return B2S_OK;
}

int b2s_peer_export(void *dev_ptr, void *handle64) {
    B2S_REQUIRE(dev_ptr != nullptr && handle64 != nullptr, B2S_ERR_INVALID, "b2s_peer_export: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    B2S_CUDA(cudaIpcGetMemHandle(&h, dev_ptr));
    memcpy(handle64, &h, sizeof(h));
    return B2S_OK;
}

int b2s_peer_open(const void *handle64, void **dev_ptr) {
    B2S_REQUIRE(dev_ptr != nullptr && handle64 != nullptr, B2S_ERR_INVALID, "b2s_peer_open: NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    B2S_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return B2S_OK;
}

int b2s_peer_close(void *dev_ptr) {
    if (dev_ptr) B2S_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return B2S_OK;
}

int b2s_peer_gather(void *dst_full_dev, const void *const *shards_dev, int world, int rank, int64_t shard_bytes,
                    void *const *streams, int n_streams) {
    B2S_REQUIRE(dst_full_dev != nullptr && shards_dev != nullptr && world >= 1 && rank >= 0 && rank < world &&
                    shard_bytes >= 0 && streams != nullptr && n_streams >= 1,
                B2S_ERR_INVALID, "b2s_peer_gather: bad arguments");
    for (int k = 1; k <= world; ++k) {
        const int src = (rank + k) % world;  // the local shard goes last
        cudaStream_t s = (cudaStream_t)streams[(k - 1) % n_streams];
        B2S_CUDA(cudaMemcpyAsync((char *)dst_full_dev + (size_t)src * (size_t)shard_bytes, shards_dev[src],
                                 (size_t)shard_bytes, cudaMemcpyDeviceToDevice, s));
    }
    return B2S_OK;
}

}  // extern "C"
