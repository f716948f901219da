// Per-GPU context of libvacmap_b200: stream, growable device arenas, tables.
#pragma once
#include "../../include/vacmap_b200.h"
#include "vm_common.cuh"
#include <cstring>
#include <string>
#include <vector>

struct VmDevBuf {
    void *p = nullptr;
    size_t cap = 0;
    bool borrowed = false;   // alias of another context's buffer (worker contexts share the score tables)
    void alias(const VmDevBuf &o) { if (!borrowed) release(); p = o.p; cap = o.cap; borrowed = true; }
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (borrowed) return cudaErrorInvalidValue;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p && !borrowed) cudaFree(p);
        p = nullptr;
        cap = 0;
        borrowed = false;
    }
    template <typename T> T *as() const { return (T *)p; }
};

// page-locked host staging buffer (D2H / H2D at full PCIe rate, no hidden bounce copy)
struct VmPinnedBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release()
    {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T *as() const { return (T *)p; }
};

// Device -> host through a page-locked bounce buffer, then a blocking wait.  A pageable destination would make the
// driver wait for the stream itself -- spinning on a host core for as long as the kernels in front of the copy run.
static inline cudaError_t vm_d2h_sync(VmPinnedBuf &pin, void *dst, const void *src, size_t bytes, cudaStream_t stream,
                                      void *dst2 = nullptr, const void *src2 = nullptr, size_t bytes2 = 0)
{
    cudaError_t e = pin.ensure(bytes + bytes2 + 16);
    if (e != cudaSuccess) return e;
    if (bytes) e = cudaMemcpyAsync(pin.p, src, bytes, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess && bytes2) e = cudaMemcpyAsync((char *)pin.p + bytes, src2, bytes2, cudaMemcpyDeviceToHost, stream);
    if (e != cudaSuccess) return e;
    e = vm_stream_sync(stream);
    if (e != cudaSuccess) return e;
    if (bytes) memcpy(dst, pin.p, bytes);
    if (bytes2) memcpy(dst2, (char *)pin.p + bytes, bytes2);
    return cudaSuccess;
}

// state of the chaining stage (stage-level upload / run / download, and the pipeline's device path)
struct VmChainState {
    bool loaded = false;
    vm_chain_params prm{};
    int64_t n_reads = 0;
    int64_t total = 0;
    std::vector<int64_t> off;            // start of each read's anchors (n_reads + 1 for the stage-level call)
    std::vector<int32_t> cnt;            // anchors per read
    std::vector<int32_t> read_len, cnt_len;
    std::vector<int64_t> gmax_host;
    VmDevBuf rows, off_dev, cnt_dev, anch, perm, sorted, sorted_rows, S, P, S_arg, gmax, opcount, ids, gcl, rgl,
        fast_scratch, fast_off, sort_scratch, pre_n_dev, head_dev;      // pre_n / head: carried prefix of the linked DP (variant 3)
    std::vector<int32_t> used_fast;
    std::vector<int64_t> opcount_host;
    VmPinnedBuf pin;                     // bounce buffer of the small per-read results read back between launches
    double opcount_last = 0;             // predecessor evaluations of the last vm_chain_core call (the reference's `opcount`, summed)
    float ms[4] = {0, 0, 0, 0};
};

struct vm_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {};
    // side streams for launches that are one slow warp deep (the capacity classes of the chain kernels): run side by side
    static const int kSide = 6;
    cudaStream_t side[kSide] = {};
    cudaEvent_t side_done[kSide] = {}, side_go = nullptr;
    std::string err;
    int64_t launches = 0;
    int sm_count = 0;
    // tables
    VmDevBuf extra, readgapcost, log2cache;
    int64_t n_extra = 0, n_readgapcost = 0, n_log2cache = 0;
    VmChainState chain;
    // persistent alignment backend (vm_backend_cuda.cu); destroyed through backend_free
    void *backend = nullptr;
    void (*backend_free)(void *) = nullptr;
    // worker contexts of the pipelined batch driver: own stream, chain state and backend; tables aliased
    std::vector<vm_ctx *> kids;
    // pool of worker threads serving this context's alignment jobs (vm_capi_align.cu); stopped before the kids go
    void *pool = nullptr;
    void (*pool_free)(void *) = nullptr;
};

// worker context i of `parent` (created on first use; its tables alias the parent's)
vm_ctx *vm_ctx_worker(vm_ctx *parent, int i);

// chaining core on device-resident anchors (vm_api.cu), used by the pipeline backend
int vm_chain_prepare(vm_ctx *c, int64_t n_reads, int64_t span, const std::vector<int64_t> &start, const std::vector<int32_t> &cnt,
                     bool want_rows);
int vm_chain_core(vm_ctx *c, const vm_chain_params &prm, const VmAnchor *d_anch, const std::vector<int64_t> &start,
                  const std::vector<int32_t> &cnt, const std::vector<int32_t> &read_len, const std::vector<int32_t> &cnt_len,
                  const std::vector<int> &ids, int64_t *sorted_rows_dev, std::vector<int32_t> *used_fast, float *ms4,
                  bool presorted = false, bool force_fast = false);
