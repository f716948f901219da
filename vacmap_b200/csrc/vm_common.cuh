// Shared device/host helpers for the vacmap_b200 CUDA library (sm_100a).
#pragma once
#include <chrono>
#include <cstdlib>
#include <sched.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#define VM_FULL 0xffffffffu
#define VM_NOPRE (-9999999)

// A seed/anchor as the reference's 4-tuple (readpos, refpos, strand, len)
// (mammap_clrnano.py:23985) packed to 16 bytes so one LDG.128 / LDS.128 fetches it.
// refpos is the GLOBAL concatenated reference coordinate; GRCh38-sized references
// exceed 2^31, so it is carried as uint32 and widened to int64 for gap arithmetic.
struct __align__(16) VmAnchor {
    int32_t x;      // read position (start)
    uint32_t y;     // global reference position (leftmost)
    int32_t s;      // strand +1 / -1
    int32_t l;      // length
};

struct VmError {
    std::string msg;
};

#define VM_CUDA_OK(ctx, call)                                                        \
    do {                                                                             \
        cudaError_t _e = (call);                                                     \
        if (_e != cudaSuccess) {                                                     \
            char _b[512];                                                            \
            snprintf(_b, sizeof(_b), "%s:%d %s -> %s", __FILE__, __LINE__, #call,    \
                     cudaGetErrorString(_e));                                        \
            (ctx)->err = _b;                                                         \
            return VM_ERR_CUDA;                                                      \
        }                                                                            \
    } while (0)

// Wait for a stream without burning a host core: the default cudaStreamSynchronize spins, and with several worker
// threads waiting on their kernels at any moment that costs whole cores the host glue needs.  One blocking-sync
// event per host thread.
struct VmSyncStats { long long ns = 0, n = 0; };
static inline VmSyncStats &vm_sync_stats() { thread_local VmSyncStats s; return s; }      // per host thread: time spent waiting, waits

static inline cudaError_t vm_stream_sync_(cudaStream_t stream);
static inline cudaError_t vm_stream_sync(cudaStream_t stream)
{
    const auto t0 = std::chrono::steady_clock::now();
    const cudaError_t e = vm_stream_sync_(stream);
    VmSyncStats &s = vm_sync_stats();
    s.ns += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
    s.n += 1;
    return e;
}
static inline cudaError_t vm_stream_sync_(cudaStream_t stream)
{
    thread_local cudaEvent_t ev = nullptr;
    if (!ev) {
        const cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventBlockingSync | cudaEventDisableTiming);
        if (e != cudaSuccess) { ev = nullptr; return e; }
    }
    const cudaError_t e = cudaEventRecord(ev, stream);
    if (e != cudaSuccess) return e;
    // VM_SYNC_SPIN_US: poll (yielding the core between polls) for that many microseconds before blocking -- a blocking
    // wait is woken by an interrupt, which on a busy multi-GPU host can take far longer than the kernel waited for;
    // 0 = block at once, negative = poll until done
    static const long spin_us = getenv("VM_SYNC_SPIN_US") ? atol(getenv("VM_SYNC_SPIN_US")) : 0;
    if (spin_us != 0) {
        const auto t0 = std::chrono::steady_clock::now();
        for (;;) {
            const cudaError_t q = cudaEventQuery(ev);
            if (q != cudaErrorNotReady) return q;
            if (spin_us > 0 && std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count() >= spin_us) break;
            sched_yield();
        }
    }
    return cudaEventSynchronize(ev);
}

// Let a kernel use all opt-in dynamic shared memory.  The value is the same on every call, so concurrent
// worker threads launching the same kernel with different sizes cannot lower each other's limit.
template <typename F>
static inline void vm_smem_optin(F kernel)
{
    cudaFuncAttributes a;
    int dev = 0, mx = 0;
    if (cudaFuncGetAttributes(&a, kernel) != cudaSuccess || cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&mx, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess)
        return;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mx - (int)a.sharedSizeBytes);
}

// pairwise gap geometry shared by all chaining variants
// (reference: mammap_clrnano.py:24953-24984 global, 27418-27456 local)
__device__ __forceinline__ void vm_pair_gaps(const VmAnchor &ai, const VmAnchor &aj,
                                             int &bonus, int &readgap, long long &refgap)
{
    const int rg = ai.x - aj.x - aj.l;
    const long long yi = (long long)ai.y, yj = (long long)aj.y;
    if (rg < 0) {
        const int b = ai.x + ai.l - aj.x - aj.l;
        const int ov = aj.x + aj.l - ai.x;
        bonus = b;
        readgap = 0;
        if (ai.s == aj.s) {
            if (ai.s == 1) refgap = yi + ov - (yj + aj.l);
            else refgap = yj - (yi + b);
        } else {
            if (aj.s == -1) refgap = yi + ov - yj + 1;
            else refgap = yi + b - 1 - (yj + aj.l);
        }
    } else {
        bonus = ai.l;
        readgap = rg;
        if (ai.s == aj.s) {
            if (ai.s == 1) refgap = yi - yj - aj.l;
            else refgap = yj - yi - ai.l;
        } else {
            if (aj.s == -1) refgap = yi - yj + 1;
            else refgap = yi + ai.l - 1 - yj - aj.l;
        }
    }
}

// asm mode (mammap_asm.py:21795-21824): the linked DPs' older gap geometry -- no +-1 between opposite strands,
// overlaps handled through the non-overlapping length of anchor i
__device__ __forceinline__ void vm_pair_gaps_asm(const VmAnchor &ai, const VmAnchor &aj,
                                                 int &bonus, int &readgap, long long &refgap)
{
    const int rg = ai.x - aj.x - aj.l;
    const long long yi = (long long)ai.y, yj = (long long)aj.y;
    if (rg < 0) {
        const long long nos = ai.x - aj.x;
        bonus = ai.x + ai.l - aj.x - aj.l;
        readgap = 0;
        if (ai.s == aj.s) {
            if (ai.s == 1) refgap = yi - yj - nos;
            else refgap = yj + aj.l - nos - yi - ai.l;
        } else {
            if (aj.s == -1) refgap = yi + aj.l - nos - yj;
            else refgap = yi + ai.l - yj - nos;
        }
    } else {
        bonus = ai.l;
        readgap = rg;
        if (ai.s == aj.s) {
            if (ai.s == 1) refgap = yi - yj - aj.l;
            else refgap = yj - yi - ai.l;
        } else {
            if (aj.s == -1) refgap = yi - yj;
            else refgap = yi + ai.l - yj - aj.l;
        }
    }
}

__device__ __forceinline__ long long vm_llabs(long long v) { return v < 0 ? -v : v; }
