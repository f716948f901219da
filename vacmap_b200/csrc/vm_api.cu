// C ABI of libvacmap_b200.so -- context, tables, and the chaining stage (see include/vacmap_b200.h).
#include "vm_ctx.cuh"
#include "vm_chain.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <malloc.h>

extern "C" {

int vm_abi_version(void) { return 9; }

int vm_host_alloc(vm_ctx *c, int64_t bytes, void **out)
{
    if (!c || !out || bytes < 0) return VM_ERR_ARG;
    *out = nullptr;
    cudaSetDevice(c->device);
    const cudaError_t e = cudaHostAlloc(out, (size_t)(bytes > 0 ? bytes : 1), cudaHostAllocPortable);
    if (e != cudaSuccess) { c->err = std::string("cudaHostAlloc: ") + cudaGetErrorString(e); cudaGetLastError(); return VM_ERR_NOMEM; }
    return VM_OK;
}

int vm_host_free(vm_ctx *c, void *p)
{
    // ctx may be NULL (a buffer outliving its context)
    if (p && cudaFreeHost(p) != cudaSuccess) { cudaGetLastError(); if (c) c->err = "cudaFreeHost failed"; return VM_ERR_CUDA; }
    return VM_OK;
}

int vm_host_register(vm_ctx *c, void *p, int64_t bytes)
{
    if (!c || !p || bytes <= 0) return VM_ERR_ARG;
    cudaSetDevice(c->device);
    const cudaError_t e = cudaHostRegister(p, (size_t)bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) { c->err = std::string("cudaHostRegister: ") + cudaGetErrorString(e); cudaGetLastError(); return VM_ERR_CUDA; }
    return VM_OK;
}

int vm_host_unregister(vm_ctx *c, void *p)
{
    if (!c || !p) return VM_ERR_ARG;
    if (cudaHostUnregister(p) != cudaSuccess) { cudaGetLastError(); c->err = "cudaHostUnregister failed"; return VM_ERR_CUDA; }
    return VM_OK;
}

int vm_ctx_create(int device, vm_ctx **out)
{
    if (!out) return VM_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) return VM_ERR_NO_DEVICE;
    if (device < 0 || device >= ndev) return VM_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return VM_ERR_CUDA;
    // Batches allocate and free large host arrays (job lists, result arenas) every call: keep them on the
    // heap instead of fresh mmaps, so they are not page-faulted in again for every batch.
    static const bool heap_tuned = [] { mallopt(M_MMAP_THRESHOLD, 1 << 30); mallopt(M_TRIM_THRESHOLD, 1 << 30); return true; }();
    (void)heap_tuned;
    vm_ctx *c = new vm_ctx();
    c->device = device;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return VM_ERR_CUDA;
    }
    for (int i = 0; i < 8; ++i) cudaEventCreateWithFlags(&c->ev[i], cudaEventBlockingSync);
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    *out = c;
    return VM_OK;
}

void vm_ctx_destroy(vm_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    vm_stream_sync(c->stream);
    if (c->pool && c->pool_free) c->pool_free(c->pool);
    c->pool = nullptr;
    for (vm_ctx *k : c->kids) vm_ctx_destroy(k);
    c->kids.clear();
    if (c->backend && c->backend_free) c->backend_free(c->backend);
    c->backend = nullptr;
    VmChainState &s = c->chain;
    VmDevBuf *bufs[] = {&s.rows, &s.off_dev, &s.cnt_dev, &s.anch, &s.perm, &s.sorted, &s.sorted_rows, &s.S, &s.P,
                        &s.S_arg, &s.gmax, &s.opcount, &s.ids, &s.gcl, &s.rgl, &s.fast_scratch, &s.fast_off, &s.sort_scratch,
                        &c->extra, &c->readgapcost, &c->log2cache};
    for (VmDevBuf *b : bufs) b->release();
    for (int i = 0; i < 8; ++i) cudaEventDestroy(c->ev[i]);
    for (int i = 0; i < vm_ctx::kSide; ++i) {
        if (c->side[i]) cudaStreamDestroy(c->side[i]);
        if (c->side_done[i]) cudaEventDestroy(c->side_done[i]);
    }
    if (c->side_go) cudaEventDestroy(c->side_go);
    cudaStreamDestroy(c->stream);
    delete c;
}

const char *vm_last_error(vm_ctx *c) { return c ? c->err.c_str() : "null ctx"; }

int64_t vm_kernel_launches(vm_ctx *c) { return c ? c->launches : 0; }

int vm_set_tables(vm_ctx *c, const float *extra, int64_t n_extra, const float *readgapcost,
                  int64_t n_readgapcost, const double *log2cache, int64_t n_log2cache)
{
    if (!c || !extra || !readgapcost || !log2cache || n_extra < 2 || n_readgapcost < 1 || n_log2cache < 2)
        return VM_ERR_ARG;
    cudaSetDevice(c->device);
    VM_CUDA_OK(c, c->extra.ensure(n_extra * sizeof(float)));
    VM_CUDA_OK(c, c->readgapcost.ensure(n_readgapcost * sizeof(float)));
    VM_CUDA_OK(c, c->log2cache.ensure(n_log2cache * sizeof(double)));
    VM_CUDA_OK(c, cudaMemcpyAsync(c->extra.p, extra, n_extra * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    VM_CUDA_OK(c, cudaMemcpyAsync(c->readgapcost.p, readgapcost, n_readgapcost * sizeof(float),
                                  cudaMemcpyHostToDevice, c->stream));
    VM_CUDA_OK(c, cudaMemcpyAsync(c->log2cache.p, log2cache, n_log2cache * sizeof(double),
                                  cudaMemcpyHostToDevice, c->stream));
    VM_CUDA_OK(c, vm_stream_sync(c->stream));
    c->n_extra = n_extra;
    c->n_readgapcost = n_readgapcost;
    c->n_log2cache = n_log2cache;
    return VM_OK;
}

} // extern "C"

vm_ctx *vm_ctx_worker(vm_ctx *parent, int i)
{
    // the table never reallocates once workers hold pointers into it (a later submit may add workers while earlier
    // ones are running): room for every worker a context can have is reserved with the first one
    if (parent->kids.capacity() < 256) parent->kids.reserve(256);
    if (i < 0 || i >= 256) return nullptr;
    while ((int)parent->kids.size() <= i) {
        vm_ctx *k = nullptr;
        if (vm_ctx_create(parent->device, &k) != VM_OK) return nullptr;
        parent->kids.push_back(k);
    }
    vm_ctx *k = parent->kids[i];
    // the score tables are aliases of the parent's; rewritten only when they changed (vm_set_tables reallocated them), so a
    // worker that is in the middle of a chunk never sees its pointers touched by a concurrent submit
    if (k->extra.p != parent->extra.p || k->readgapcost.p != parent->readgapcost.p || k->log2cache.p != parent->log2cache.p ||
        k->n_extra != parent->n_extra) {
        k->extra.alias(parent->extra);
        k->readgapcost.alias(parent->readgapcost);
        k->log2cache.alias(parent->log2cache);
        k->n_extra = parent->n_extra;
        k->n_readgapcost = parent->n_readgapcost;
        k->n_log2cache = parent->n_log2cache;
    }
    return k;
}

// gapcost_list as built inside the reference njit functions with libm log2
// (global :24843-24846; local :27317-27322)
static void vm_host_gapcost(int kmersize, int maxdiff, bool local, std::vector<double> &out)
{
    out.assign(maxdiff + 1, 0.0);
    for (int g = 1; g <= maxdiff; ++g) {
        double lg = std::log2((double)g);
        if (!local || g <= 10) out[g] = 0.01 * kmersize * g + 0.5 * lg;
        else out[g] = 0.01 * kmersize * g + 2 * lg;
    }
}

// large_readgapcost_list (:28270-28275), float32
static void vm_host_large_readgap(int maxgap, int large_readgap, std::vector<float> &out)
{
    out.assign(maxgap + 1, 0.0f);
    for (int r = 1; r <= maxgap; ++r) {
        if (large_readgap <= r) out[r] = (float)(0.5 * r);
        else out[r] = (float)(0.1 * std::log2((double)(r + 1)));
    }
}

// shared-memory capacity classes of the DP kernels (12 B per anchor): fine steps where the reads are (a 15 kb read has
// 0.5-2 k anchors per DP), so that a class wastes little of the SM's shared memory and more reads are resident
static const int kCaps[] = {128, 256, 384, 512, 768, 1024, 1536, 2048, 3072, 4096, 6144, 8192, VM_CHAIN_SMEM_CAP};
static const int kNumCaps = 13;

// Fill VmChainArgs from ctx + state (device pointers).
static int vm_chain_args(vm_ctx *c, VmChainState &s, const vm_chain_params &p, VmChainArgs &A)
{
    if (c->n_extra == 0) { c->err = "vm_set_tables must be called first"; return VM_ERR_STATE; }
    if (p.maxdiff + 1 > VM_GCL_MAX) { c->err = "maxdiff too large"; return VM_ERR_ARG; }
    std::vector<double> gcl;
    vm_host_gapcost(p.kmersize, p.maxdiff, p.variant == 1 || p.variant == 2, gcl);
    VM_CUDA_OK(c, s.gcl.ensure(gcl.size() * sizeof(double)));
    VM_CUDA_OK(c, cudaMemcpyAsync(s.gcl.p, gcl.data(), gcl.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    A.rgcost = nullptr;
    A.n_rg = 0;
    std::vector<float> rg;
    if (p.variant == 1) {
        A.rgcost = c->readgapcost.as<float>();
        A.n_rg = (int)c->n_readgapcost;
    } else if (p.variant == 4) {
        // asm mode's own readgapcost_list (mammap_asm.py:16536-16538): float32[100], 0.1 log2(r)
        if (p.maxgap + 1 > 100) { c->err = "maxgap too large for the linked local DP"; return VM_ERR_ARG; }
        rg.assign(100, 0.0f);
        for (int r = 1; r < 100; ++r) rg[(size_t)r] = (float)(0.1 * std::log2((double)r));
        VM_CUDA_OK(c, s.rgl.ensure(rg.size() * sizeof(float)));
        VM_CUDA_OK(c, cudaMemcpyAsync(s.rgl.p, rg.data(), rg.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        A.rgcost = s.rgl.as<float>();
        A.n_rg = (int)rg.size();
    } else if (p.variant == 2) {
        if (p.maxgap + 1 > VM_RGL_MAX) { c->err = "maxgap too large for the local DP"; return VM_ERR_ARG; }
        vm_host_large_readgap(p.maxgap, p.large_readgap, rg);
        VM_CUDA_OK(c, s.rgl.ensure(rg.size() * sizeof(float)));
        VM_CUDA_OK(c, cudaMemcpyAsync(s.rgl.p, rg.data(), rg.size() * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        A.rgcost = s.rgl.as<float>();
        A.n_rg = (int)rg.size();
    }
    // (the staging vectors above are pageable: cudaMemcpyAsync has staged them when it returns, no sync needed)
    A.anchors = s.sorted.as<VmAnchor>();
    A.off = s.off_dev.as<int64_t>();
    A.cnt = s.cnt_dev.as<int32_t>();
    A.S = s.S.as<double>();
    A.P = s.P.as<int32_t>();
    A.S_arg = s.S_arg.as<int32_t>();
    A.gmax = s.gmax.as<int64_t>();
    A.opcount = s.opcount.as<int64_t>();
    A.extra = c->extra.as<float>();
    A.extra_size = c->n_extra - 1;
    A.log2cache = c->log2cache.as<double>();
    A.log2cache_size = c->n_log2cache - 1;
    A.gapcost_list = s.gcl.as<double>();
    A.skipcost = p.skipcost;
    A.maxdiff = p.maxdiff;
    A.maxgap = p.maxgap;
    A.max_factor = p.max_factor;
    A.pre_n = s.pre_n_dev.as<int32_t>();
    A.head = s.head_dev.as<double>();
    return VM_OK;
}

// Chaining core on DEVICE input: unsorted anchors of read r at d_anch[start[r] .. start[r]+cnt[r]).
// Runs argsort replay + exact DP (+ fast DP where hit2work_1 / the local DP would) for the reads in
// `ids`; results land in s.sorted / s.S / s.P / s.S_arg (same offsets) and s.gmax[r].  The device
// copies of start/cnt must already be in s.off_dev / s.cnt_dev (vm_chain_prepare).  cnt_len[r] bounds
// the integer score of a chain of read r (sizes the fast DP's per-score counters).
int vm_chain_core(vm_ctx *c, const vm_chain_params &prm, const VmAnchor *d_anch, const std::vector<int64_t> &start,
                  const std::vector<int32_t> &cnt, const std::vector<int32_t> &read_len, const std::vector<int32_t> &cnt_len,
                  const std::vector<int> &ids, int64_t *sorted_rows_dev, std::vector<int32_t> *used_fast, float *ms4,
                  bool presorted, bool force_fast)
{
    (void)start;
    VmChainState &s = c->chain;
    VmChainArgs A;
    int rc = vm_chain_args(c, s, prm, A);
    if (rc != VM_OK) return rc;
    const bool by_end = prm.variant == 1 || prm.variant == 2;
    const bool linked = prm.variant >= 3;      // asm mode: jobs, not reads -- no n / read_len short cut to the fast DP
    std::vector<std::vector<int>> cls(kNumCaps + 1);
    std::vector<int> fast_ids;
    bool may_bail = false;
    for (int r : ids) {
        const int64_t n = cnt[r];
        if (n <= 0) continue;
        // hit2work_1 :23570 -- n / read_len > 5 goes straight to the fast DP (global only)
        if (!linked && (force_fast || (!by_end && (double)n / (double)read_len[r] > 5.0))) { fast_ids.push_back(r); continue; }
        int k = 0;
        while (k < kNumCaps && n > kCaps[k]) ++k;
        cls[k].push_back(r);
        // global: opcount/i > 1000 needs i > 2000; local: opcount > 100000 needs n(n-1)/2 > 100000
        if (prm.variant != 4 && (by_end ? n > 440 : n > 2000)) may_bail = true;      // variant 4 never bails out
    }
    std::vector<int> ids_host;
    std::vector<int> cls_start(kNumCaps + 2, 0);
    for (int k = 0; k <= kNumCaps; ++k) {
        cls_start[k] = (int)ids_host.size();
        ids_host.insert(ids_host.end(), cls[k].begin(), cls[k].end());
    }
    cls_start[kNumCaps + 1] = (int)ids_host.size();
    const int n_exact = (int)ids_host.size();
    ids_host.insert(ids_host.end(), fast_ids.begin(), fast_ids.end());   // fast-path reads still need sorting
    VM_CUDA_OK(c, s.ids.ensure(ids_host.size() * 4 + 64));
    if (!ids_host.empty())
        VM_CUDA_OK(c, cudaMemcpyAsync(s.ids.p, ids_host.data(), ids_host.size() * 4, cudaMemcpyHostToDevice, c->stream));

    cudaEvent_t *ev = c->ev;
    VM_CUDA_OK(c, cudaEventRecord(ev[1], c->stream));
    if (presorted) {
        // stage-level entry with anchors already in DP order (the reference functions take sorted input): no argsort
        int64_t span = 0;
        for (size_t r = 0; r < cnt.size(); ++r) span = std::max<int64_t>(span, start[r] + cnt[r]);
        if (span > 0)
            VM_CUDA_OK(c, cudaMemcpyAsync(s.sorted.p, d_anch, (size_t)span * sizeof(VmAnchor), cudaMemcpyDeviceToDevice, c->stream));
    }
    // Every capacity class is its own pair of launches (argsort replay, then the DP), and each launch lasts as long as its
    // slowest read -- a launch of 2 reads takes what one of 2 000 takes (ncu: 0.8 ms either way).  The classes are
    // independent, so they go to side streams and run side by side: the stage costs its slowest class, not their sum.
    if (!c->side[0]) {
        for (int i = 0; i < vm_ctx::kSide; ++i) {
            VM_CUDA_OK(c, cudaStreamCreateWithFlags(&c->side[i], cudaStreamNonBlocking));
            VM_CUDA_OK(c, cudaEventCreateWithFlags(&c->side_done[i], cudaEventDisableTiming));
        }
        VM_CUDA_OK(c, cudaEventCreateWithFlags(&c->side_go, cudaEventDisableTiming));
    }
    VM_CUDA_OK(c, cudaEventRecord(c->side_go, c->stream));
    for (int i = 0; i < vm_ctx::kSide; ++i) VM_CUDA_OK(c, cudaStreamWaitEvent(c->side[i], c->side_go, 0));
    int rr = 0;
    for (int k = kNumCaps; k >= 0; --k) {          // the classes of the longest reads first
        const int n_k = cls_start[k + 1] - cls_start[k];
        if (n_k == 0) continue;
        cudaStream_t st = c->side[rr++ % vm_ctx::kSide];
        if (!presorted) {
            const bool smem_sort = k < kNumCaps && kCaps[k] <= VM_SORT_SMEM_CAP;
            c->launches += vm_launch_sort_anchors(d_anch, s.off_dev.as<int64_t>(), s.cnt_dev.as<int32_t>(),
                                                  s.ids.as<int>() + cls_start[k], n_k, smem_sort ? kCaps[k] : 0, smem_sort, by_end ? 1 : 0,
                                                  s.perm.as<int32_t>(), s.sort_scratch.as<int32_t>(), s.sorted.as<VmAnchor>(),
                                                  sorted_rows_dev, st);
        }
        static const int no_smem_below = getenv("VM_CHAIN_GLOBAL_BELOW") ? atoi(getenv("VM_CHAIN_GLOBAL_BELOW")) : 0;   // experiment knob
        const bool smem = k < kNumCaps && !(kCaps[k] <= no_smem_below);
        c->launches += vm_launch_chain_exact(prm.variant, A, s.ids.as<int>() + cls_start[k], n_k, smem ? kCaps[k] : 0, smem, st);
    }
    if (!fast_ids.empty() && !presorted)
        c->launches += vm_launch_sort_anchors(d_anch, s.off_dev.as<int64_t>(), s.cnt_dev.as<int32_t>(), s.ids.as<int>() + n_exact,
                                              (int)fast_ids.size(), 0, false, by_end ? 1 : 0, s.perm.as<int32_t>(),
                                              s.sort_scratch.as<int32_t>(), s.sorted.as<VmAnchor>(), sorted_rows_dev, c->side[rr++ % vm_ctx::kSide]);
    for (int i = 0; i < vm_ctx::kSide; ++i) {
        VM_CUDA_OK(c, cudaEventRecord(c->side_done[i], c->side[i]));
        VM_CUDA_OK(c, cudaStreamWaitEvent(c->stream, c->side_done[i], 0));
    }
    VM_CUDA_OK(c, cudaEventRecord(ev[2], c->stream));
    VM_CUDA_OK(c, cudaEventRecord(ev[3], c->stream));
    // reads whose exact DP bailed out (opcount rule) join the fast list
    if (may_bail) {
        s.gmax_host.resize(cnt.size());
        VM_CUDA_OK(c, vm_d2h_sync(s.pin, s.gmax_host.data(), s.gmax.p, cnt.size() * 8, c->stream));
        for (int t = 0; t < n_exact; ++t)
            if (s.gmax_host[ids_host[t]] < 0) fast_ids.push_back(ids_host[t]);
    }
    if (!fast_ids.empty()) {
        std::vector<int64_t> soff(fast_ids.size() + 1, 0);
        for (size_t t = 0; t < fast_ids.size(); ++t) {
            const int r = fast_ids[t];
            soff[t + 1] = soff[t] + 2LL * cnt[r] + (int64_t)cnt_len[r] + 64;
            if (used_fast) (*used_fast)[r] = 1;
        }
        VM_CUDA_OK(c, s.fast_scratch.ensure((size_t)soff.back() * 8));
        VM_CUDA_OK(c, s.fast_off.ensure(soff.size() * 8 + fast_ids.size() * 4));
        int *fids_dev = (int *)((char *)s.fast_off.p + soff.size() * 8);
        VM_CUDA_OK(c, cudaMemcpyAsync(s.fast_off.p, soff.data(), soff.size() * 8, cudaMemcpyHostToDevice, c->stream));
        VM_CUDA_OK(c, cudaMemcpyAsync(fids_dev, fast_ids.data(), fast_ids.size() * 4, cudaMemcpyHostToDevice, c->stream));
        c->launches += vm_launch_chain_fast(prm.variant, A, prm.fast_t, fids_dev, (int)fast_ids.size(),
                                            s.fast_scratch.as<long long>(), s.fast_off.as<int64_t>(), c->stream);
    }
    VM_CUDA_OK(c, cudaEventRecord(ev[4], c->stream));
    // the DPs' own count of predecessor evaluations (the reference's `opcount`): B_chain_alg of the roofline report
    s.opcount_host.resize(cnt.size());
    VM_CUDA_OK(c, vm_d2h_sync(s.pin, s.opcount_host.data(), s.opcount.p, cnt.size() * 8, c->stream));
    VM_CUDA_OK(c, cudaGetLastError());
    s.opcount_last = 0;
    for (int t = 0; t < n_exact; ++t) s.opcount_last += (double)s.opcount_host[(size_t)ids_host[(size_t)t]];
    if (ms4) {
        ms4[0] = 0;
        for (int t = 1; t < 4; ++t) cudaEventElapsedTime(&ms4[t], ev[t], ev[t + 1]);
    }
    return VM_OK;
}

// size the chaining state for `span` anchor slots and `n_reads` reads; upload start / cnt
int vm_chain_prepare(vm_ctx *c, int64_t n_reads, int64_t span, const std::vector<int64_t> &start, const std::vector<int32_t> &cnt,
                     bool want_rows)
{
    VmChainState &s = c->chain;
    const size_t T = (size_t)std::max<int64_t>(span, 1);
    VM_CUDA_OK(c, s.off_dev.ensure((size_t)(n_reads + 1) * 8));
    VM_CUDA_OK(c, s.cnt_dev.ensure((size_t)(n_reads + 1) * 4));
    VM_CUDA_OK(c, s.perm.ensure(T * 4));
    VM_CUDA_OK(c, s.sorted.ensure(T * 16));
    if (want_rows) VM_CUDA_OK(c, s.sorted_rows.ensure(T * 32));
    VM_CUDA_OK(c, s.S.ensure(T * 8));
    VM_CUDA_OK(c, s.P.ensure(T * 4));
    VM_CUDA_OK(c, s.S_arg.ensure(T * 4));
    VM_CUDA_OK(c, s.gmax.ensure((size_t)(n_reads + 1) * 8));
    VM_CUDA_OK(c, s.opcount.ensure((size_t)(n_reads + 1) * 8));
    VM_CUDA_OK(c, s.sort_scratch.ensure(T * 12));
    if (n_reads > 0) {
        VM_CUDA_OK(c, cudaMemcpyAsync(s.off_dev.p, start.data(), (size_t)n_reads * 8, cudaMemcpyHostToDevice, c->stream));
        VM_CUDA_OK(c, cudaMemcpyAsync(s.cnt_dev.p, cnt.data(), (size_t)n_reads * 4, cudaMemcpyHostToDevice, c->stream));
        VM_CUDA_OK(c, cudaMemsetAsync(s.gmax.p, 0xff, (size_t)n_reads * 8, c->stream));   // -1: not chained
        VM_CUDA_OK(c, vm_stream_sync(c->stream));
    }
    return VM_OK;
}

extern "C" {

int vm_chain_global_upload(vm_ctx *c, const vm_chain_params *prm, int64_t n_reads, const int64_t *anchors,
                           const int64_t *off, const int32_t *read_len)
{
    if (!c) return VM_ERR_ARG;
    if (!prm || n_reads < 0 || !off || (n_reads > 0 && !read_len)) { c->err = "bad argument"; return VM_ERR_ARG; }
    if (n_reads >= (1LL << 31)) { c->err = "too many reads in one batch"; return VM_ERR_ARG; }
    cudaSetDevice(c->device);
    VmChainState &s = c->chain;
    s.loaded = false;
    s.prm = *prm;
    s.n_reads = n_reads;
    s.total = off[n_reads];
    if (s.total > 0 && !anchors) { c->err = "anchors is null"; return VM_ERR_ARG; }
    for (int64_t r = 0; r < n_reads; ++r) {
        if (off[r + 1] < off[r] || off[r + 1] - off[r] >= (1LL << 31) - 16) { c->err = "bad offsets"; return VM_ERR_ARG; }
    }
    s.off.assign(off, off + n_reads + 1);
    s.read_len.assign(read_len, read_len + n_reads);
    // the fast DP indexes a per-score counter by integer chain score (<= last read position + k):
    // size it from the larger of the declared read length and the largest anchor end actually present
    s.cnt_len.assign(n_reads, 0);
    s.cnt.assign(n_reads, 0);
    for (int64_t r = 0; r < n_reads; ++r) {
        int64_t mx = 0;
        for (int64_t t = off[r]; t < off[r + 1]; ++t) mx = std::max(mx, anchors[t * 4] + anchors[t * 4 + 3]);
        s.cnt_len[r] = (int32_t)std::min<int64_t>(std::max<int64_t>(mx, s.read_len[r]), INT32_MAX - 128);
        s.cnt[r] = (int32_t)(off[r + 1] - off[r]);
    }
    const size_t T = (size_t)std::max<int64_t>(s.total, 1);
    VM_CUDA_OK(c, s.rows.ensure(T * 32));
    VM_CUDA_OK(c, s.anch.ensure(T * 16));
    int rc = vm_chain_prepare(c, n_reads, s.total, s.off, s.cnt, true);
    if (rc != VM_OK) return rc;
    if (s.total > 0)
        VM_CUDA_OK(c, cudaMemcpyAsync(s.rows.p, anchors, (size_t)s.total * 32, cudaMemcpyHostToDevice, c->stream));
    VM_CUDA_OK(c, vm_stream_sync(c->stream));
    s.loaded = true;
    return VM_OK;
}

int vm_chain_global_run(vm_ctx *c, float *kernel_ms)
{
    if (!c) return VM_ERR_ARG;
    VmChainState &s = c->chain;
    if (!s.loaded) { c->err = "vm_chain_global_upload first"; return VM_ERR_STATE; }
    cudaSetDevice(c->device);
    const int n_reads = (int)s.n_reads;
    s.used_fast.assign(n_reads, 0);
    std::vector<int> ids(n_reads);
    for (int r = 0; r < n_reads; ++r) ids[r] = r;
    VM_CUDA_OK(c, cudaEventRecord(c->ev[0], c->stream));
    c->launches += vm_launch_pack(s.rows.as<int64_t>(), s.anch.as<VmAnchor>(), s.total, c->stream);
    VM_CUDA_OK(c, cudaEventRecord(c->ev[5], c->stream));
    int rc = vm_chain_core(c, s.prm, s.anch.as<VmAnchor>(), s.off, s.cnt, s.read_len, s.cnt_len, ids, s.sorted_rows.as<int64_t>(),
                           &s.used_fast, s.ms);
    if (rc != VM_OK) return rc;
    cudaEventElapsedTime(&s.ms[0], c->ev[0], c->ev[5]);
    if (kernel_ms) *kernel_ms = s.ms[0] + s.ms[1] + s.ms[2] + s.ms[3];
    return VM_OK;
}

int vm_chain_global_times(vm_ctx *c, float *ms4)
{
    if (!c || !ms4) return VM_ERR_ARG;
    memcpy(ms4, c->chain.ms, sizeof(float) * 4);
    return VM_OK;
}

int vm_chain_global_download(vm_ctx *c, int64_t *sorted, double *S, int32_t *P, int32_t *S_arg,
                             int64_t *g_max_index, int32_t *used_fast)
{
    if (!c) return VM_ERR_ARG;
    VmChainState &s = c->chain;
    if (!s.loaded) { c->err = "nothing to download"; return VM_ERR_STATE; }
    cudaSetDevice(c->device);
    const size_t T = (size_t)s.total;
    if (T > 0) {
        if (sorted) VM_CUDA_OK(c, cudaMemcpyAsync(sorted, s.sorted_rows.p, T * 32, cudaMemcpyDeviceToHost, c->stream));
        if (S) VM_CUDA_OK(c, cudaMemcpyAsync(S, s.S.p, T * 8, cudaMemcpyDeviceToHost, c->stream));
        if (P) VM_CUDA_OK(c, cudaMemcpyAsync(P, s.P.p, T * 4, cudaMemcpyDeviceToHost, c->stream));
        if (S_arg) VM_CUDA_OK(c, cudaMemcpyAsync(S_arg, s.S_arg.p, T * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    if (g_max_index && s.n_reads > 0)
        VM_CUDA_OK(c, cudaMemcpyAsync(g_max_index, s.gmax.p, (size_t)s.n_reads * 8, cudaMemcpyDeviceToHost, c->stream));
    VM_CUDA_OK(c, vm_stream_sync(c->stream));
    if (used_fast && s.n_reads > 0) memcpy(used_fast, s.used_fast.data(), (size_t)s.n_reads * 4);
    return VM_OK;
}

int vm_chain_global_batch(vm_ctx *c, const vm_chain_params *prm, int64_t n_reads, const int64_t *anchors,
                          const int64_t *off, const int32_t *read_len, int64_t *sorted, double *S, int32_t *P,
                          int32_t *S_arg, int64_t *g_max_index, int32_t *used_fast, float *kernel_ms)
{
    int rc = vm_chain_global_upload(c, prm, n_reads, anchors, off, read_len);
    if (rc != VM_OK) return rc;
    rc = vm_chain_global_run(c, kernel_ms);
    if (rc != VM_OK) return rc;
    return vm_chain_global_download(c, sorted, S, P, S_arg, g_max_index, used_fast);
}

int vm_chain_linked_batch(vm_ctx *c, const vm_chain_params *prm, int64_t n_jobs, const int64_t *anchors, const int64_t *off,
                          const int32_t *pre_n, const double *head, double *S, int32_t *P, int32_t *S_arg, int64_t *g_max_index,
                          int32_t *used_fast)
{
    if (!c) return VM_ERR_ARG;
    if (!prm || !off || n_jobs < 0 || (n_jobs > 0 && (!pre_n || !head || !S || !P))) { c->err = "bad argument"; return VM_ERR_ARG; }
    vm_chain_params p = *prm;
    p.variant = prm->variant == 4 ? 4 : 3;      // 4: the second-round twin linked_..._fine_list_all (:21505-21686)
    for (int64_t r = 0; r < n_jobs; ++r)
        if (pre_n[r] < 0 || pre_n[r] > off[r + 1] - off[r]) { c->err = "pre_n out of range"; return VM_ERR_ARG; }
    // "read length" of a job = what sizes the fast DP's per-score counters: the largest chain score it can reach
    // (carried scores + the read span of its anchors)
    std::vector<int32_t> rl((size_t)std::max<int64_t>(n_jobs, 1), 0);
    for (int64_t r = 0; r < n_jobs; ++r) {
        double carried = 0;
        for (int64_t t = 0; t < pre_n[r]; ++t) carried = std::max(carried, S[off[r] + t]);
        int64_t span = 0;
        for (int64_t t = off[r]; t < off[r + 1]; ++t) span = std::max(span, anchors[t * 4] + anchors[t * 4 + 3]);
        rl[(size_t)r] = (int32_t)std::min<double>((double)span + carried + 1064.0, (double)(INT32_MAX - 256));
    }
    int rc = vm_chain_global_upload(c, &p, n_jobs, anchors, off, rl.data());
    if (rc != VM_OK) return rc;
    VmChainState &s = c->chain;
    const size_t T = (size_t)s.total;
    VM_CUDA_OK(c, s.pre_n_dev.ensure((size_t)(n_jobs + 1) * 4));
    VM_CUDA_OK(c, s.head_dev.ensure((size_t)(n_jobs + 1) * 24));
    if (n_jobs > 0) {
        VM_CUDA_OK(c, cudaMemcpyAsync(s.pre_n_dev.p, pre_n, (size_t)n_jobs * 4, cudaMemcpyHostToDevice, c->stream));
        VM_CUDA_OK(c, cudaMemcpyAsync(s.head_dev.p, head, (size_t)n_jobs * 24, cudaMemcpyHostToDevice, c->stream));
    }
    if (T > 0) {      // the carried scores / negated back-pointers sit in front of every job's S / P
        VM_CUDA_OK(c, cudaMemcpyAsync(s.S.p, S, T * 8, cudaMemcpyHostToDevice, c->stream));
        VM_CUDA_OK(c, cudaMemcpyAsync(s.P.p, P, T * 4, cudaMemcpyHostToDevice, c->stream));
    }
    std::vector<int> ids((size_t)n_jobs);
    for (int64_t r = 0; r < n_jobs; ++r) ids[(size_t)r] = (int)r;
    c->launches += vm_launch_pack(s.rows.as<int64_t>(), s.anch.as<VmAnchor>(), s.total, c->stream);
    s.used_fast.assign((size_t)n_jobs, 0);
    rc = vm_chain_core(c, p, s.anch.as<VmAnchor>(), s.off, s.cnt, s.read_len, s.cnt_len, ids, nullptr, &s.used_fast, s.ms, true, false);
    if (rc != VM_OK) return rc;
    return vm_chain_global_download(c, nullptr, S, P, S_arg, g_max_index, used_fast);
}

} // extern "C"
