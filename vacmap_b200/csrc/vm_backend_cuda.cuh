// CUDA backend of the batch driver (vm_pipeline.hpp).
//
// One instance per vm_ctx, alive for the life of the context: every device arena and pinned
// staging buffer is reused across batches.  Consecutive hot loops hand their data over on the
// device (seeding -> sort -> global DP; re-seeding -> concat -> sort -> local DP); the host only
// receives what its glue needs (sorted anchors, S, P, S_arg, g_max; CIGAR ops compacted on the
// device) through page-locked buffers.  There is no CPU implementation of any stage here.
#pragma once
#include "vm_ctx.cuh"
#include "vm_chain.cuh"
#include "vm_index.cuh"
#include "vm_seed.cuh"
#include "vm_reseed.cuh"
#include "vm_align.cuh"
#include "vm_extract.cuh"
#include "vm_pipeline.hpp"
#include "vm_dglue.cuh"
#include <chrono>
#include <ctime>
#include <map>
#include <mutex>

using namespace vmp;

struct vm_index_handle {
    VmIndex *ix = nullptr;
    vmg::Contigs ctg;
};

#define BE_OK(call)                                                                                   \
    do {                                                                                              \
        cudaError_t _e = (call);                                                                      \
        if (_e != cudaSuccess) throw std::runtime_error(std::string(#call) + ": " + cudaGetErrorString(_e)); \
    } while (0)

// ---- small plumbing kernels ----
// reverse complement of every read of the batch (one thread per base)
__global__ void vm_revcomp_kernel(const uint8_t *__restrict__ fwd, const int64_t *__restrict__ off, int n_reads, int64_t total,
                                  uint8_t *__restrict__ rc)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int lo = 0, hi = n_reads;   // read containing base i
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= i) lo = mid; else hi = mid;
    }
    const int64_t b = off[lo], e = off[lo + 1];
    const uint8_t ch = fwd[e - 1 - (i - b)];
    rc[i] = ch == 'A' ? 'T' : ch == 'C' ? 'G' : ch == 'G' ? 'C' : ch == 'T' ? 'A' : 'N';
}

// gather variable-length segments: segment j = src[src_off[j] .. +len[j]) -> dst[dst_off[j] ..)
template <typename T>
__global__ void vm_gather_segments_kernel(const T *__restrict__ src, const int64_t *__restrict__ src_off,
                                          const int64_t *__restrict__ dst_off, const int32_t *__restrict__ len,
                                          T *__restrict__ dst)
{
    const int j = blockIdx.x;
    const int n = len[j];
    const T *s = src + src_off[j];
    T *d = dst + dst_off[j];
    for (int t = threadIdx.x; t < n; t += blockDim.x) d[t] = s[t];
}

namespace {

struct StageTimer {
    std::map<std::string, double> ms;
    void add(const char *k, double v) { ms[k] += v; }
};

// Optional wall-clock timeline of every timed phase of every worker (VM_TIMELINE=<file>): one line per phase,
// "worker name start_ms end_ms", for looking at how host glue and kernels of the pipelined workers overlap.
struct Timeline {
    std::mutex mu;
    std::vector<std::string> lines;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    const char *path = getenv("VM_TIMELINE");
    static Timeline &get() { static Timeline t; return t; }
    void add(const void *who, const char *name, std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b)
    {
        if (!path) return;
        char buf[160];
        snprintf(buf, sizeof(buf), "%p %s %.3f %.3f", who, name, std::chrono::duration<double, std::milli>(a - t0).count(),
                 std::chrono::duration<double, std::milli>(b - t0).count());
        std::lock_guard<std::mutex> lk(mu);
        lines.push_back(buf);
    }
    void flush()
    {
        if (!path) return;
        std::lock_guard<std::mutex> lk(mu);
        FILE *f = fopen(path, "a");
        if (!f) return;
        for (const std::string &l : lines) fprintf(f, "%s\n", l.c_str());
        fprintf(f, "# flush\n");
        fclose(f);
        lines.clear();
    }
};

class CudaBackend : public Backend {
public:
    CudaBackend(vm_ctx *c, vm_index_handle *ih) : c_(c), ih_(ih)
    {
        device_extension = getenv("VM_HOST_GLUE") == nullptr;      // A/B switch: the vector-based host glue of round 1
        device_front = device_extension && getenv("VM_HOST_FRONT") == nullptr;
    }
    ~CudaBackend() override
    {
        for (int i = 0; i < 3; ++i) {
            if (side_[i]) cudaStreamDestroy(side_[i]);
            if (side_done_[i]) cudaEventDestroy(side_done_[i]);
        }
        if (side_go_) cudaEventDestroy(side_go_);
        seed_.release();
        gx_.release();
        lx_.release();
        { VmDevBuf *xb[] = {&x_ids_, &x_used_, &x_tmp_anc_, &x_tmp_S_, &x_tmp_len_, &x_tmp_score_, &d_ctg_}; for (VmDevBuf *x : xb) x->release(); }
        VmDevBuf *b[] = {&reads_fwd_, &reads_rc_, &read_off_, &jobs_, &d_wlo_, &d_whi_, &d_gx_, &d_gy_, &d_nh_, &d_hits_, &d_tab_,
                         &d_order_, &d_rout_, &d_dense_, &d_seg_, &d_dir_, &d_sc_, &d_cig_, &d_cigd_, &d_pairs_, &d_msegs_};
        for (VmDevBuf *x : b) x->release();
        { VmDevBuf *xb[] = {&dx_nrev_, &d_mask_, &d_fstats_, &d_cigd2_}; for (VmDevBuf *x : xb) x->release(); }
        bplanbufs_.release();
        fplanbufs_.release();
        back_.reset();
        front_.reset();
        VmPinnedBuf *p[] = {&h_sorted_, &h_S_, &h_P_, &h_A_, &h_gmax_, &h_jobs_, &h_cig_, &h_lsorted_, &h_lP_, &h_lgmax_, &h_misc_, &h_gx_, &h_gy_,
                            &h_segs_};
        for (VmPinnedBuf *x : p) x->release();
    }
    StageTimer timer;
    bool reads_resident = false;    // vm_reads_upload already put this batch in HBM
    bool device_extension = true;   // extend_func's glue runs on the device (extend_device); false: the host glue of vm_glue.hpp
    bool device_front = true;       // hit2work_1's bookkeeping, guide selection and re-seeding jobs on the device (front_device)
    int host_threads = 1;           // host threads this backend may use for staging loops
    void set_index(vm_index_handle *ih) { ih_ = ih; }
    double fill_cells_ = 0, fill_bases_ = 0, fill_jobs_ = 0, ed_cells_ = 0, reseed_hits_ = 0, chain_anchors_ = 0, ed_upper_jobs_ = 0, fill_band_jobs_ = 0,
           fill_band_redo_ = 0, fill_dir_bytes_ = 0, chain_opcount_ = 0;
    void reset_counters()
    {
        timer.ms.clear();
        fill_cells_ = fill_bases_ = fill_jobs_ = ed_cells_ = reseed_hits_ = chain_anchors_ = ed_upper_jobs_ = fill_band_jobs_ = fill_band_redo_ = fill_dir_bytes_ = chain_opcount_ = 0;
    }

    // device time of a group of launches, CUDA events on the ctx stream
    struct KTimer {
        CudaBackend *be; const char *name; cudaEvent_t a, b; std::chrono::steady_clock::time_point w0;
        KTimer(CudaBackend *be_, const char *n) : be(be_), name(n), w0(std::chrono::steady_clock::now())
        {
            cudaEventCreateWithFlags(&a, cudaEventBlockingSync); cudaEventCreateWithFlags(&b, cudaEventBlockingSync);
            cudaEventRecord(a, be->c_->stream);
        }
        void stop()
        {
            cudaEventRecord(b, be->c_->stream);
            cudaEventSynchronize(b);
            float ms = 0;
            cudaEventElapsedTime(&ms, a, b);
            be->timer.add(name, ms);
            cudaEventDestroy(a); cudaEventDestroy(b);
            Timeline::get().add(be, name, w0, std::chrono::steady_clock::now());
        }
    };
    static double process_cpu_ms()
    {
        timespec ts;
        clock_gettime(CLOCK_PROCESS_CPUTIME_ID, &ts);
        return 1e3 * (double)ts.tv_sec + 1e-6 * (double)ts.tv_nsec;
    }
    // wall time of a stage, plus ("cpu_<name>") the CPU time the whole process burnt meanwhile -- in a lock-step
    // run (one worker) that is the stage's own host cost on all threads
    struct WallTimer {
        CudaBackend *be; const char *name; std::chrono::steady_clock::time_point t0; double c0;
        WallTimer(CudaBackend *b, const char *n) : be(b), name(n), t0(std::chrono::steady_clock::now()), c0(process_cpu_ms()) {}
        ~WallTimer()
        {
            const auto t1 = std::chrono::steady_clock::now();
            be->timer.add(name, std::chrono::duration<double, std::milli>(t1 - t0).count());
            be->timer.add((std::string("cpu_") + name).c_str(), process_cpu_ms() - c0);
            Timeline::get().add(be, name, t0, t1);
        }
    };

    // b.off may start anywhere inside b.seq (a sub-batch of a larger batch): the device copy is rebased to 0
    void upload_reads(const ReadBatch &b)
    {
        const int64_t base = b.off[0];
        if (reads_resident && off_host_.size() == (size_t)b.n + 1) {
            bool same = true;
            for (int64_t i = 0; i <= b.n && same; ++i) same = off_host_[i] == b.off[i] - base;
            if (same) return;
        }
        const size_t total = (size_t)(b.off[b.n] - base);
        off_host_.resize((size_t)b.n + 1);
        for (int64_t i = 0; i <= b.n; ++i) off_host_[i] = b.off[i] - base;
        BE_OK(reads_fwd_.ensure(total + 64));
        BE_OK(reads_rc_.ensure(total + 64));
        BE_OK(read_off_.ensure((size_t)(b.n + 1) * 8));
        BE_OK(cudaMemcpyAsync(reads_fwd_.p, b.seq + base, total, cudaMemcpyHostToDevice, c_->stream));
        BE_OK(cudaMemcpyAsync(read_off_.p, off_host_.data(), (size_t)(b.n + 1) * 8, cudaMemcpyHostToDevice, c_->stream));
        if (total > 0) {
            vm_revcomp_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c_->stream>>>(reads_fwd_.as<uint8_t>(), read_off_.as<int64_t>(),
                                                                                      (int)b.n, (int64_t)total, reads_rc_.as<uint8_t>());
            c_->launches += 1;
        }
        BE_OK(vm_stream_sync(c_->stream));
    }

    // reads [r0, r0 + n) of a batch another backend of this device already holds in HBM (device-to-device)
    void copy_reads_from(const CudaBackend &src, int64_t r0, int64_t n)
    {
        const int64_t base = src.off_host_[(size_t)r0];
        const size_t total = (size_t)(src.off_host_[(size_t)(r0 + n)] - base);
        off_host_.resize((size_t)n + 1);
        for (int64_t i = 0; i <= n; ++i) off_host_[i] = src.off_host_[(size_t)(r0 + i)] - base;
        BE_OK(reads_fwd_.ensure(total + 64));
        BE_OK(reads_rc_.ensure(total + 64));
        BE_OK(read_off_.ensure((size_t)(n + 1) * 8));
        BE_OK(cudaMemcpyAsync(reads_fwd_.p, src.reads_fwd_.as<uint8_t>() + base, total, cudaMemcpyDeviceToDevice, c_->stream));
        BE_OK(cudaMemcpyAsync(reads_rc_.p, src.reads_rc_.as<uint8_t>() + base, total, cudaMemcpyDeviceToDevice, c_->stream));
        BE_OK(cudaMemcpyAsync(read_off_.p, off_host_.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, c_->stream));
        BE_OK(vm_stream_sync(c_->stream));
        reads_resident = true;
    }

    // stage-level: seeding only, anchors copied to the host (parity tests)
    void seed_only(const ReadBatch &b, int check_num, std::vector<VmAnchor> &flat, std::vector<int64_t> &a_off,
                   std::vector<int32_t> &n_out, std::vector<int32_t> &nrev)
    {
        upload_reads(b);
        std::string err;
        if (vm_seed_batch(seed_, ih_->ix->dev, reads_fwd_.as<uint8_t>(), read_off_.as<int64_t>(), off_host_, check_num, -1,
                          c_->stream, n_out, nrev, a_off, &c_->launches, err))
            throw std::runtime_error(err);
        flat.resize((size_t)a_off[b.n]);
        if (!flat.empty()) BE_OK(cudaMemcpy(flat.data(), seed_.out.p, flat.size() * sizeof(VmAnchor), cudaMemcpyDeviceToHost));
    }

    // ---- seeding + global chaining, anchors never leave the device in between ----
    void seed_chain(const ReadBatch &b, int check_num, int kmersize, double skipcost, int maxdiff, int maxgap, double accept,
                    std::vector<char> &need_reverse, ChainOut &out) override
    {
        WallTimer wt(this, "seed_chain");
        if (const char *inj = getenv("VM_TEST_FAIL_LEN")) {      // test hook: a read of this length cannot be processed
            const int64_t bad = atoll(inj);
            for (int64_t r = 0; r < b.n; ++r)
                if (b.len(r) == bad) throw std::runtime_error("injected failure (VM_TEST_FAIL_LEN)");
        }
        upload_reads(b);
        std::vector<int32_t> n_out, nrev;
        std::vector<int64_t> a_off;
        std::string err;
        {
            KTimer kt(this, "k_seed");
            if (vm_seed_batch(seed_, ih_->ix->dev, reads_fwd_.as<uint8_t>(), read_off_.as<int64_t>(), off_host_, check_num, -1,
                              c_->stream, n_out, nrev, a_off, &c_->launches, err))
                throw std::runtime_error(err);
            kt.stop();
        }
        const int64_t n = b.n;
        need_reverse.assign((size_t)n, 0);
        for (int64_t r = 0; r < n; ++r) need_reverse[r] = (char)nrev[r];
        out = ChainOut();
        out.start.assign(a_off.begin(), a_off.begin() + n);
        out.cnt = n_out;
        out.gmax.assign((size_t)n, -1);
        const int64_t span = a_off[n];
        chain_anchors_ += (double)span;
        std::vector<int32_t> rl((size_t)n);
        std::vector<int> ids;
        for (int64_t r = 0; r < n; ++r) {
            rl[r] = (int32_t)std::min<int64_t>(b.len(r), INT32_MAX - 128);
            if (n_out[r] > 2) ids.push_back((int)r);      // decode_hit :23986 -- <= 2 anchors: unmapped, not chained
        }
        vm_chain_params prm{kmersize, skipcost, maxdiff, maxgap, 1000, 5, 30, 0};
        if (vm_chain_prepare(c_, n, span, out.start, out.cnt, false) != VM_OK) throw std::runtime_error("chain: " + c_->err);
        float ms4[4] = {0, 0, 0, 0};
        out.used_fast.assign((size_t)n, 0);
        if (vm_chain_core(c_, prm, seed_.out.as<VmAnchor>(), out.start, out.cnt, rl, rl, ids, nullptr, &out.used_fast, ms4) != VM_OK)
            throw std::runtime_error("chain: " + c_->err);
        timer.add("chain_global_kernels", ms4[1] + ms4[2] + ms4[3]);
        chain_opcount_ += c_->chain.opcount_last;
        // chains extracted on the device: only what hit2work_1 keeps goes to the host
        extract(true, n, span, ids, accept, gx_, out);
    }

    // result buffers of one extraction (global / local stage), reused across batches
    struct Extracted {
        VmDevBuf rec, anc, S, len, score, counters;
        VmPinnedBuf h_rec, h_anc, h_S, h_len, h_score, h_counters;
        // local stage: rebuild_chain_break on the device
        VmDevBuf al_rec, al_anc, al_len;
        VmPinnedBuf h_al_rec, h_al_anc, h_al_len;
        size_t al_n_anc = 0, al_n_al = 0;
        size_t n_anc_total = 0, n_chain_total = 0;      // global stage: extracted anchors / chains of the chunk
        void release()
        {
            VmDevBuf *d[] = {&rec, &anc, &S, &len, &score, &counters, &al_rec, &al_anc, &al_len};
            for (VmDevBuf *x : d) x->release();
            VmPinnedBuf *h[] = {&h_rec, &h_anc, &h_S, &h_len, &h_score, &h_counters, &h_al_rec, &h_al_anc, &h_al_len};
            for (VmPinnedBuf *x : h) x->release();
        }
    };

    void extract(bool global, int64_t n, int64_t span, const std::vector<int> &ids, double accept, Extracted &X, ChainOut &out)
    {
        static_assert(sizeof(ExtractRec) == sizeof(VmExtractRec), "extract record layout");
        VmChainState &s = c_->chain;
        const size_t T = (size_t)std::max<int64_t>(span, 1);
        BE_OK(X.rec.ensure((size_t)(n + 1) * sizeof(VmExtractRec)));
        BE_OK(X.anc.ensure(T * 16));
        BE_OK(X.counters.ensure(64));
        BE_OK(x_tmp_anc_.ensure(T * 16));
        BE_OK(x_ids_.ensure(ids.size() * 4 + 64));
        if (!global) BE_OK(X.score.ensure((size_t)(n + 1) * 8));
        if (global) {
            BE_OK(X.S.ensure(T * 8));
            BE_OK(X.len.ensure(T * 4));
            BE_OK(X.score.ensure(T * 8));
            BE_OK(x_used_.ensure(T));
            BE_OK(x_tmp_S_.ensure(T * 8));
            BE_OK(x_tmp_len_.ensure(T * 4));
            BE_OK(x_tmp_score_.ensure(T * 8));
            BE_OK(cudaMemsetAsync(x_used_.p, 0, T, c_->stream));
        }
        BE_OK(cudaMemsetAsync(X.counters.p, 0, 64, c_->stream));
        BE_OK(cudaMemsetAsync(X.rec.p, 0, (size_t)(n + 1) * sizeof(VmExtractRec), c_->stream));
        if (!ids.empty()) BE_OK(cudaMemcpyAsync(x_ids_.p, ids.data(), ids.size() * 4, cudaMemcpyHostToDevice, c_->stream));
        VmExtractOut O;
        O.rec = X.rec.as<VmExtractRec>();
        O.anc = X.anc.as<VmAnchor>();
        O.S = X.S.as<double>();
        O.chain_len = X.len.as<int32_t>();
        O.chain_score = X.score.as<double>();
        O.n_anc_total = X.counters.as<unsigned long long>();
        O.n_chain_total = X.counters.as<unsigned long long>() + 1;
        {
            KTimer kt(this, global ? "k_extract_global" : "k_extract_local");
            if (global)
                c_->launches += vm_launch_extract_global(x_ids_.as<int>(), (int)ids.size(), s.off_dev.as<int64_t>(), s.cnt_dev.as<int32_t>(),
                                                         s.sorted.as<VmAnchor>(), s.S.as<double>(), s.P.as<int32_t>(),
                                                         s.S_arg.as<int32_t>(), s.gmax.as<int64_t>(), accept, x_used_.as<uint8_t>(),
                                                         x_tmp_anc_.as<VmAnchor>(), x_tmp_S_.as<double>(), x_tmp_len_.as<int32_t>(),
                                                         x_tmp_score_.as<double>(), O, c_->stream);
            else
                c_->launches += vm_launch_extract_local(x_ids_.as<int>(), (int)ids.size(), s.off_dev.as<int64_t>(), s.cnt_dev.as<int32_t>(),
                                                        s.sorted.as<VmAnchor>(), s.S.as<double>(), s.P.as<int32_t>(), s.gmax.as<int64_t>(),
                                                        x_tmp_anc_.as<VmAnchor>(), O, c_->stream);
            kt.stop();
        }
        const bool rebuild = !global && rebuild_large_cost_ >= 0 && ih_ != nullptr;
        if (rebuild) {
            static_assert(sizeof(RebuildRec) == sizeof(VmRebuildRec), "rebuild record layout");
            BE_OK(X.al_rec.ensure((size_t)(n + 1) * sizeof(VmRebuildRec)));
            BE_OK(X.al_anc.ensure(T * 16));
            BE_OK(X.al_len.ensure(T * 4));
            BE_OK(x_tmp_len_.ensure(T * 4));
            BE_OK(x_tmp_S_.ensure(T * 16));          // second anchor scratch (the extract kernel's is still being read)
            BE_OK(cudaMemsetAsync(X.al_rec.p, 0, (size_t)(n + 1) * sizeof(VmRebuildRec), c_->stream));
            upload_contig_starts();
            VmRebuildOut R;
            R.rec = X.al_rec.as<VmRebuildRec>();
            R.anc = X.al_anc.as<VmAnchor>();
            R.len = X.al_len.as<int32_t>();
            R.n_anc_total = X.counters.as<unsigned long long>() + 2;
            R.n_len_total = X.counters.as<unsigned long long>() + 3;
            KTimer kt(this, "k_rebuild");
            c_->launches += vm_launch_rebuild(x_ids_.as<int>(), (int)ids.size(), s.off_dev.as<int64_t>(), X.rec.as<VmExtractRec>(),
                                              X.anc.as<VmAnchor>(), d_ctg_.as<int64_t>(), n_ctg_dev_, rebuild_large_cost_, 50,
                                              x_tmp_S_.as<VmAnchor>(), x_tmp_len_.as<int32_t>(), R, c_->stream);
            kt.stop();
        }
        BE_OK(X.h_counters.ensure(64));
        BE_OK(X.h_rec.ensure((size_t)(n + 1) * sizeof(VmExtractRec)));
        BE_OK(cudaMemcpyAsync(X.h_counters.p, X.counters.p, 32, cudaMemcpyDeviceToHost, c_->stream));
        if (global && device_front) {
            // hit2work_1's bookkeeping runs on the device (front_device): only the totals cross PCIe
            BE_OK(vm_stream_sync(c_->stream));
            BE_OK(cudaGetLastError());
            X.n_anc_total = (size_t)X.h_counters.as<unsigned long long>()[0];
            X.n_chain_total = (size_t)X.h_counters.as<unsigned long long>()[1];
            return;
        }
        if (!global && device_extension && rebuild) {
            // the extension stage runs on the device (extend_device): only the totals cross PCIe
            BE_OK(vm_stream_sync(c_->stream));
            BE_OK(cudaGetLastError());
            X.al_n_anc = (size_t)X.h_counters.as<unsigned long long>()[2];
            X.al_n_al = (size_t)X.h_counters.as<unsigned long long>()[3];
            return;
        }
        BE_OK(cudaMemcpyAsync(X.h_rec.p, X.rec.p, (size_t)n * sizeof(VmExtractRec), cudaMemcpyDeviceToHost, c_->stream));
        BE_OK(vm_stream_sync(c_->stream));
        BE_OK(cudaGetLastError());
        const size_t na = (size_t)X.h_counters.as<unsigned long long>()[0], nc = (size_t)X.h_counters.as<unsigned long long>()[1];
        BE_OK(X.h_anc.ensure(std::max<size_t>(na, 1) * 16));
        if (na) BE_OK(cudaMemcpyAsync(X.h_anc.p, X.anc.p, na * 16, cudaMemcpyDeviceToHost, c_->stream));
        if (global) {
            BE_OK(X.h_S.ensure(std::max<size_t>(na, 1) * 8));
            BE_OK(X.h_len.ensure(std::max<size_t>(nc, 1) * 4));
            BE_OK(X.h_score.ensure(std::max<size_t>(nc, 1) * 8));
            if (na) BE_OK(cudaMemcpyAsync(X.h_S.p, X.S.p, na * 8, cudaMemcpyDeviceToHost, c_->stream));
            if (nc) BE_OK(cudaMemcpyAsync(X.h_len.p, X.len.p, nc * 4, cudaMemcpyDeviceToHost, c_->stream));
            if (nc) BE_OK(cudaMemcpyAsync(X.h_score.p, X.score.p, nc * 8, cudaMemcpyDeviceToHost, c_->stream));
        }
        if (!global && want_local_score_) {
            BE_OK(X.h_score.ensure((size_t)(n + 1) * 8));
            BE_OK(cudaMemcpyAsync(X.h_score.p, X.score.p, (size_t)n * 8, cudaMemcpyDeviceToHost, c_->stream));
        }
        if (rebuild) {
            const size_t ra = (size_t)X.h_counters.as<unsigned long long>()[2], rl = (size_t)X.h_counters.as<unsigned long long>()[3];
            X.al_n_anc = ra;
            X.al_n_al = rl;
            BE_OK(X.h_al_rec.ensure((size_t)(n + 1) * sizeof(VmRebuildRec)));
            BE_OK(X.h_al_anc.ensure(std::max<size_t>(ra, 1) * 16));
            BE_OK(X.h_al_len.ensure(std::max<size_t>(rl, 1) * 4));
            BE_OK(cudaMemcpyAsync(X.h_al_rec.p, X.al_rec.p, (size_t)n * sizeof(VmRebuildRec), cudaMemcpyDeviceToHost, c_->stream));
            if (ra) BE_OK(cudaMemcpyAsync(X.h_al_anc.p, X.al_anc.p, ra * 16, cudaMemcpyDeviceToHost, c_->stream));
            if (rl) BE_OK(cudaMemcpyAsync(X.h_al_len.p, X.al_len.p, rl * 4, cudaMemcpyDeviceToHost, c_->stream));
        }
        BE_OK(vm_stream_sync(c_->stream));
        if (rebuild) {
            out.al_rec = X.h_al_rec.as<RebuildRec>();
            out.al_anc = X.h_al_anc.as<Anc32>();
            out.al_len = X.h_al_len.as<int32_t>();
        }
        out.rec = X.h_rec.as<ExtractRec>();
        out.x_anc = X.h_anc.as<Anc32>();
        out.x_S = X.h_S.as<double>();
        out.x_len = X.h_len.as<int32_t>();
        out.x_score = X.h_score.as<double>();
    }

    // local 9-mer re-seeding of every guide job (hits + same-diagonal merge): afterwards job j's anchors are
    // d_rout_[2 * J[j].dense_off .. + n_out[j]) on the device
    void reseed_device(const std::vector<char> &need_reverse, const std::vector<GuideJobRef> &jobs, std::vector<VmReseedJobDev> &J,
                       std::vector<int32_t> &n_out, int32_t *&d_n_out)
    {
        const int nj = (int)jobs.size();
        WallTimer *hs = new WallTimer(this, "h_reseed_stage");
        const bool from_device = false;
        J.assign((size_t)nj, VmReseedJobDev());
        std::vector<int64_t> w_off((size_t)nj + 1, 0), g_off((size_t)nj + 1, 0);
        for (int j = 0; j < nj; ++j) {
            w_off[j + 1] = w_off[j] + (int64_t)jobs[j].job.win_lo.size();
            g_off[j + 1] = g_off[j] + (int64_t)jobs[j].job.gx.size();
        }
        std::vector<int64_t> wlo((size_t)w_off[nj]), whi((size_t)w_off[nj]);
        // the guide points are the bulk of this stage's upload: staged in page-locked memory
        const size_t n_g = (size_t)g_off[nj];
        BE_OK(h_gy_.ensure(n_g * 8 + 64));
        BE_OK(h_gx_.ensure(n_g * 4 + 64));
        int64_t *gy = h_gy_.as<int64_t>();
        int32_t *gx = h_gx_.as<int32_t>();
        parallel_for(nj, host_threads, [&](int64_t j) {
            const vmg::GuideJob &g = jobs[j].job;
            VmReseedJobDev &d = J[j];
            memset(&d, 0, sizeof(d));
            d.read = jobs[j].read;
            d.need_reverse = need_reverse[jobs[j].read] ? 1 : 0;
            d.readstart = g.readstart;
            d.readend = g.readend;
            d.n_win = (int32_t)g.win_lo.size();
            d.n_guide = (int32_t)g.gx.size();
            d.win_off = w_off[j];
            d.g_off = g_off[j];
            std::copy(g.win_lo.begin(), g.win_lo.end(), wlo.begin() + w_off[j]);
            std::copy(g.win_hi.begin(), g.win_hi.end(), whi.begin() + w_off[j]);
            std::copy(g.gx.begin(), g.gx.end(), gx + g_off[j]);
            std::copy(g.gy.begin(), g.gy.end(), gy + g_off[j]);
        }, 64);
        BE_OK(d_wlo_.ensure(wlo.size() * 8 + 64));
        BE_OK(d_whi_.ensure(whi.size() * 8 + 64));
        BE_OK(d_gx_.ensure(n_g * 4 + 64));
        BE_OK(d_gy_.ensure(n_g * 8 + 64));
        BE_OK(cudaMemcpyAsync(d_wlo_.p, wlo.data(), wlo.size() * 8, cudaMemcpyHostToDevice, c_->stream));
        BE_OK(cudaMemcpyAsync(d_whi_.p, whi.data(), whi.size() * 8, cudaMemcpyHostToDevice, c_->stream));
        BE_OK(cudaMemcpyAsync(d_gx_.p, gx, n_g * 4, cudaMemcpyHostToDevice, c_->stream));
        BE_OK(cudaMemcpyAsync(d_gy_.p, gy, n_g * 8, cudaMemcpyHostToDevice, c_->stream));
        reseed_run(J, n_out, d_n_out, hs, from_device);
    }

    // The re-seeding kernels over the job slots J (host copy; windows and guide points already in d_wlo_ / d_whi_ / d_gx_ /
    // d_gy_ on the device).  A slot with n_guide == 0 is empty.  hs: the staging timer to close before the kernels.
    void reseed_run(std::vector<VmReseedJobDev> &J, std::vector<int32_t> &n_out, int32_t *&d_n_out, WallTimer *hs, bool from_device)
    {
        (void)from_device;
        const int nj = (int)J.size();
        // one pass with room for 1.5 hits per read position (measured: 0.2 on 15 kb ONT reads, 0.6 on 20 kb HiFi reads -- at 3
        // per position the hit buffers of a GRCh38-sized HiFi chunk took 9 GB per worker); the rare job that needs more is
        // re-run below with its exact count
        int64_t hit_off = 0;
        for (int j = 0; j < nj; ++j) {
            const int64_t span = std::max<int64_t>(0, (int64_t)J[j].readend - J[j].readstart);
            J[j].hit_off = hit_off;
            J[j].hit_cap = J[j].n_guide > 0 ? (int32_t)std::min<int64_t>(span + span / 2 + 64, INT32_MAX) : 0;
            hit_off += J[j].hit_cap;
        }
        BE_OK(jobs_.ensure(J.size() * sizeof(VmReseedJobDev) + 64));
        BE_OK(d_nh_.ensure((size_t)nj * 8 + 64));
        BE_OK(d_hits_.ensure((size_t)hit_off * vm_reseed_hit_bytes() + 64));
        BE_OK(cudaMemcpyAsync(jobs_.p, J.data(), J.size() * sizeof(VmReseedJobDev), cudaMemcpyHostToDevice, c_->stream));
        int32_t *d_n_hits = d_nh_.as<int32_t>();
        d_n_out = d_nh_.as<int32_t>() + nj;
        const VmIndexDev &ix = ih_->ix->dev;
        delete hs;
        {
            KTimer kt(this, "k_reseed_hits");
            c_->launches += vm_reseed_launch(ix, jobs_.as<VmReseedJobDev>(), nj, reads_fwd_.as<uint8_t>(), reads_rc_.as<uint8_t>(),
                                             read_off_.as<int64_t>(), d_wlo_.as<int64_t>(), d_whi_.as<int64_t>(),
                                             d_gx_.as<int32_t>(), d_gy_.as<int64_t>(), d_hits_.p, d_n_hits, c_->stream);
            kt.stop();
        }
        std::vector<int32_t> n_hits((size_t)nj);
        BE_OK(vm_d2h_sync(h_small_, n_hits.data(), d_n_hits, (size_t)nj * 4, c_->stream));
        {
            // jobs that overflowed their capacity: exact room at the end of the hit buffer, second launch
            std::vector<int> redo;
            std::vector<VmReseedJobDev> RJ;
            for (int j = 0; j < nj; ++j)
                if (n_hits[j] > J[j].hit_cap) {
                    J[j].hit_off = hit_off;
                    J[j].hit_cap = n_hits[j];
                    hit_off += n_hits[j];
                    redo.push_back(j);
                    RJ.push_back(J[j]);
                }
            if (!redo.empty()) {
                // growing the buffer must keep the hits already written
                VmDevBuf bigger;
                BE_OK(bigger.ensure((size_t)hit_off * vm_reseed_hit_bytes() + 64));
                BE_OK(cudaMemcpyAsync(bigger.p, d_hits_.p, (size_t)RJ[0].hit_off * vm_reseed_hit_bytes(), cudaMemcpyDeviceToDevice,
                                      c_->stream));
                BE_OK(vm_stream_sync(c_->stream));
                d_hits_.release();
                d_hits_ = bigger;
                BE_OK(d_seg_.ensure(RJ.size() * (sizeof(VmReseedJobDev) + 4) + 64));
                BE_OK(cudaMemcpyAsync(d_seg_.p, RJ.data(), RJ.size() * sizeof(VmReseedJobDev), cudaMemcpyHostToDevice, c_->stream));
                int32_t *d_rn = (int32_t *)(d_seg_.as<VmReseedJobDev>() + RJ.size());
                KTimer kt(this, "k_reseed_hits");
                c_->launches += vm_reseed_launch(ix, d_seg_.as<VmReseedJobDev>(), (int)RJ.size(), reads_fwd_.as<uint8_t>(),
                                                 reads_rc_.as<uint8_t>(), read_off_.as<int64_t>(), d_wlo_.as<int64_t>(),
                                                 d_whi_.as<int64_t>(), d_gx_.as<int32_t>(), d_gy_.as<int64_t>(), d_hits_.p, d_rn,
                                                 c_->stream);
                kt.stop();
            }
        }
        int64_t dense_hits = 0, tab_off = 0;
        for (int j = 0; j < nj; ++j) {
            J[j].dense_off = dense_hits;
            dense_hits += n_hits[j];
            // the merge kernel gives every lane a 32nd of the table; the diagonals are a fraction of the hits, so a table of
            // the next power of two above the hit count is about a third full (a lane slice that does fill up sends the job to
            // the sequential kernel)
            int ts = J[j].n_guide > 0 ? 1024 : 32;      // an empty slot has no table
            while (ts < n_hits[j] + 8) ts <<= 1;
            J[j].tab_off = tab_off;
            J[j].tab_size = ts;
            tab_off += ts;
        }
        reseed_hits_ += (double)dense_hits;
        BE_OK(d_tab_.ensure((size_t)tab_off * vm_reseed_point_bytes() + 64));
        BE_OK(d_order_.ensure((size_t)dense_hits * 4 + 64));
        BE_OK(d_rout_.ensure((size_t)dense_hits * 2 * sizeof(VmAnchor) + 64));
        BE_OK(cudaMemcpyAsync(jobs_.p, J.data(), J.size() * sizeof(VmReseedJobDev), cudaMemcpyHostToDevice, c_->stream));
        {
            KTimer kt(this, "k_reseed_merge");
            c_->launches += vm_reseed_merge_launch(jobs_.as<VmReseedJobDev>(), nj, d_hits_.p, d_n_hits, d_tab_.p,
                                                   d_order_.as<int32_t>(), d_rout_.as<VmAnchor>(), d_n_out, c_->stream);
            kt.stop();
        }
        n_out.assign((size_t)nj, 0);
        BE_OK(vm_d2h_sync(h_small_, n_out.data(), d_n_out, (size_t)nj * 4, c_->stream));
    }

    // stage-level: the anchors of every guide job on the host (parity tests)
    void reseed_only(const ReadBatch &b, const std::vector<char> &need_reverse, const std::vector<GuideJobRef> &jobs,
                     std::vector<VmAnchor> &flat, std::vector<int64_t> &job_off)
    {
        upload_reads(b);
        std::vector<VmReseedJobDev> J;
        std::vector<int32_t> n_out;
        int32_t *d_n_out = nullptr;
        const int nj = (int)jobs.size();
        job_off.assign((size_t)nj + 1, 0);
        if (nj == 0) { flat.clear(); return; }
        reseed_device(need_reverse, jobs, J, n_out, d_n_out);
        for (int j = 0; j < nj; ++j) job_off[(size_t)j + 1] = job_off[(size_t)j] + n_out[(size_t)j];
        flat.resize((size_t)job_off[(size_t)nj]);
        for (int j = 0; j < nj; ++j)
            if (n_out[(size_t)j] > 0)
                BE_OK(cudaMemcpyAsync(flat.data() + job_off[(size_t)j], d_rout_.as<VmAnchor>() + 2 * J[(size_t)j].dense_off,
                                      (size_t)n_out[(size_t)j] * sizeof(VmAnchor), cudaMemcpyDeviceToHost, c_->stream));
        BE_OK(vm_stream_sync(c_->stream));
    }

    // ---- local re-seeding + local chaining ----
    void reseed_chain(const ReadBatch &b, const std::vector<char> &need_reverse, const std::vector<GuideJobRef> &jobs,
                      const std::vector<int> &variant, const std::vector<double> &skipcost, int maxdiff, int maxgap,
                      ChainOut &out) override
    {
        WallTimer wt(this, "reseed_chain");
        const int64_t n = b.n;
        const int nj = (int)jobs.size();
        out = ChainOut();
        out.start.assign((size_t)n, 0);
        out.cnt.assign((size_t)n, 0);
        out.gmax.assign((size_t)n, -1);
        if (nj == 0) return;
        std::vector<VmReseedJobDev> J;
        std::vector<int32_t> n_out;
        int32_t *d_n_out = nullptr;
        reseed_device(need_reverse, jobs, J, n_out, d_n_out);
        // the jobs of a read are consecutive, reads in order
        std::vector<int32_t> job_lo((size_t)n, 0), job_n((size_t)n, 0);
        {
            size_t q = 0;
            for (int64_t r = 0; r < n; ++r) {
                job_lo[(size_t)r] = (int32_t)q;
                while (q < jobs.size() && jobs[q].read == r) ++q;
                job_n[(size_t)r] = (int32_t)(q - (size_t)job_lo[(size_t)r]);
            }
        }
        reseed_chain_finish(b, J, n_out, d_n_out, job_lo, job_n, variant, skipcost, maxdiff, maxgap, out);
    }

    // concatenate the anchors of each read's jobs (slots [job_lo[r], job_lo[r] + job_n[r]) of J), chain them, trace back
    void reseed_chain_finish(const ReadBatch &b, const std::vector<VmReseedJobDev> &J, const std::vector<int32_t> &n_out, int32_t *d_n_out,
                             const std::vector<int32_t> &job_lo, const std::vector<int32_t> &job_n, const std::vector<int> &variant,
                             const std::vector<double> &skipcost, int maxdiff, int maxgap, ChainOut &out)
    {
        const int64_t n = b.n;
        const int nj = (int)J.size();
        // one dense anchor list per read on the device
        std::vector<int64_t> seg(2 * (size_t)nj, 0);   // [src_off | dst_off]
        int64_t dense = 0;
        for (int64_t r = 0; r < n; ++r) {
            out.start[r] = dense;
            for (int32_t q = job_lo[(size_t)r]; q < job_lo[(size_t)r] + job_n[(size_t)r]; ++q) {
                seg[(size_t)q] = 2 * J[(size_t)q].dense_off;
                seg[(size_t)nj + (size_t)q] = dense;
                dense += n_out[(size_t)q];
            }
            out.cnt[r] = variant[r] ? (int32_t)(dense - out.start[r]) : 0;
        }
        BE_OK(d_dense_.ensure((size_t)std::max<int64_t>(dense, 1) * sizeof(VmAnchor)));
        BE_OK(d_seg_.ensure(seg.size() * 8 + 64));
        BE_OK(cudaMemcpyAsync(d_seg_.p, seg.data(), seg.size() * 8, cudaMemcpyHostToDevice, c_->stream));
        vm_gather_segments_kernel<VmAnchor><<<nj, 128, 0, c_->stream>>>(d_rout_.as<VmAnchor>(), d_seg_.as<int64_t>(),
                                                                        d_seg_.as<int64_t>() + nj, d_n_out, d_dense_.as<VmAnchor>());
        c_->launches += 1;
        chain_anchors_ += (double)dense;
        // local DP per (variant, skipcost) group
        if (vm_chain_prepare(c_, n, dense, out.start, out.cnt, false) != VM_OK) throw std::runtime_error("chain: " + c_->err);
        std::map<std::pair<int, double>, std::vector<int>> groups;
        std::vector<int32_t> rl((size_t)n);
        for (int64_t r = 0; r < n; ++r) {
            rl[r] = (int32_t)std::min<int64_t>(b.len(r) + 64, INT32_MAX - 128);
            if (variant[r] != 0 && out.cnt[r] > 0) groups[{variant[r], skipcost[r]}].push_back((int)r);
        }
        out.used_fast.assign((size_t)n, 0);
        for (auto &g : groups) {
            vm_chain_params prm{9, g.first.second, maxdiff, maxgap, 1000, 5, 30, g.first.first};
            float ms4[4] = {0, 0, 0, 0};
            if (vm_chain_core(c_, prm, d_dense_.as<VmAnchor>(), out.start, out.cnt, rl, rl, g.second, nullptr, &out.used_fast, ms4) != VM_OK)
                throw std::runtime_error("chain: " + c_->err);
            timer.add("chain_local_kernels", ms4[1] + ms4[2] + ms4[3]);
            chain_opcount_ += c_->chain.opcount_last;
        }
        // best chain of every read traced back (and trimmed) on the device
        std::vector<int> xids;
        for (int64_t r = 0; r < n; ++r)
            if (variant[r] != 0 && out.cnt[r] > 0) xids.push_back((int)r);
        rebuild_large_cost_ = maxdiff;        // rebuild_chain_break's large_cost is the local maxdiff (:19243)
        extract(false, n, dense, xids, 0.0, lx_, out);
        rebuild_large_cost_ = -1;
    }

    // contig start offsets on the device (pos2contig for the rebuild kernel), refreshed when the index changes
    void upload_contig_starts()
    {
        upload_contig_table();
        return;
        if (d_ctg_for_ == ih_) return;
        const std::vector<int64_t> &st = ih_->ctg.start;
        BE_OK(d_ctg_.ensure(st.size() * 8 + 64));
        BE_OK(cudaMemcpyAsync(d_ctg_.p, st.data(), st.size() * 8, cudaMemcpyHostToDevice, c_->stream));
        BE_OK(vm_stream_sync(c_->stream));
        n_ctg_dev_ = (int)st.size();
        d_ctg_for_ = ih_;
    }

    // side streams of this backend (created on first use), ordered against the main stream with events
    static const int kSide = 3;
    void side_fork()
    {
        if (!side_[0]) {
            for (int i = 0; i < kSide; ++i) {
                BE_OK(cudaStreamCreateWithFlags(&side_[i], cudaStreamNonBlocking));
                BE_OK(cudaEventCreateWithFlags(&side_done_[i], cudaEventDisableTiming));
            }
            BE_OK(cudaEventCreateWithFlags(&side_go_, cudaEventDisableTiming));
        }
        BE_OK(cudaEventRecord(side_go_, c_->stream));
        for (int i = 0; i < kSide; ++i) BE_OK(cudaStreamWaitEvent(side_[i], side_go_, 0));
    }
    void side_join()
    {
        for (int i = 0; i < kSide; ++i) {
            BE_OK(cudaEventRecord(side_done_[i], side_[i]));
            BE_OK(cudaStreamWaitEvent(c_->stream, side_done_[i], 0));
        }
    }

    // Stage-level local chaining (parity tests): anchors as int64 rows, already in the DP's order when
    // `presorted` (the reference functions take sorted input); the exact DP with its own fall-back, or the _fast
    // variant outright.  Per read: the best chain's score and its trimmed path in ASCENDING read order.
    void chain_local_stage(const vm_chain_params &prm, bool presorted, bool force_fast, int64_t n, const int64_t *rows,
                           const int64_t *off, const int32_t *read_len, ChainOut &out, std::vector<double> &score,
                           std::vector<int32_t> &used_fast)
    {
        const int64_t total = off[n];
        std::vector<VmAnchor> h((size_t)std::max<int64_t>(total, 1));
        for (int64_t t = 0; t < total; ++t) {
            h[(size_t)t].x = (int32_t)rows[4 * t];
            h[(size_t)t].y = (uint32_t)rows[4 * t + 1];
            h[(size_t)t].s = (int32_t)rows[4 * t + 2];
            h[(size_t)t].l = (int32_t)rows[4 * t + 3];
        }
        BE_OK(d_dense_.ensure(h.size() * sizeof(VmAnchor)));
        BE_OK(cudaMemcpyAsync(d_dense_.p, h.data(), (size_t)total * sizeof(VmAnchor), cudaMemcpyHostToDevice, c_->stream));
        out = ChainOut();
        out.start.assign(off, off + n);
        out.cnt.resize((size_t)n);
        out.gmax.assign((size_t)n, -1);
        std::vector<int32_t> rl((size_t)n);
        std::vector<int> ids;
        for (int64_t r = 0; r < n; ++r) {
            out.cnt[(size_t)r] = (int32_t)(off[r + 1] - off[r]);
            rl[(size_t)r] = read_len[r] + 64;
            if (out.cnt[(size_t)r] > 0) ids.push_back((int)r);
        }
        if (vm_chain_prepare(c_, n, total, out.start, out.cnt, false) != VM_OK) throw std::runtime_error("chain: " + c_->err);
        used_fast.assign((size_t)n, 0);
        float ms4[4] = {0, 0, 0, 0};
        if (vm_chain_core(c_, prm, d_dense_.as<VmAnchor>(), out.start, out.cnt, rl, rl, ids, nullptr, &used_fast, ms4, presorted,
                          force_fast) != VM_OK)
            throw std::runtime_error("chain: " + c_->err);
        want_local_score_ = true;
        extract(false, n, total, ids, 0.0, lx_, out);
        want_local_score_ = false;
        score.assign(lx_.h_score.as<double>(), lx_.h_score.as<double>() + n);
    }

    static VmSeqSpec spec(const vmg::SeqRef &s)
    {
        VmSeqSpec d;
        d.lo = s.lo;
        d.len = (int32_t)(s.hi - s.lo);
        d.src = s.src;
        d.reverse = s.reverse;
        d.comp = s.comp;
        return d;
    }

    VmSeqSources sources() const
    {
        VmSeqSources S;
        S.ref = ih_ ? ih_->ix->dev.ref : nullptr;
        S.reads_fwd = reads_fwd_.as<uint8_t>();
        S.reads_rc = reads_rc_.as<uint8_t>();
        S.read_off = read_off_.as<int64_t>();
        return S;
    }

    VmAlnJobDev *stage_jobs(size_t nj)
    {
        BE_OK(h_jobs_.ensure(nj * sizeof(VmAlnJobDev) + 64));
        BE_OK(jobs_.ensure(nj * sizeof(VmAlnJobDev) + 64));
        return h_jobs_.as<VmAlnJobDev>();
    }

    // Divergence filter distances.  Jobs that come with their chain's match segments are first bounded from
    // above by the alignment through those segments (vm_ed_upper_kernel); a bound within the job's band settles
    // the filter, so only the jobs it leaves open (none on typical reads) run the exact banded kernel.
    vmg::MatchSeg *seg_staging(size_t n_segs) override
    {
        BE_OK(h_segs_.ensure(n_segs * sizeof(vmg::MatchSeg) + 64));
        return h_segs_.as<vmg::MatchSeg>();
    }

    void edit_distance(const ReadBatch &, std::vector<EdJob> &jobs, const vmg::MatchSeg *segs, size_t n_segs) override
    {
        WallTimer wt(this, "edit_distance");
        const int nj = (int)jobs.size();
        if (nj == 0) return;
        const bool check = getenv("VM_ED_CHECK") != nullptr;   // debug: also run the exact kernel and compare
        std::vector<int> open_jobs;
        std::vector<int> ub;
        bool on_device = false;
        for (int j = 0; j < nj; ++j) {
            if (jobs[j].seg_n > 0 && jobs[j].band >= 0) { ub.push_back(j); on_device = on_device || jobs[j].segs_on_device; }
            else open_jobs.push_back(j);
        }
        std::vector<int64_t> bound;
        if (!ub.empty()) {
            const int nu = (int)ub.size();
            VmAlnJobDev *J = stage_jobs((size_t)nu);
            parallel_for(nu, host_threads, [&](int64_t t) {
                const EdJob &e = jobs[ub[t]];
                memset(&J[t], 0, sizeof(VmAlnJobDev));
                J[t].q = spec(e.a);
                J[t].t = spec(e.b);
                J[t].read = e.read;
                J[t].dir_off = e.seg_off;
                J[t].n_out = e.seg_n;
            }, 1024);
            static_assert(sizeof(vmg::MatchSeg) == 12, "match segment layout");
            BE_OK(d_msegs_.ensure(std::max(n_segs, on_device ? lx_.al_n_anc : (size_t)0) * sizeof(vmg::MatchSeg) + 64));
            BE_OK(d_seg_.ensure((size_t)nu * 4 + 64));
            std::vector<int> ids((size_t)nu);
            for (int t = 0; t < nu; ++t) ids[t] = t;
            if (!on_device && segs != h_segs_.as<vmg::MatchSeg>()) {      // a caller that did not use seg_staging()
                BE_OK(h_segs_.ensure(n_segs * sizeof(vmg::MatchSeg) + 64));
                memcpy(h_segs_.p, segs, n_segs * sizeof(vmg::MatchSeg));
            }
            if (!on_device)
                BE_OK(cudaMemcpyAsync(d_msegs_.p, h_segs_.p, n_segs * sizeof(vmg::MatchSeg), cudaMemcpyHostToDevice, c_->stream));
            BE_OK(cudaMemcpyAsync(d_seg_.p, ids.data(), (size_t)nu * 4, cudaMemcpyHostToDevice, c_->stream));
            BE_OK(cudaMemcpyAsync(jobs_.p, J, (size_t)nu * sizeof(VmAlnJobDev), cudaMemcpyHostToDevice, c_->stream));
            KTimer kt(this, "k_ed_upper");
            if (on_device)      // the segments are derived from the sub-alignments' anchors the rebuild kernel left in HBM
                c_->launches += vm_launch_match_segments(jobs_.as<VmAlnJobDev>(), nu, lx_.al_anc.as<VmAnchor>(), d_msegs_.p, c_->stream);
            c_->launches += vm_launch_ed_upper(jobs_.as<VmAlnJobDev>(), d_seg_.as<int>(), nu, d_msegs_.p, sources(), c_->stream);
            kt.stop();
            BE_OK(cudaMemcpyAsync(J, jobs_.p, (size_t)nu * sizeof(VmAlnJobDev), cudaMemcpyDeviceToHost, c_->stream));
            BE_OK(vm_stream_sync(c_->stream));
            BE_OK(cudaGetLastError());
            bound.assign((size_t)nj, -1);
            for (int t = 0; t < nu; ++t) {
                const int j = ub[t];
                bound[j] = J[t].result0;
                if (J[t].result0 <= jobs[j].band && !check) jobs[j].dist = J[t].result0;
                else open_jobs.push_back(j);
            }
            ed_upper_jobs_ += nu;
        }
        if (open_jobs.empty()) return;
        ed_exact(jobs, open_jobs);
        if (check)
            for (int j : ub)
                if (bound[j] < jobs[j].dist && jobs[j].dist <= jobs[j].band)
                    throw std::runtime_error("VM_ED_CHECK: upper bound " + std::to_string(bound[j]) + " below the distance " +
                                             std::to_string(jobs[j].dist));
    }

    void ed_exact(std::vector<EdJob> &jobs, const std::vector<int> &which)
    {
        const int nj = (int)which.size();
        VmAlnJobDev *J = stage_jobs((size_t)nj);
        std::vector<std::vector<int>> cls(VM_ED_NCLASS);
        std::vector<int64_t> work((size_t)nj);
        int class_words[VM_ED_NCLASS] = {0};
        for (int j = 0; j < nj; ++j) {
            const EdJob &e = jobs[which[j]];
            memset(&J[j], 0, sizeof(VmAlnJobDev));
            // the shorter sequence is the bit-vector pattern (fewer 64-row blocks)
            const bool a_short = e.a.len() <= e.b.len();
            J[j].q = spec(a_short ? e.a : e.b);
            J[j].t = spec(a_short ? e.b : e.a);
            J[j].read = e.read;
            J[j].out_off = e.band;
            const int m = J[j].q.len, n = J[j].t.len, W = (m + 63) / 64;
            const int G = vm_ed_slots(m, n, e.band);
            int k = 0;
            while (k < VM_ED_NCLASS && VM_ED_CLASS_G[k] < G) ++k;
            if (k == VM_ED_NCLASS) throw std::runtime_error("edit distance: sequence pair too long for the register-resident band");
            cls[k].push_back(j);
            class_words[k] = std::max(class_words[k], W);
            work[j] = (int64_t)(n + W) * VM_ED_CLASS_G[k];
            const int64_t kk = e.band < 0 ? std::max(m, n) : e.band;
            ed_cells_ += (double)std::min<int64_t>(m, kk + 64) * (double)n;   // cells inside the (k + 1)-diagonal band
        }
        std::vector<int> ids;
        int class_start[VM_ED_NCLASS + 1];
        for (int k = 0; k < VM_ED_NCLASS; ++k) {
            // longest jobs first: the tail of a launch is one warp deep
            std::sort(cls[k].begin(), cls[k].end(), [&](int x, int y) { return work[x] != work[y] ? work[x] > work[y] : x < y; });
            class_start[k] = (int)ids.size();
            ids.insert(ids.end(), cls[k].begin(), cls[k].end());
        }
        class_start[VM_ED_NCLASS] = (int)ids.size();
        BE_OK(d_seg_.ensure(ids.size() * 4 + 64));
        BE_OK(cudaMemcpyAsync(d_seg_.p, ids.data(), ids.size() * 4, cudaMemcpyHostToDevice, c_->stream));
        BE_OK(cudaMemcpyAsync(jobs_.p, J, (size_t)nj * sizeof(VmAlnJobDev), cudaMemcpyHostToDevice, c_->stream));
        KTimer kt(this, "k_edit_distance");
        c_->launches += vm_launch_edit_distance(jobs_.as<VmAlnJobDev>(), d_seg_.as<int>(), class_start, class_words, sources(),
                                                c_->stream);
        kt.stop();
        BE_OK(cudaMemcpyAsync(J, jobs_.p, (size_t)nj * sizeof(VmAlnJobDev), cudaMemcpyDeviceToHost, c_->stream));
        BE_OK(vm_stream_sync(c_->stream));
        BE_OK(cudaGetLastError());
        for (int j = 0; j < nj; ++j) jobs[which[j]].dist = J[j].result0;
    }

    void extend(const ReadBatch &, std::vector<ExtJobRef> &jobs) override
    {
        WallTimer wt(this, "extend");
        const int nj = (int)jobs.size();
        if (nj == 0) return;
        VmAlnJobDev *J = stage_jobs((size_t)nj);
        for (int j = 0; j < nj; ++j) {
            memset(&J[j], 0, sizeof(VmAlnJobDev));
            J[j].t = spec(jobs[j].job.target);
            J[j].q = spec(jobs[j].job.query);
            J[j].read = jobs[j].read;
        }
        BE_OK(cudaMemcpyAsync(jobs_.p, J, (size_t)nj * sizeof(VmAlnJobDev), cudaMemcpyHostToDevice, c_->stream));
        KTimer kt(this, "k_extend");
        c_->launches += vm_launch_extend(jobs_.as<VmAlnJobDev>(), nj, sources(), c_->stream);
        kt.stop();
        BE_OK(cudaMemcpyAsync(J, jobs_.p, (size_t)nj * sizeof(VmAlnJobDev), cudaMemcpyDeviceToHost, c_->stream));
        BE_OK(vm_stream_sync(c_->stream));
        BE_OK(cudaGetLastError());
        for (int j = 0; j < nj; ++j) { jobs[j].job.q_e = (int32_t)J[j].result0; jobs[j].job.t_e = (int32_t)J[j].result1; }
    }

    const uint32_t *fill(const ReadBatch &, bool eqx, std::vector<FillJobRef> &jobs) override
    {
        WallTimer wt(this, "fill");
        const int nj = (int)jobs.size();
        if (nj == 0) return nullptr;
        VmAlnJobDev *J = stage_jobs((size_t)nj);
        int64_t out_off = 0;
        {
            WallTimer w1(this, "h_fill_stage");
            parallel_for(nj, host_threads, [&](int64_t j) {
                memset(&J[j], 0, sizeof(VmAlnJobDev));
                J[j].t = spec(jobs[j].job.target);
                J[j].q = spec(jobs[j].job.query);
                J[j].read = jobs[j].read;
            }, 1024);
            double cells = 0, bases = 0;
            for (int j = 0; j < nj; ++j) {
                J[j].out_off = out_off;
                out_off += (int64_t)J[j].t.len + J[j].q.len + 2;
                cells += (double)J[j].t.len * (double)J[j].q.len;
            }
            bases = (double)(out_off - 2LL * nj);
            fill_cells_ += cells;
            fill_bases_ += bases;
            fill_jobs_ += nj;
        }
        // Jobs the banded kernel can take (with its optimality certificate) and the rest for the full-matrix kernel
        static const bool no_band = getenv("VM_FILL_NO_BAND") != nullptr;      // debug / A-B switch
        std::vector<uint8_t> full_mask((size_t)nj, 1);
        {
            WallTimer w2(this, "h_fill_plan");
            if (!no_band) vm_fillb_plan(J, nj, c_->sm_count > 0 ? c_->sm_count : 148, host_threads, bplan_, full_mask.data());
            else { bplan_.pairs.clear(); bplan_.launches.clear(); bplan_.dir_words = 0; }
            vm_fill_plan(J, nj, c_->sm_count > 0 ? c_->sm_count : 148, plan_, host_threads, full_mask.data());
        }
        fill_dir_bytes_ += plan_.dir_bytes + bplan_.dir_bytes;
        const size_t n_launch = plan_.launches.size() + bplan_.launches.size();
        BE_OK(d_dir_.ensure((plan_.dir_words + bplan_.dir_words) * 4 + 64));
        BE_OK(d_sc_.ensure(plan_.band_words * 4 + 64));
        BE_OK(d_cig_.ensure((size_t)out_off * 4 + 64));
        BE_OK(d_cigd_.ensure((size_t)out_off * 4 + 64));
        BE_OK(d_seg_.ensure((size_t)nj * 8 + 64));
        BE_OK(d_pairs_.ensure(plan_.pairs.size() * sizeof(VmFillPair) + bplan_.pairs.size() * sizeof(VmFillBandPair) + n_launch * 4 + 128));
        // [full pairs | band pairs | dense counter (8 bytes, 8-aligned) | one launch counter each]
        VmFillBandPair *d_bpairs = (VmFillBandPair *)(d_pairs_.as<VmFillPair>() + plan_.pairs.size());
        unsigned long long *d_count = (unsigned long long *)(d_bpairs + bplan_.pairs.size());
        int *d_ctr = (int *)(d_count + 1);
        BE_OK(cudaMemcpyAsync(jobs_.p, J, (size_t)nj * sizeof(VmAlnJobDev), cudaMemcpyHostToDevice, c_->stream));
        if (!plan_.pairs.empty())
            BE_OK(cudaMemcpyAsync(d_pairs_.p, plan_.pairs.data(), plan_.pairs.size() * sizeof(VmFillPair), cudaMemcpyHostToDevice,
                                  c_->stream));
        if (!bplan_.pairs.empty())
            BE_OK(cudaMemcpyAsync(d_bpairs, bplan_.pairs.data(), bplan_.pairs.size() * sizeof(VmFillBandPair), cudaMemcpyHostToDevice,
                                  c_->stream));
        BE_OK(cudaMemsetAsync(d_count, 0, 8 + n_launch * 4 + 4, c_->stream));
        BE_OK(cudaMemsetAsync(d_seg_.p, 0, (size_t)nj * 8, c_->stream));
        BE_OK(h_misc_.ensure((size_t)nj * 8 + 64));
        unsigned long long *h_count = h_misc_.as<unsigned long long>();
        uint32_t *h_res = (uint32_t *)(h_count + 1);
        {
            KTimer kt(this, "k_fill");
            // launches that fill the device run on the worker's stream, the small ones (rare job classes, one warp's
            // latency deep) beside them on side streams: fork after the uploads, join before the read-back
            side_fork();
            int rr = 0;
            size_t dir_cur = 0, band_cur = 0;
            c_->launches += vm_fillb_launch(bplan_, jobs_.as<VmAlnJobDev>(), d_bpairs, sources(), eqx ? 1 : 0, d_dir_.as<uint32_t>(),
                                            d_ctr + plan_.launches.size(), d_cig_.as<uint32_t>(), d_cigd_.as<uint32_t>(), d_count,
                                            d_seg_.p, c_->stream, side_, kSide, &rr, &dir_cur);
            c_->launches += vm_fill_launch(plan_, jobs_.as<VmAlnJobDev>(), d_pairs_.as<VmFillPair>(), sources(), eqx ? 1 : 0,
                                           d_dir_.as<uint32_t>(), d_sc_.as<uint32_t>(), d_ctr, d_cig_.as<uint32_t>(),
                                           d_cigd_.as<uint32_t>(), d_count, d_seg_.p, c_->stream, side_, kSide, &rr, &dir_cur, &band_cur);
            side_join();
            kt.stop();
        }
        if (!bplan_.pairs.empty()) {
            // jobs whose certificate failed: once more, in the full-matrix kernel
            BE_OK(cudaMemcpyAsync(h_res, d_seg_.p, (size_t)nj * 8, cudaMemcpyDeviceToHost, c_->stream));
            BE_OK(vm_stream_sync(c_->stream));
            BE_OK(cudaGetLastError());
            int n_redo = 0, n_band = 0;
            for (int j = 0; j < nj; ++j) {
                n_band += full_mask[j] == 0;
                full_mask[j] = h_res[2 * j] == 0xffffffffu;
                n_redo += full_mask[j];
            }
            fill_band_jobs_ += (double)n_band;
            fill_band_redo_ += n_redo;
            if (n_redo) {
                vm_fill_plan(J, nj, c_->sm_count > 0 ? c_->sm_count : 148, plan_, host_threads, full_mask.data());
                BE_OK(d_dir_.ensure(plan_.dir_words * 4 + 64));
                BE_OK(d_sc_.ensure(plan_.band_words * 4 + 64));
                BE_OK(d_pairs_.ensure(plan_.pairs.size() * sizeof(VmFillPair) + plan_.launches.size() * 4 + 128));
                int *d_ctr2 = (int *)(d_pairs_.as<VmFillPair>() + plan_.pairs.size());
                BE_OK(cudaMemcpyAsync(d_pairs_.p, plan_.pairs.data(), plan_.pairs.size() * sizeof(VmFillPair), cudaMemcpyHostToDevice,
                                      c_->stream));
                BE_OK(cudaMemsetAsync(d_ctr2, 0, plan_.launches.size() * 4 + 4, c_->stream));
                BE_OK(d_nh_.ensure(64));
                BE_OK(cudaMemcpyAsync(d_nh_.p, d_count, 8, cudaMemcpyDeviceToDevice, c_->stream));   // the counter moves with us
                d_count = d_nh_.as<unsigned long long>();
                KTimer kt(this, "k_fill");
                int rr = 0;
                size_t dir_cur = 0, band_cur = 0;
                c_->launches += vm_fill_launch(plan_, jobs_.as<VmAlnJobDev>(), d_pairs_.as<VmFillPair>(), sources(), eqx ? 1 : 0,
                                               d_dir_.as<uint32_t>(), d_sc_.as<uint32_t>(), d_ctr2, d_cig_.as<uint32_t>(),
                                               d_cigd_.as<uint32_t>(), d_count, d_seg_.p, c_->stream, nullptr, 0, &rr, &dir_cur, &band_cur);
                kt.stop();
            }
        }
        WallTimer w3(this, "h_fill_post");
        // per job (offset, length) into the dense CIGAR arena the kernel filled, then one dense D2H
        BE_OK(cudaMemcpyAsync(h_count, d_count, 8, cudaMemcpyDeviceToHost, c_->stream));
        BE_OK(cudaMemcpyAsync(h_res, d_seg_.p, (size_t)nj * 8, cudaMemcpyDeviceToHost, c_->stream));
        BE_OK(vm_stream_sync(c_->stream));
        BE_OK(cudaGetLastError());
        const size_t dense = (size_t)*h_count;
        BE_OK(h_cig_.ensure(std::max<size_t>(dense, 1) * 4));
        if (dense > 0) BE_OK(cudaMemcpyAsync(h_cig_.p, d_cigd_.p, dense * 4, cudaMemcpyDeviceToHost, c_->stream));
        parallel_for(nj, host_threads, [&](int64_t j) {
            jobs[j].cig_off = h_res[2 * j];
            jobs[j].cig_len = (int32_t)h_res[2 * j + 1];
        }, 4096);
        BE_OK(vm_stream_sync(c_->stream));
        return h_cig_.as<uint32_t>();
    }

    // =========================================================================================================
    // The extension stage with its glue on the device (vm_dgrun.hpp / vm_dglue.hpp): CudaExec is the execution
    // policy that runs the per-read functors as kernels on this backend's stream and the hot loops on the kernels
    // this backend already owns (divergence-filter bounds, exact banded distance, fill with device-side planning).
    // =========================================================================================================
    struct CudaExec {
        typedef VmDevBuf Buf;
        typedef VmPinnedBuf HostBuf;
        static constexpr bool kBoundsAreUpper = true;
        static void release(Buf &b) { b.release(); }
        static void release_host(HostBuf &b) { b.release(); }
        template <typename T> T *host(HostBuf &b, size_t n)
        {
            BE_OK(b.ensure(n * sizeof(T) + 64));
            return b.as<T>();
        }
        CudaBackend *be;
        template <typename T> T *ensure(Buf &b, size_t n)
        {
            BE_OK(b.ensure(n * sizeof(T) + 64));
            return b.as<T>();
        }
        void zero(void *p, size_t bytes) { if (bytes) BE_OK(cudaMemsetAsync(p, 0, bytes, be->c_->stream)); }
        void to_host(void *dst, const void *src, size_t bytes) { if (bytes) BE_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, be->c_->stream)); }
        void to_exec(void *dst, const void *src, size_t bytes) { if (bytes) BE_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, be->c_->stream)); }
        void sync()
        {
            BE_OK(vm_stream_sync(be->c_->stream));
            BE_OK(cudaGetLastError());
        }
        template <typename F> void per_item(int64_t n, const F &f)
        {
            if (n <= 0) return;
            vm_dg_item_kernel<F><<<(unsigned)((n + 63) / 64), 64, 0, be->c_->stream>>>(n, f);
            be->c_->launches += 1;
        }
        template <typename F> void per_item_warp(int64_t n, const F &f)
        {
            if (n <= 0) return;
            KTimer kt(be, "k_extend");
            vm_dg_warp_kernel<F><<<(unsigned)((n + 3) / 4), 128, 0, be->c_->stream>>>(n, f, be->sources());
            be->c_->launches += 1;
            kt.stop();
        }
        template <typename F> void per_ids_write(const int32_t *ids, int64_t n, const F &f)
        {
            if (n <= 0) return;
            vm_dg_write_kernel<F><<<(unsigned)((n + 3) / 4), 128, 0, be->c_->stream>>>(ids, n, f);
            be->c_->launches += 1;
        }
        void ed_bounds(vmd::Job *jobs, int64_t n, const vmd::A32 *al_anc)
        {
            if (n <= 0) return;
            BE_OK(be->d_msegs_.ensure(std::max<size_t>(be->lx_.al_n_anc, 1) * sizeof(vmg::MatchSeg) + 64));
            KTimer kt(be, "k_ed_upper");
            be->c_->launches += vm_launch_match_segments((VmAlnJobDev *)jobs, (int)n, (const VmAnchor *)al_anc, be->d_msegs_.p, be->c_->stream);
            be->c_->launches += vm_launch_ed_upper((VmAlnJobDev *)jobs, nullptr, (int)n, be->d_msegs_.p, be->sources(), be->c_->stream);
            kt.stop();
            be->ed_upper_jobs_ += (double)n;
        }
        // the jobs the bound left open: exact banded distance (host-planned by register class, as the stage-level path)
        void ed_exact(vmd::Job *jobs, const int32_t *open, int64_t n_open)
        {
            std::vector<int32_t> ids((size_t)n_open);
            to_host(ids.data(), open, (size_t)n_open * 4);
            sync();
            std::vector<VmAlnJobDev> J((size_t)n_open);
            for (int64_t i = 0; i < n_open; ++i) to_host(&J[(size_t)i], (VmAlnJobDev *)jobs + ids[(size_t)i], sizeof(VmAlnJobDev));
            sync();
            std::vector<EdJob> ed((size_t)n_open);
            std::vector<int> which((size_t)n_open);
            for (int64_t i = 0; i < n_open; ++i) {
                const VmAlnJobDev &j = J[(size_t)i];
                EdJob &e = ed[(size_t)i];
                e.read = j.read;
                e.a.src = j.q.src; e.a.lo = j.q.lo; e.a.hi = j.q.lo + j.q.len; e.a.reverse = j.q.reverse; e.a.comp = j.q.comp;
                e.b.src = j.t.src; e.b.lo = j.t.lo; e.b.hi = j.t.lo + j.t.len; e.b.reverse = j.t.reverse; e.b.comp = j.t.comp;
                e.band = j.out_off;
                which[(size_t)i] = (int)i;
            }
            be->ed_exact(ed, which);
            for (int64_t i = 0; i < n_open; ++i) {
                int64_t d = ed[(size_t)i].dist;
                to_exec(&((VmAlnJobDev *)jobs + ids[(size_t)i])->result0, &d, 8);
                sync();      // `d` is a stack variable
            }
        }
        void fill(vmd::Job *jobs, int64_t nj, int64_t scratch_words, bool eqx, vmd::U2 *res, const uint32_t **ops)
        {
            *ops = be->fill_device((VmAlnJobDev *)jobs, (int)nj, scratch_words, eqx, (uint2 *)res, fill_slot);
        }
        int fill_slot = 0;       // pass 1 / pass 2 keep their dense CIGAR arenas apart
    };

    // per-job statistics of a fill batch, summed on the device
    struct FillStats { double cells, bases; unsigned long long n_band, n_redo; };

    // Global fill of device-resident jobs, planned on the device.  results[j] = (offset, length) of job j's ops in the
    // returned dense arena (valid until the next fill_device call with the same slot).
    const uint32_t *fill_device(VmAlnJobDev *jobs, int nj, int64_t scratch_words, bool eqx, uint2 *results, int slot)
    {
        WallTimer wt(this, "fill");
        VmDevBuf &dense = slot ? d_cigd2_ : d_cigd_;
        const int sms = c_->sm_count > 0 ? c_->sm_count : 148;
        BE_OK(d_cig_.ensure((size_t)scratch_words * 4 + 64));
        BE_OK(dense.ensure((size_t)scratch_words * 4 + 64));
        BE_OK(d_mask_.ensure((size_t)nj + 64));
        BE_OK(d_pairs_.ensure((size_t)nj * (sizeof(VmFillPair) + sizeof(VmFillBandPair)) + 4096));
        BE_OK(d_fstats_.ensure(256));
        VmFillPair *d_fpairs = d_pairs_.as<VmFillPair>();
        VmFillBandPair *d_bpairs = (VmFillBandPair *)(d_fpairs + nj);
        unsigned long long *d_count = (unsigned long long *)d_fstats_.p;       // [dense count | FillStats | launch counters]
        FillStats *d_stats = (FillStats *)(d_count + 1);
        int *d_ctr = (int *)(d_stats + 1);
        static const bool no_band = getenv("VM_FILL_NO_BAND") != nullptr;
        {
            WallTimer w2(this, "h_fill_plan");
            BE_OK(cudaMemsetAsync(d_fstats_.p, 0, 256, c_->stream));
            BE_OK(cudaMemsetAsync(results, 0, (size_t)nj * 8, c_->stream));
            int nl;
            if (!no_band) {
                nl = vm_fillb_plan_dev(jobs, nj, bplanbufs_, d_bpairs, d_mask_.as<uint8_t>(), c_->stream);
                if (nl < 0) throw std::runtime_error("fill plan (banded): CUDA error");
                c_->launches += nl;
            } else BE_OK(cudaMemsetAsync(d_mask_.p, 1, (size_t)nj, c_->stream));
            nl = vm_fill_plan_dev(jobs, nj, d_mask_.as<uint8_t>(), fplanbufs_, d_fpairs, c_->stream);
            if (nl < 0) throw std::runtime_error("fill plan: CUDA error");
            c_->launches += nl + vm_launch_fill_stats(jobs, nj, &d_stats->cells, &d_stats->bases, c_->stream);
            BE_OK(vm_stream_sync(c_->stream));
            BE_OK(cudaGetLastError());
            if (!no_band) vm_fillb_plan_finish(bplanbufs_, sms, bplan_);
            else { bplan_.pairs.clear(); bplan_.launches.clear(); bplan_.dir_words = 0; bplan_.dir_bytes = 0; }
            vm_fill_plan_finish(fplanbufs_, sms, plan_);
        }
        fill_dir_bytes_ += plan_.dir_bytes + bplan_.dir_bytes;
        fill_jobs_ += nj;
        if (plan_.launches.size() + bplan_.launches.size() > 40) throw std::runtime_error("fill: too many launch classes");
        if (getenv("VM_DEBUG_MEM")) {
            size_t fr = 0, tot = 0;
            cudaMemGetInfo(&fr, &tot);
            fprintf(stderr, "[fill_device] nj=%d scratch_words=%lld band dir_words=%zu (%zu launches) full dir_words=%zu band_words=%zu (%zu launches) free=%.1f GB of %.1f\n",
                    nj, (long long)scratch_words, bplan_.dir_words, bplan_.launches.size(), plan_.dir_words, plan_.band_words, plan_.launches.size(),
                    fr / 1e9, tot / 1e9);
            for (const VmFillBandLaunch &L : bplan_.launches)
                fprintf(stderr, "   band cls %d pairs %d blocks %d words/warp %lld\n", L.cls, L.pair_end - L.pair_begin, L.blocks, L.dir_words_per_warp);
            for (const VmFillLaunch &L : plan_.launches)
                fprintf(stderr, "   full R %d mb %d pairs %d blocks %d words/warp %lld band/warp %lld\n", L.R, L.multiband, L.pair_end - L.pair_begin, L.blocks,
                        L.dir_words_per_warp, L.band_words_per_warp);
        }
        BE_OK(d_dir_.ensure((plan_.dir_words + bplan_.dir_words) * 4 + 64));
        BE_OK(d_sc_.ensure(plan_.band_words * 4 + 64));
        {
            KTimer kt(this, "k_fill");
            side_fork();
            int rr = 0;
            size_t dir_cur = 0, band_cur = 0;
            c_->launches += vm_fillb_launch(bplan_, jobs, d_bpairs, sources(), eqx ? 1 : 0, d_dir_.as<uint32_t>(), d_ctr + plan_.launches.size(),
                                            d_cig_.as<uint32_t>(), dense.as<uint32_t>(), d_count, results, c_->stream, side_, kSide, &rr, &dir_cur);
            c_->launches += vm_fill_launch(plan_, jobs, d_fpairs, sources(), eqx ? 1 : 0, d_dir_.as<uint32_t>(), d_sc_.as<uint32_t>(), d_ctr,
                                           d_cig_.as<uint32_t>(), dense.as<uint32_t>(), d_count, results, c_->stream, side_, kSide, &rr, &dir_cur,
                                           &band_cur);
            side_join();
            kt.stop();
        }
        BE_OK(h_misc_.ensure(256));
        FillStats *h_stats = h_misc_.as<FillStats>();
        if (!bplan_.launches.empty()) {
            // jobs whose certificate failed: once more, in the full-matrix kernel
            c_->launches += vm_launch_fill_redo_mask(results, nj, d_mask_.as<uint8_t>(), &d_stats->n_redo, c_->stream);
            BE_OK(cudaMemcpyAsync(h_stats, d_stats, sizeof(FillStats), cudaMemcpyDeviceToHost, c_->stream));
            BE_OK(vm_stream_sync(c_->stream));
            BE_OK(cudaGetLastError());
            fill_band_jobs_ += (double)bplanbufs_.table.as<VmFbTable>()->n_live;
            fill_band_redo_ += (double)h_stats->n_redo;
            if (h_stats->n_redo) {
                int nl = vm_fill_plan_dev(jobs, nj, d_mask_.as<uint8_t>(), fplanbufs_, d_fpairs, c_->stream);
                if (nl < 0) throw std::runtime_error("fill plan: CUDA error");
                c_->launches += nl;
                BE_OK(vm_stream_sync(c_->stream));
                vm_fill_plan_finish(fplanbufs_, sms, plan_);
                BE_OK(d_dir_.ensure(plan_.dir_words * 4 + 64));
                BE_OK(d_sc_.ensure(plan_.band_words * 4 + 64));
                BE_OK(cudaMemsetAsync(d_ctr, 0, 160, c_->stream));
                KTimer kt(this, "k_fill");
                int rr = 0;
                size_t dir_cur = 0, band_cur = 0;
                c_->launches += vm_fill_launch(plan_, jobs, d_fpairs, sources(), eqx ? 1 : 0, d_dir_.as<uint32_t>(), d_sc_.as<uint32_t>(), d_ctr,
                                               d_cig_.as<uint32_t>(), dense.as<uint32_t>(), d_count, results, c_->stream, nullptr, 0, &rr, &dir_cur,
                                               &band_cur);
                kt.stop();
            }
        } else {
            BE_OK(cudaMemcpyAsync(h_stats, d_stats, sizeof(FillStats), cudaMemcpyDeviceToHost, c_->stream));
            BE_OK(vm_stream_sync(c_->stream));
        }
        fill_cells_ += h_stats->cells;
        fill_bases_ += h_stats->bases;
        return dense.as<uint32_t>();
    }

    bool has_device_extension() const override { return device_extension; }
    bool has_device_front() const override { return device_front; }

    // hit2work_1's bookkeeping after the DP, guide selection and the re-seeding jobs of every read, on the device
    bool front_device(const ReadBatch &b, const std::vector<char> &need_reverse, const ChainOut &g, int max_guides,
                      std::vector<vmd::FrontOut> &fo) override
    {
        if (!device_front) return false;
        WallTimer wt(this, "front_device");
        static_assert(sizeof(vmd::RJob) == sizeof(VmReseedJobDev), "re-seeding job layout");
        const int64_t n = b.n;
        upload_contig_table();
        std::vector<int32_t> ids, nrev((size_t)n);
        for (int64_t r = 0; r < n; ++r) {
            nrev[(size_t)r] = need_reverse[(size_t)r] ? 1 : 0;
            if (g.cnt[(size_t)r] > 2) ids.push_back((int32_t)r);
        }
        const size_t NA = gx_.n_anc_total, NC = gx_.n_chain_total;
        BE_OK(dx_nrev_.ensure((size_t)n * 8 + 64));
        BE_OK(cudaMemcpyAsync(dx_nrev_.p, nrev.data(), (size_t)n * 4, cudaMemcpyHostToDevice, c_->stream));      // pageable: staged on return
        BE_OK(jobs_.ensure((NC + 1) * sizeof(VmReseedJobDev) + 64));
        BE_OK(d_wlo_.ensure((NA + 1) * 8 + 64));
        BE_OK(d_whi_.ensure((NA + 1) * 8 + 64));
        BE_OK(d_gx_.ensure((NA + 1) * 4 + 64));
        BE_OK(d_gy_.ensure((NA + 1) * 8 + 64));
        vmd::FrontInput in;
        in.n_reads = n;
        in.read_off = read_off_.as<int64_t>();
        in.ctg.start = d_ctg_.as<int64_t>();
        in.ctg.len = d_ctg_.as<int64_t>() + n_ctg_dev_;
        in.ctg.n = n_ctg_dev_;
        in.need_reverse = dx_nrev_.as<int32_t>();
        in.xrec = (const vmd::ExtractRec *)gx_.rec.p;
        in.anc = (const vmd::A32 *)gx_.anc.p;
        in.S = gx_.S.as<double>();
        in.chain_len = gx_.len.as<int32_t>();
        in.chain_score = gx_.score.as<double>();
        in.NA = (int64_t)NA;
        in.NC = (int64_t)NC;
        in.max_guides = max_guides;
        in.kmer = 9;
        exec_.be = this;
        if (!front_) front_.reset(new vmd::FrontHalf<CudaExec>(exec_));
        front_->run(in, ids, (vmd::RJob *)jobs_.p, d_wlo_.as<int64_t>(), d_whi_.as<int64_t>(), d_gx_.as<int32_t>(), d_gy_.as<int64_t>(), fo,
                    front_J_, front_xrec_);
        return true;
    }

    void reseed_chain_front(const ReadBatch &b, const std::vector<char> &need_reverse, const std::vector<vmd::FrontOut> &fo,
                            const std::vector<int> &variant, const std::vector<double> &skipcost, int maxdiff, int maxgap,
                            ChainOut &out) override
    {
        (void)need_reverse;
        WallTimer wt(this, "reseed_chain");
        const int64_t n = b.n;
        out = ChainOut();
        out.start.assign((size_t)n, 0);
        out.cnt.assign((size_t)n, 0);
        out.gmax.assign((size_t)n, -1);
        const size_t nj = front_J_.size();
        if (nj == 0) return;
        std::vector<VmReseedJobDev> J(nj);
        memcpy(J.data(), front_J_.data(), nj * sizeof(VmReseedJobDev));
        std::vector<int32_t> job_lo((size_t)n, 0), job_n((size_t)n, 0), n_out;
        for (int64_t r = 0; r < n; ++r)
            if (variant[(size_t)r] != 0) { job_lo[(size_t)r] = (int32_t)front_xrec_[(size_t)r].meta_off; job_n[(size_t)r] = fo[(size_t)r].n_jobs; }
        int32_t *d_n_out = nullptr;
        reseed_run(J, n_out, d_n_out, new WallTimer(this, "h_reseed_stage"), true);
        reseed_chain_finish(b, J, n_out, d_n_out, job_lo, job_n, variant, skipcost, maxdiff, maxgap, out);
    }

    // extend_func (+ second pass) for every read of `ids` (the reads that came out of the local stage), on the device
    bool extend_device(const ReadBatch &b, const std::vector<int32_t> &ids, const std::vector<char> &need_reverse,
                       const std::vector<int32_t> &mapq, const vmg::Options &opt, std::vector<int32_t> &status, FlatRecords &out) override
    {
        if (!device_extension) return false;
        WallTimer wt(this, "extend_device");
        const int64_t n = b.n;
        VmChainState &cs = c_->chain;
        upload_contig_table();
        BE_OK(dx_nrev_.ensure((size_t)n * 8 + 64));
        int32_t *d_nrev = dx_nrev_.as<int32_t>(), *d_mapq = d_nrev + n;
        std::vector<int32_t> tmp(2 * (size_t)n);
        for (int64_t r = 0; r < n; ++r) { tmp[(size_t)r] = need_reverse[(size_t)r] ? 1 : 0; tmp[(size_t)(n + r)] = mapq[(size_t)r]; }
        BE_OK(cudaMemcpyAsync(d_nrev, tmp.data(), tmp.size() * 4, cudaMemcpyHostToDevice, c_->stream));      // pageable: staged on return
        vmd::BackInput in;
        in.n_reads = n;
        in.read_off = read_off_.as<int64_t>();
        in.reads_fwd = reads_fwd_.as<uint8_t>();
        in.reads_rc = reads_rc_.as<uint8_t>();
        in.ref = ih_->ix->dev.ref;
        in.ctg.start = d_ctg_.as<int64_t>();
        in.ctg.len = d_ctg_.as<int64_t>() + n_ctg_dev_;
        in.ctg.n = n_ctg_dev_;
        in.need_reverse = d_nrev;
        in.mapq = d_mapq;
        in.local_cnt = cs.cnt_dev.as<int32_t>();
        in.xrec = (const vmd::ExtractRec *)lx_.rec.p;
        in.rrec = (const vmd::RebuildRec *)lx_.al_rec.p;
        in.al_anc = (const vmd::A32 *)lx_.al_anc.p;
        in.al_len = lx_.al_len.as<int32_t>();
        in.NA = (int64_t)lx_.al_n_al;
        in.NT = (int64_t)lx_.al_n_anc;
        in.total_bases = b.off[n] - b.off[0];
        vmd::BackParams p;
        p.maxdivergence = opt.maxdivergence;
        p.eqx = opt.eqx; p.hardclip = opt.hardclip; p.nodiscard = opt.nodiscard;
        if (!back_) back_.reset(new vmd::BackHalf<CudaExec>(exec_));
        exec_.be = this;
        vmd::BackResult br;
        back_->run(in, p, ids, status, br);
        out.rec_off.swap(br.rec_off);
        out.recs = br.recs;
        out.cigar = br.cigar;
        out.n_rec = br.n_rec;
        out.n_ops = br.n_ops;
        for (int k = 0; k < vmd::CT_COUNT; ++k) out.counters[k] = br.counters[k];
        return true;
    }

    // contig starts and lengths on the device: [starts | lens]
    void upload_contig_table()
    {
        if (d_ctg_for_ == ih_ && d_ctg_has_len_) return;
        const std::vector<int64_t> &st = ih_->ctg.start, &ln = ih_->ctg.len;
        std::vector<int64_t> both(st);
        both.insert(both.end(), ln.begin(), ln.end());
        BE_OK(d_ctg_.ensure(both.size() * 8 + 64));
        BE_OK(cudaMemcpyAsync(d_ctg_.p, both.data(), both.size() * 8, cudaMemcpyHostToDevice, c_->stream));
        BE_OK(vm_stream_sync(c_->stream));
        n_ctg_dev_ = (int)st.size();
        d_ctg_for_ = ih_;
        d_ctg_has_len_ = true;
    }

private:
    CudaExec exec_{nullptr};
    std::unique_ptr<vmd::BackHalf<CudaExec>> back_;
    std::unique_ptr<vmd::FrontHalf<CudaExec>> front_;
    std::vector<vmd::RJob> front_J_;
    std::vector<vmd::ExtractRec> front_xrec_;
    VmDevBuf dx_nrev_, d_mask_, d_fstats_, d_cigd2_;
    VmFillPlanBufs bplanbufs_, fplanbufs_;
    bool d_ctg_has_len_ = false;
    vm_ctx *c_;
    vm_index_handle *ih_;
    VmSeedBufs seed_;
    VmDevBuf reads_fwd_, reads_rc_, read_off_, jobs_, d_wlo_, d_whi_, d_gx_, d_gy_, d_nh_, d_hits_, d_tab_, d_order_, d_rout_,
        d_dense_, d_seg_, d_dir_, d_sc_, d_cig_, d_cigd_, d_pairs_, d_msegs_;
    VmFillPlan plan_;
    VmFillBandPlan bplan_;
    cudaStream_t side_[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t side_done_[3] = {nullptr, nullptr, nullptr}, side_go_ = nullptr;
    Extracted gx_, lx_;
    bool want_local_score_ = false;
    int rebuild_large_cost_ = -1;          // >= 0 while the local extraction should also rebuild the sub-alignments
    VmDevBuf d_ctg_;
    int n_ctg_dev_ = 0;
    const vm_index_handle *d_ctg_for_ = nullptr;
    VmDevBuf x_ids_, x_used_, x_tmp_anc_, x_tmp_S_, x_tmp_len_, x_tmp_score_;
    VmPinnedBuf h_sorted_, h_S_, h_P_, h_A_, h_gmax_, h_jobs_, h_cig_, h_lsorted_, h_lP_, h_lgmax_, h_misc_, h_gx_, h_gy_, h_segs_, h_small_;
    std::vector<int64_t> off_host_;
};

} // namespace
