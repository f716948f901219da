// CUDA backend of the batch driver (vm_pipeline.hpp) and the vm_index_* / vm_align_* C ABI.
//
// Every Backend method uploads its job descriptors, launches the kernels of its stage on the
// context stream and reads the results back; reads, reference and index stay resident in HBM
// for the whole batch.  There is no CPU implementation of any stage behind this interface.
#include "vm_ctx.cuh"
#include "vm_chain.cuh"
#include "vm_index.cuh"
#include "vm_seed.cuh"
#include "vm_reseed.cuh"
#include "vm_align.cuh"
#include "vm_pipeline.hpp"
#include <chrono>
#include <map>

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

namespace {

struct StageTimer {
    std::map<std::string, double> ms;
    void add(const char *k, double v) { ms[k] += v; }
};

class CudaBackend : public Backend {
public:
    CudaBackend(vm_ctx *c, vm_index_handle *ih) : c_(c), ih_(ih) {}
    ~CudaBackend() override
    {
        seed_.release();
        VmDevBuf *b[] = {&reads_fwd_, &reads_rc_, &read_off_, &jobs_, &aux0_, &aux1_, &aux2_, &aux3_, &aux4_, &aux5_, &aux6_,
                         &aux7_, &aux8_};
        for (VmDevBuf *x : b) x->release();
    }
    StageTimer timer;
    bool reads_resident = false;    // vm_reads_upload already put this batch in HBM
    void set_index(vm_index_handle *ih) { ih_ = ih; }

    // device time of a group of launches, CUDA events on the ctx stream
    struct KTimer {
        CudaBackend *be; const char *name; cudaEvent_t a, b;
        KTimer(CudaBackend *be_, const char *n) : be(be_), name(n)
        {
            cudaEventCreate(&a); cudaEventCreate(&b);
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
        }
    };

    void upload_reads(const ReadBatch &b)
    {
        if (reads_resident && off_host_.size() == (size_t)b.n + 1 && std::equal(off_host_.begin(), off_host_.end(), b.off)) return;
        const size_t total = (size_t)b.off[b.n];
        std::string rc(total, 'N');
        for (int64_t r = 0; r < b.n; ++r) {
            const int64_t lo = b.off[r], L = b.len(r);
            for (int64_t i = 0; i < L; ++i) {
                const char ch = b.seq[lo + L - 1 - i];
                rc[lo + i] = ch == 'A' ? 'T' : ch == 'C' ? 'G' : ch == 'G' ? 'C' : ch == 'T' ? 'A' : 'N';
            }
        }
        BE_OK(reads_fwd_.ensure(total + 64));
        BE_OK(reads_rc_.ensure(total + 64));
        BE_OK(read_off_.ensure((size_t)(b.n + 1) * 8));
        BE_OK(cudaMemcpyAsync(reads_fwd_.p, b.seq, total, cudaMemcpyHostToDevice, c_->stream));
        BE_OK(cudaMemcpyAsync(reads_rc_.p, rc.data(), total, cudaMemcpyHostToDevice, c_->stream));
        BE_OK(cudaMemcpyAsync(read_off_.p, b.off, (size_t)(b.n + 1) * 8, cudaMemcpyHostToDevice, c_->stream));
        BE_OK(cudaStreamSynchronize(c_->stream));
        off_host_.assign(b.off, b.off + b.n + 1);
    }

    void seed(const ReadBatch &b, int check_num, Ragged<Anc> &anchors, std::vector<char> &need_reverse) override
    {
        auto t0 = std::chrono::steady_clock::now();
        upload_reads(b);
        std::vector<int32_t> n_out, nrev;
        std::vector<int64_t> a_off;
        std::string err;
        if (vm_seed_batch(seed_, ih_->ix->dev, reads_fwd_.as<uint8_t>(), read_off_.as<int64_t>(), off_host_, check_num, -1,
                          c_->stream, n_out, nrev, a_off, &c_->launches, err))
            throw std::runtime_error(err);
        std::vector<VmAnchor> flat((size_t)a_off[b.n]);
        if (!flat.empty())
            BE_OK(cudaMemcpy(flat.data(), seed_.out.p, flat.size() * sizeof(VmAnchor), cudaMemcpyDeviceToHost));
        anchors.clear();
        need_reverse.assign((size_t)b.n, 0);
        for (int64_t r = 0; r < b.n; ++r) {
            need_reverse[r] = (char)nrev[r];
            for (int32_t t = 0; t < n_out[r]; ++t) {
                const VmAnchor &a = flat[(size_t)a_off[r] + t];
                anchors.data.push_back(Anc{a.x, (int64_t)a.y, a.s, a.l});
            }
            anchors.close_row();
        }
        timer.add("seed", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }

    // shared by the global and local DPs: rows in, sorted / S / P / S_arg / gmax out
    void run_chain(const Ragged<Anc> &anchors, const std::vector<int64_t> &ids, const std::vector<int64_t> &read_len,
                   const vm_chain_params &prm, ChainOut &out, bool want_S)
    {
        std::vector<int64_t> rows, off(1, 0);
        std::vector<int32_t> rl;
        for (int64_t r : ids) {
            const Anc *a = anchors.row(r);
            for (int64_t t = 0; t < anchors.size(r); ++t) {
                rows.push_back(a[t].x); rows.push_back(a[t].y); rows.push_back(a[t].s); rows.push_back(a[t].l);
            }
            off.push_back((int64_t)rows.size() / 4);
            rl.push_back((int32_t)std::min<int64_t>(read_len[r], INT32_MAX));
        }
        const int64_t n = (int64_t)ids.size(), T = off.back();
        chain_anchors_ += (double)T;
        std::vector<int64_t> srt((size_t)T * 4), gmax((size_t)n);
        std::vector<double> S((size_t)T);
        std::vector<int32_t> P((size_t)T), A((size_t)T), uf((size_t)n);
        float ms = 0;
        int rc = vm_chain_global_batch(c_, &prm, n, rows.data(), off.data(), rl.data(), srt.data(), S.data(), P.data(), A.data(),
                                       gmax.data(), uf.data(), &ms);
        if (rc != VM_OK) throw std::runtime_error("chain: " + c_->err);
        timer.add(prm.variant == 0 ? "chain_global_kernels" : "chain_local_kernels", ms);
        for (int64_t t = 0; t < n; ++t) {
            const int64_t r = ids[t];
            const int64_t o = out.sorted.off[r], m = off[t + 1] - off[t];
            for (int64_t q = 0; q < m; ++q) {
                const int64_t *s = srt.data() + (off[t] + q) * 4;
                out.sorted.data[o + q] = Anc{s[0], s[1], (int32_t)s[2], (int32_t)s[3]};
                out.P[o + q] = P[off[t] + q];
                if (want_S) { out.S[o + q] = S[off[t] + q]; out.S_arg[o + q] = A[off[t] + q]; }
            }
            out.gmax[r] = gmax[t];
        }
    }

    static void prepare_out(const Ragged<Anc> &anchors, ChainOut &out, bool want_S)
    {
        out = ChainOut();
        out.sorted.off = anchors.off;
        out.sorted.data.resize(anchors.data.size());
        out.P.resize(anchors.data.size());
        if (want_S) { out.S.resize(anchors.data.size()); out.S_arg.resize(anchors.data.size()); }
        out.gmax.assign((size_t)anchors.rows(), -1);
    }

    void chain_global(const Ragged<Anc> &anchors, const std::vector<int64_t> &read_len, int kmersize, double skipcost,
                      int maxdiff, int maxgap, ChainOut &out) override
    {
        auto t0 = std::chrono::steady_clock::now();
        prepare_out(anchors, out, true);
        std::vector<int64_t> ids;
        for (int64_t r = 0; r < anchors.rows(); ++r)
            if (anchors.size(r) > 0) ids.push_back(r);
        vm_chain_params prm{kmersize, skipcost, maxdiff, maxgap, 1000, 5, 30, 0};
        if (!ids.empty()) run_chain(anchors, ids, read_len, prm, out, true);
        timer.add("chain_global", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }

    void chain_local(const Ragged<Anc> &anchors, const std::vector<int> &variant, const std::vector<double> &skipcost,
                     int maxdiff, int maxgap, ChainOut &out) override
    {
        auto t0 = std::chrono::steady_clock::now();
        prepare_out(anchors, out, false);
        std::map<std::pair<int, double>, std::vector<int64_t>> groups;
        for (int64_t r = 0; r < anchors.rows(); ++r)
            if (variant[r] != 0 && anchors.size(r) > 0) groups[{variant[r], skipcost[r]}].push_back(r);
        std::vector<int64_t> read_len((size_t)anchors.rows(), 1 << 30);   // unused by the local DP (no n/L rule)
        for (auto &g : groups) {
            vm_chain_params prm{9, g.first.second, maxdiff, maxgap, 1000, 5, 30, g.first.first};
            run_chain(anchors, g.second, read_len, prm, out, false);
        }
        timer.add("chain_local", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }

    void reseed(const ReadBatch &b, const std::vector<char> &need_reverse, const std::vector<GuideJobRef> &jobs,
                Ragged<Anc> &local) override
    {
        auto t0 = std::chrono::steady_clock::now();
        const int nj = (int)jobs.size();
        local.clear();
        if (nj == 0) {
            for (int64_t r = 0; r < b.n; ++r) local.close_row();
            return;
        }
        std::vector<VmReseedJobDev> J((size_t)nj);
        std::vector<int64_t> wlo, whi, gy;
        std::vector<int32_t> gx;
        for (int j = 0; j < nj; ++j) {
            const vmg::GuideJob &g = jobs[j].job;
            VmReseedJobDev &d = J[j];
            memset(&d, 0, sizeof(d));
            d.read = jobs[j].read;
            d.need_reverse = need_reverse[jobs[j].read] ? 1 : 0;
            d.readstart = g.readstart;
            d.readend = g.readend;
            d.n_win = (int32_t)g.win_lo.size();
            d.n_guide = (int32_t)g.gx.size();
            d.win_off = (int64_t)wlo.size();
            d.g_off = (int64_t)gx.size();
            wlo.insert(wlo.end(), g.win_lo.begin(), g.win_lo.end());
            whi.insert(whi.end(), g.win_hi.begin(), g.win_hi.end());
            gx.insert(gx.end(), g.gx.begin(), g.gx.end());
            gy.insert(gy.end(), g.gy.begin(), g.gy.end());
            d.count_only = 1;
        }
        VmDevBuf &d_jobs = jobs_, &d_wlo = aux0_, &d_whi = aux1_, &d_gx = aux2_, &d_gy = aux3_, &d_nh = aux4_, &d_hits = aux5_,
                 &d_tab = aux6_, &d_order = aux7_, &d_out = aux8_;
        BE_OK(d_jobs.ensure(J.size() * sizeof(VmReseedJobDev)));
        BE_OK(d_wlo.ensure(wlo.size() * 8 + 64));
        BE_OK(d_whi.ensure(whi.size() * 8 + 64));
        BE_OK(d_gx.ensure(gx.size() * 4 + 64));
        BE_OK(d_gy.ensure(gy.size() * 8 + 64));
        BE_OK(d_nh.ensure((size_t)nj * 8 + 64));
        BE_OK(cudaMemcpyAsync(d_jobs.p, J.data(), J.size() * sizeof(VmReseedJobDev), cudaMemcpyHostToDevice, c_->stream));
        BE_OK(cudaMemcpyAsync(d_wlo.p, wlo.data(), wlo.size() * 8, cudaMemcpyHostToDevice, c_->stream));
        BE_OK(cudaMemcpyAsync(d_whi.p, whi.data(), whi.size() * 8, cudaMemcpyHostToDevice, c_->stream));
        BE_OK(cudaMemcpyAsync(d_gx.p, gx.data(), gx.size() * 4, cudaMemcpyHostToDevice, c_->stream));
        BE_OK(cudaMemcpyAsync(d_gy.p, gy.data(), gy.size() * 8, cudaMemcpyHostToDevice, c_->stream));
        int32_t *d_n_hits = d_nh.as<int32_t>(), *d_over = d_nh.as<int32_t>() + nj;
        BE_OK(cudaMemsetAsync(d_over, 0, 4, c_->stream));
        const VmIndexDev &ix = ih_->ix->dev;
        // pass 1: count hits
        KTimer kt1(this, "k_reseed_hits");
        c_->launches += vm_reseed_launch(ix, d_jobs.as<VmReseedJobDev>(), nj, reads_fwd_.as<uint8_t>(), reads_rc_.as<uint8_t>(),
                                         read_off_.as<int64_t>(), d_wlo.as<int64_t>(), d_whi.as<int64_t>(), d_gx.as<int32_t>(),
                                         d_gy.as<int64_t>(), nullptr, d_n_hits, d_over, nullptr, nullptr, nullptr, nullptr,
                                         c_->stream);
        kt1.stop();
        std::vector<int32_t> n_hits((size_t)nj);
        BE_OK(cudaMemcpyAsync(n_hits.data(), d_n_hits, (size_t)nj * 4, cudaMemcpyDeviceToHost, c_->stream));
        BE_OK(cudaStreamSynchronize(c_->stream));
        int64_t hit_off = 0, tab_off = 0;
        for (int j = 0; j < nj; ++j) {
            J[j].count_only = 0;
            J[j].hit_off = hit_off;
            J[j].hit_cap = n_hits[j];
            hit_off += n_hits[j];
            int ts = 64;
            while (ts < n_hits[j] + 8) ts <<= 1;
            J[j].tab_off = tab_off;
            J[j].tab_size = ts;
            tab_off += ts;
        }
        BE_OK(d_hits.ensure((size_t)hit_off * vm_reseed_hit_bytes() + 64));
        BE_OK(d_tab.ensure((size_t)tab_off * vm_reseed_point_bytes() + 64));
        BE_OK(d_order.ensure((size_t)hit_off * 4 + 64));
        BE_OK(d_out.ensure((size_t)hit_off * 2 * sizeof(VmAnchor) + 64));
        BE_OK(cudaMemcpyAsync(d_jobs.p, J.data(), J.size() * sizeof(VmReseedJobDev), cudaMemcpyHostToDevice, c_->stream));
        int32_t *d_n_out = d_nh.as<int32_t>();   // reuse after the counts are on the host
        KTimer kt2(this, "k_reseed_hits");
        c_->launches += vm_reseed_launch(ix, d_jobs.as<VmReseedJobDev>(), nj, reads_fwd_.as<uint8_t>(), reads_rc_.as<uint8_t>(),
                                         read_off_.as<int64_t>(), d_wlo.as<int64_t>(), d_whi.as<int64_t>(), d_gx.as<int32_t>(),
                                         d_gy.as<int64_t>(), d_hits.p, d_n_hits, d_over, nullptr, nullptr, nullptr, nullptr,
                                         c_->stream);
        kt2.stop();
        KTimer kt3(this, "k_reseed_merge");
        c_->launches += vm_reseed_merge_launch(d_jobs.as<VmReseedJobDev>(), nj, d_hits.p, d_n_hits, d_tab.p, d_order.as<int32_t>(),
                                               d_out.as<VmAnchor>(), d_n_out + nj + 1, c_->stream);
        kt3.stop();
        reseed_hits_ += (double)hit_off;
        std::vector<int32_t> n_out((size_t)nj), over(1);
        BE_OK(d_nh.ensure((size_t)nj * 8 + 64));
        BE_OK(cudaMemcpyAsync(n_out.data(), d_n_out + nj + 1, (size_t)nj * 4, cudaMemcpyDeviceToHost, c_->stream));
        BE_OK(cudaMemcpyAsync(over.data(), d_over, 4, cudaMemcpyDeviceToHost, c_->stream));
        std::vector<VmAnchor> flat((size_t)hit_off * 2);
        if (!flat.empty())
            BE_OK(cudaMemcpyAsync(flat.data(), d_out.p, flat.size() * sizeof(VmAnchor), cudaMemcpyDeviceToHost, c_->stream));
        BE_OK(cudaStreamSynchronize(c_->stream));
        BE_OK(cudaGetLastError());
        if (over[0]) throw std::runtime_error("reseed: hit buffer overflow (count and fill passes disagree)");
        size_t q = 0;
        for (int64_t r = 0; r < b.n; ++r) {
            while (q < jobs.size() && jobs[q].read == r) {
                const VmAnchor *src = flat.data() + 2 * J[q].hit_off;
                for (int32_t t = 0; t < n_out[q]; ++t) local.data.push_back(Anc{src[t].x, (int64_t)src[t].y, src[t].s, src[t].l});
                ++q;
            }
            local.close_row();
        }
        timer.add("reseed", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
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

    void edit_distance(const ReadBatch &, std::vector<EdJob> &jobs) override
    {
        auto t0 = std::chrono::steady_clock::now();
        const int nj = (int)jobs.size();
        if (nj == 0) return;
        std::vector<VmAlnJobDev> J((size_t)nj);
        int max_words = 1;
        for (int j = 0; j < nj; ++j) {
            memset(&J[j], 0, sizeof(VmAlnJobDev));
            // the shorter sequence is the bit-vector pattern (fewer 64-row blocks)
            const bool a_short = jobs[j].a.len() <= jobs[j].b.len();
            J[j].q = spec(a_short ? jobs[j].a : jobs[j].b);
            J[j].t = spec(a_short ? jobs[j].b : jobs[j].a);
            J[j].read = jobs[j].read;
            max_words = std::max(max_words, (J[j].q.len + 63) / 64);
        }
        if (max_words > 32 * 64) throw std::runtime_error("edit distance: sequence longer than 131072 bases is not supported yet");
        BE_OK(jobs_.ensure(J.size() * sizeof(VmAlnJobDev)));
        BE_OK(cudaMemcpyAsync(jobs_.p, J.data(), J.size() * sizeof(VmAlnJobDev), cudaMemcpyHostToDevice, c_->stream));
        KTimer kt(this, "k_edit_distance");
        c_->launches += vm_launch_edit_distance(jobs_.as<VmAlnJobDev>(), nj, sources(), max_words, c_->stream);
        kt.stop();
        for (int j = 0; j < nj; ++j) ed_cells_ += (double)J[j].q.len * (double)J[j].t.len;
        BE_OK(cudaMemcpyAsync(J.data(), jobs_.p, J.size() * sizeof(VmAlnJobDev), cudaMemcpyDeviceToHost, c_->stream));
        BE_OK(cudaStreamSynchronize(c_->stream));
        BE_OK(cudaGetLastError());
        for (int j = 0; j < nj; ++j) jobs[j].dist = J[j].result0;
        timer.add("edit_distance", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }

    void extend(const ReadBatch &, std::vector<ExtJobRef> &jobs) override
    {
        auto t0 = std::chrono::steady_clock::now();
        const int nj = (int)jobs.size();
        if (nj == 0) return;
        std::vector<VmAlnJobDev> J((size_t)nj);
        for (int j = 0; j < nj; ++j) {
            memset(&J[j], 0, sizeof(VmAlnJobDev));
            J[j].t = spec(jobs[j].job.target);
            J[j].q = spec(jobs[j].job.query);
            J[j].read = jobs[j].read;
        }
        BE_OK(jobs_.ensure(J.size() * sizeof(VmAlnJobDev)));
        BE_OK(cudaMemcpyAsync(jobs_.p, J.data(), J.size() * sizeof(VmAlnJobDev), cudaMemcpyHostToDevice, c_->stream));
        KTimer kt(this, "k_extend");
        c_->launches += vm_launch_extend(jobs_.as<VmAlnJobDev>(), nj, sources(), c_->stream);
        kt.stop();
        BE_OK(cudaMemcpyAsync(J.data(), jobs_.p, J.size() * sizeof(VmAlnJobDev), cudaMemcpyDeviceToHost, c_->stream));
        BE_OK(cudaStreamSynchronize(c_->stream));
        BE_OK(cudaGetLastError());
        for (int j = 0; j < nj; ++j) { jobs[j].job.q_e = (int32_t)J[j].result0; jobs[j].job.t_e = (int32_t)J[j].result1; }
        timer.add("extend", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }

    void fill(const ReadBatch &, bool eqx, std::vector<FillJobRef> &jobs) override
    {
        auto t0 = std::chrono::steady_clock::now();
        const int nj = (int)jobs.size();
        if (nj == 0) return;
        std::vector<VmAlnJobDev> J((size_t)nj);
        int64_t out_off = 0, dir_off = 0, sc_off = 0;
        const int band_rows = vm_fill_band_rows();
        for (int j = 0; j < nj; ++j) {
            memset(&J[j], 0, sizeof(VmAlnJobDev));
            J[j].t = spec(jobs[j].job.target);
            J[j].q = spec(jobs[j].job.query);
            J[j].read = jobs[j].read;
            J[j].out_off = out_off;
            J[j].dir_off = dir_off;
            out_off += (int64_t)J[j].t.len + J[j].q.len + 2;
            dir_off += (int64_t)((vm_fill_dir_bytes(J[j].t.len, J[j].q.len) + 7) & ~(size_t)7);
            if (J[j].t.len > band_rows) { J[j].sc_off = sc_off; sc_off += 3LL * J[j].q.len; }
            else J[j].sc_off = -1;
            fill_cells_ += (double)J[j].t.len * (double)J[j].q.len;
            fill_bases_ += (double)J[j].t.len + (double)J[j].q.len;
        }
        fill_jobs_ += nj;
        VmDevBuf &d_dir = aux5_, &d_sc = aux6_, &d_cig = aux8_;
        BE_OK(jobs_.ensure(J.size() * sizeof(VmAlnJobDev)));
        BE_OK(d_dir.ensure((size_t)dir_off + 64));
        BE_OK(d_sc.ensure((size_t)sc_off * 4 + 64));
        BE_OK(d_cig.ensure((size_t)out_off * 4 + 64));
        BE_OK(cudaMemcpyAsync(jobs_.p, J.data(), J.size() * sizeof(VmAlnJobDev), cudaMemcpyHostToDevice, c_->stream));
        KTimer kt(this, "k_fill");
        c_->launches += vm_launch_fill(jobs_.as<VmAlnJobDev>(), nj, sources(), eqx ? 1 : 0, d_dir.as<uint8_t>(),
                                       d_sc.as<int32_t>(), d_cig.as<uint32_t>(), c_->stream);
        kt.stop();
        std::vector<uint32_t> cig((size_t)out_off);
        BE_OK(cudaMemcpyAsync(J.data(), jobs_.p, J.size() * sizeof(VmAlnJobDev), cudaMemcpyDeviceToHost, c_->stream));
        BE_OK(cudaMemcpyAsync(cig.data(), d_cig.p, cig.size() * 4, cudaMemcpyDeviceToHost, c_->stream));
        BE_OK(cudaStreamSynchronize(c_->stream));
        BE_OK(cudaGetLastError());
        for (int j = 0; j < nj; ++j) jobs[j].cigar.assign(cig.begin() + J[j].out_off, cig.begin() + J[j].out_off + J[j].n_out);
        timer.add("fill", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }

    double fill_cells_ = 0, fill_bases_ = 0, fill_jobs_ = 0, ed_cells_ = 0, reseed_hits_ = 0, chain_anchors_ = 0;
    void reset_counters()
    {
        timer.ms.clear();
        fill_cells_ = fill_bases_ = fill_jobs_ = ed_cells_ = reseed_hits_ = chain_anchors_ = 0;
    }

private:
    vm_ctx *c_;
    vm_index_handle *ih_;
    VmSeedBufs seed_;
    VmDevBuf reads_fwd_, reads_rc_, read_off_, jobs_, aux0_, aux1_, aux2_, aux3_, aux4_, aux5_, aux6_, aux7_, aux8_;
    std::vector<int64_t> off_host_;
};

} // namespace

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
struct vm_result {
    std::vector<int64_t> rec_off;        // per read
    std::vector<vm_record> recs;
    std::vector<uint32_t> cigar;
    std::vector<double> stage_ms;
    std::vector<std::string> stage_names;
    std::string stage_text;
};

extern "C" {

int vm_index_create(vm_ctx *c, int32_t n_contigs, const char *const *names, const char *const *seqs, const int64_t *lens,
                    int32_t w, int32_t k, vm_index_handle **out)
{
    if (!c) return VM_ERR_ARG;
    if (!out || n_contigs <= 0 || !names || !seqs || !lens || k < 1 || k > 28 || w < 1 || w > 255) {
        c->err = "bad argument";
        return VM_ERR_ARG;
    }
    cudaSetDevice(c->device);
    std::vector<std::string> nm, sq;
    for (int i = 0; i < n_contigs; ++i) {
        nm.emplace_back(names[i]);
        sq.emplace_back(seqs[i], (size_t)lens[i]);
    }
    vm_index_handle *h = new vm_index_handle();
    h->ix = vm_index_build_host(nm, sq, w, k);
    if ((int64_t)h->ix->ref.size() >= (1LL << 32) - 64) {
        c->err = "reference longer than 2^32 bases is not supported";
        vm_index_free(h->ix);
        delete h;
        return VM_ERR_ARG;
    }
    std::string err;
    if (vm_index_upload(h->ix, err)) {
        c->err = err;
        vm_index_free(h->ix);
        delete h;
        return VM_ERR_CUDA;
    }
    h->ctg.names = h->ix->names;
    h->ctg.start = h->ix->ctg_start;
    h->ctg.len = h->ix->ctg_len;
    h->ctg.seq = h->ix->ref.data();
    h->ctg.total = (int64_t)h->ix->ref.size();
    *out = h;
    return VM_OK;
}

void vm_index_destroy(vm_index_handle *h)
{
    if (!h) return;
    vm_index_free(h->ix);
    delete h;
}

int vm_index_info(vm_index_handle *h, int32_t *k, int32_t *w, int32_t *n_contigs, int64_t *n_minimizers, int64_t *n_keys,
                  int32_t *mid_occ)
{
    if (!h) return VM_ERR_ARG;
    if (k) *k = h->ix->k;
    if (w) *w = h->ix->w;
    if (n_contigs) *n_contigs = (int32_t)h->ix->names.size();
    if (n_minimizers) *n_minimizers = (int64_t)h->ix->occ.size();
    if (n_keys) *n_keys = h->ix->n_keys;
    if (mid_occ) *mid_occ = h->ix->mid_occ_default;
    return VM_OK;
}

int vm_index_contig(vm_index_handle *h, int32_t i, const char **name, int64_t *start, int64_t *len, const char **seq)
{
    if (!h || i < 0 || i >= (int32_t)h->ix->names.size()) return VM_ERR_ARG;
    if (name) *name = h->ix->names[i].c_str();
    if (start) *start = h->ix->ctg_start[i];
    if (len) *len = h->ix->ctg_len[i];
    if (seq) *seq = h->ix->ref.data() + h->ix->ctg_start[i];
    return VM_OK;
}

static int vm_align_impl(vm_ctx *c, vm_index_handle *h, const vm_align_params *p, int64_t n_reads, const char *seqs,
                         const int64_t *seq_off, int resident, vm_result **out)
{
    if (!c) return VM_ERR_ARG;
    if (!h || !p || !out || n_reads < 0 || !seq_off || (n_reads > 0 && !seqs)) { c->err = "bad argument"; return VM_ERR_ARG; }
    if (c->n_extra == 0) { c->err = "vm_set_tables must be called first"; return VM_ERR_STATE; }
    cudaSetDevice(c->device);
    vmg::Options opt;
    opt.global_skipcost = p->global_skipcost;
    opt.local_skipcost = p->local_skipcost;
    opt.maxdivergence = p->maxdivergence;
    opt.global_maxdiff = p->global_maxdiff;
    opt.local_maxdiff = p->local_maxdiff;
    opt.check_num = p->check_num;
    opt.eqx = p->eqx != 0;
    opt.hardclip = p->hardclip != 0;
    opt.nodiscard = p->nodiscard != 0;
    opt.mode = vmg::ModeConst{p->accept_score, p->max_guides, p->local_maxgap, p->clamp40 != 0};
    vm_result *res = new vm_result();
    try {
        if (!c->backend) {
            c->backend = new CudaBackend(c, h);
            c->backend_free = [](void *p) { delete (CudaBackend *)p; };
        }
        CudaBackend &be = *(CudaBackend *)c->backend;
        be.set_index(h);
        be.reset_counters();
        be.reads_resident = resident != 0;
        int threads = p->host_threads > 0 ? p->host_threads : (int)std::max(1u, std::thread::hardware_concurrency());
        Driver drv(be, h->ctg, opt, h->ix->k, threads);
        ReadBatch b;
        b.n = n_reads;
        b.seq = seqs;
        b.off = seq_off;
        BatchResult br;
        auto t0 = std::chrono::steady_clock::now();
        drv.align_batch(b, br);
        const double total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        res->rec_off.assign((size_t)n_reads + 1, 0);
        for (int64_t r = 0; r < n_reads; ++r) {
            for (const vmg::Record &rec : br.records[r]) {
                vm_record o;
                o.contig = rec.contig;
                o.strand = rec.strand;
                o.q_st = rec.q_st; o.q_en = rec.q_en; o.r_st = rec.r_st; o.r_en = rec.r_en;
                o.mapq = rec.mapq;
                o.cigar_off = (int64_t)res->cigar.size();
                o.cigar_len = (int32_t)rec.cigar.size();
                res->cigar.insert(res->cigar.end(), rec.cigar.begin(), rec.cigar.end());
                res->recs.push_back(o);
            }
            res->rec_off[r + 1] = (int64_t)res->recs.size();
        }
        be.timer.add("total", total);
        be.timer.add("n_fill_cells", be.fill_cells_);
        be.timer.add("n_fill_bases", be.fill_bases_);
        be.timer.add("n_fill_jobs", be.fill_jobs_);
        be.timer.add("n_ed_cells", be.ed_cells_);
        be.timer.add("n_reseed_hits", be.reseed_hits_);
        be.timer.add("n_chain_anchors", be.chain_anchors_);
        for (auto &kv : be.timer.ms) {
            res->stage_names.push_back(kv.first);
            res->stage_ms.push_back(kv.second);
            res->stage_text += kv.first + "=" + std::to_string(kv.second) + ";";
        }
    } catch (const std::exception &e) {
        c->err = e.what();
        delete res;
        return VM_ERR_CUDA;
    }
    *out = res;
    return VM_OK;
}

int vm_align_batch(vm_ctx *c, vm_index_handle *h, const vm_align_params *p, int64_t n_reads, const char *seqs,
                   const int64_t *seq_off, vm_result **out)
{
    return vm_align_impl(c, h, p, n_reads, seqs, seq_off, 0, out);
}

// Stage-level entry point for the base-level kernels on raw sequence pairs (parity tests).
// kind 0: global edit distance -> out0[j]; kind 1: z-drop edge extension -> out0 = q_e, out1 = t_e;
// kind 2: global fill -> CIGAR ops of pair j at cigar[cig_off[j] .. cig_off[j] + out0[j]),
//         cig_off[j] = sum over i < j of (tlen_i + qlen_i + 2).
int vm_pairs_batch(vm_ctx *c, int32_t kind, int32_t eqx, int64_t n_pairs, const char *targets, const int64_t *t_off,
                   const char *queries, const int64_t *q_off, int64_t *out0, int64_t *out1, uint32_t *cigar)
{
    if (!c) return VM_ERR_ARG;
    if (n_pairs < 0 || !t_off || !q_off || !out0 || kind < 0 || kind > 2) { c->err = "bad argument"; return VM_ERR_ARG; }
    cudaSetDevice(c->device);
    try {
        if (!c->backend) {
            c->backend = new CudaBackend(c, nullptr);
            c->backend_free = [](void *p) { delete (CudaBackend *)p; };
        }
        CudaBackend &be = *(CudaBackend *)c->backend;
        // one pseudo-read holding every target followed by every query
        const int64_t tt = t_off[n_pairs], tq = q_off[n_pairs];
        std::string cat((size_t)(tt + tq), 'N');
        if (tt) memcpy(&cat[0], targets, (size_t)tt);
        if (tq) memcpy(&cat[(size_t)tt], queries, (size_t)tq);
        for (char &ch : cat) ch = "ACGTN"[vm_nt4((unsigned char)ch)];
        const int64_t off[2] = {0, tt + tq};
        ReadBatch b;
        b.n = 1; b.seq = cat.data(); b.off = off;
        be.reads_resident = false;
        be.upload_reads(b);
        be.reset_counters();
        auto ref_of = [&](int64_t lo, int64_t hi) { vmg::SeqRef s; s.src = 1; s.lo = lo; s.hi = hi; return s; };
        if (kind == 0) {
            std::vector<EdJob> jobs((size_t)n_pairs);
            for (int64_t j = 0; j < n_pairs; ++j) {
                jobs[j].read = 0;
                jobs[j].a = ref_of(tt + q_off[j], tt + q_off[j + 1]);
                jobs[j].b = ref_of(t_off[j], t_off[j + 1]);
            }
            be.edit_distance(b, jobs);
            for (int64_t j = 0; j < n_pairs; ++j) out0[j] = jobs[j].dist;
        } else if (kind == 1) {
            std::vector<ExtJobRef> jobs((size_t)n_pairs);
            for (int64_t j = 0; j < n_pairs; ++j) {
                jobs[j].read = 0;
                jobs[j].job.target = ref_of(t_off[j], t_off[j + 1]);
                jobs[j].job.query = ref_of(tt + q_off[j], tt + q_off[j + 1]);
            }
            be.extend(b, jobs);
            for (int64_t j = 0; j < n_pairs; ++j) { out0[j] = jobs[j].job.q_e; if (out1) out1[j] = jobs[j].job.t_e; }
        } else {
            std::vector<FillJobRef> jobs((size_t)n_pairs);
            for (int64_t j = 0; j < n_pairs; ++j) {
                jobs[j].read = 0;
                jobs[j].job.target = ref_of(t_off[j], t_off[j + 1]);
                jobs[j].job.query = ref_of(tt + q_off[j], tt + q_off[j + 1]);
            }
            be.fill(b, eqx != 0, jobs);
            int64_t co = 0;
            for (int64_t j = 0; j < n_pairs; ++j) {
                out0[j] = (int64_t)jobs[j].cigar.size();
                if (cigar && !jobs[j].cigar.empty()) memcpy(cigar + co, jobs[j].cigar.data(), jobs[j].cigar.size() * 4);
                co += (t_off[j + 1] - t_off[j]) + (q_off[j + 1] - q_off[j]) + 2;
            }
        }
    } catch (const std::exception &e) {
        c->err = e.what();
        return VM_ERR_CUDA;
    }
    return VM_OK;
}

int vm_reads_upload(vm_ctx *c, vm_index_handle *h, int64_t n_reads, const char *seqs, const int64_t *seq_off)
{
    if (!c) return VM_ERR_ARG;
    if (!h || n_reads < 0 || !seq_off || (n_reads > 0 && !seqs)) { c->err = "bad argument"; return VM_ERR_ARG; }
    cudaSetDevice(c->device);
    try {
        if (!c->backend) {
            c->backend = new CudaBackend(c, h);
            c->backend_free = [](void *p) { delete (CudaBackend *)p; };
        }
        CudaBackend &be = *(CudaBackend *)c->backend;
        be.reads_resident = false;
        ReadBatch b;
        b.n = n_reads; b.seq = seqs; b.off = seq_off;
        be.upload_reads(b);
    } catch (const std::exception &e) {
        c->err = e.what();
        return VM_ERR_CUDA;
    }
    return VM_OK;
}

int vm_align_resident(vm_ctx *c, vm_index_handle *h, const vm_align_params *p, int64_t n_reads, const char *seqs,
                      const int64_t *seq_off, vm_result **out)
{
    return vm_align_impl(c, h, p, n_reads, seqs, seq_off, 1, out);
}

int64_t vm_result_num_records(vm_result *r) { return r ? (int64_t)r->recs.size() : 0; }
int64_t vm_result_num_cigar_ops(vm_result *r) { return r ? (int64_t)r->cigar.size() : 0; }
const int64_t *vm_result_read_offsets(vm_result *r) { return r ? r->rec_off.data() : nullptr; }
const vm_record *vm_result_records(vm_result *r) { return r ? r->recs.data() : nullptr; }
const uint32_t *vm_result_cigar(vm_result *r) { return r ? r->cigar.data() : nullptr; }
const char *vm_result_stage_times(vm_result *r) { return r ? r->stage_text.c_str() : ""; }
void vm_result_free(vm_result *r) { delete r; }

} // extern "C"
