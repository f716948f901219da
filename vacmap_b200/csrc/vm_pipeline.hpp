// Batch driver of the per-read alignment pipeline (host C++17).
//
// Mirrors get_readmap_DP_test (mammap_clrnano.py:24023-24084) for a whole batch of reads:
// the hot loops run as batched kernel launches behind `Backend`; everything between them is
// the host glue of vm_glue.hpp, run in parallel over the reads of the batch.  The driver is
// backend-agnostic so that the glue can be unit-tested on a CPU-only box with a test backend
// (tests/gluetest); the product only ever instantiates the CUDA backend (vm_backend_cuda.cu).
#pragma once
#include "vm_glue.hpp"
#include "vm_dgrun.hpp"
#include "vm_hostpool.hpp"
#include <atomic>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <cmath>
#include <cstring>
#include <ctime>
#include <chrono>
#include <functional>
#include <thread>

namespace vmp {

using vmg::Anc;
using vmg::Anc32;
using vmg::Path;

struct ReadBatch {
    int64_t n = 0;
    const char *seq = nullptr;     // upper-case bases, all reads concatenated
    const int64_t *off = nullptr;  // [n+1]
    int64_t len(int64_t r) const { return off[r + 1] - off[r]; }
};

// Result of a chaining stage for a batch.  Read r owns slots [start[r], start[r] + cnt[r]) of the
// flat arrays, which live in backend-owned (pinned) memory until the next call of the same stage;
// the *_store vectors are optional owning storage for test backends.
// where a read's device-extracted chains sit in ChainOut::x_* (layout of VmExtractRec, vm_extract.cuh)
struct ExtractRec {
    long long anc_off, meta_off;
    int32_t n_anc, n_chains;
};

// where a read's device-rebuilt sub-alignments sit in ChainOut::al_* (layout of VmRebuildRec, vm_extract.cuh)
struct RebuildRec {
    long long anc_off, len_off;
    int32_t n_anc, n_al;
};

struct ChainOut {
    std::vector<int64_t> start;
    std::vector<int32_t> cnt;
    std::vector<int64_t> gmax;      // g_max_index per read (-1: not chained)
    std::vector<int32_t> used_fast; // per read, 1 when a heuristic _fast DP produced the result (empty: not reported)
    const Anc32 *sorted = nullptr;  // anchors after the argsort
    const double *S = nullptr;      // global stage only
    const int32_t *P = nullptr;
    const int32_t *S_arg = nullptr; // global stage only
    // Compact form, when the backend extracts the chains itself (sorted / S / P / S_arg are then not filled):
    // global stage -- the chains hit2work_1 keeps (score > 40), primary first, each in descending read order,
    // with the primary chain's S values; local stage -- the trimmed best chain in ascending read order.
    const ExtractRec *rec = nullptr;
    const Anc32 *x_anc = nullptr;
    const double *x_S = nullptr;
    const int32_t *x_len = nullptr;
    const double *x_score = nullptr;
    // Local stage only, when the backend also ran rebuild_chain_break (:23437-23484) on the path: per read its
    // colinear sub-alignments (anchors back to back in al_anc, anchors per sub-alignment in al_len).  The same
    // anchors stay on the device; EdJob::seg_off then counts in that device array (al_rec[r].anc_off + ...).
    const RebuildRec *al_rec = nullptr;
    const Anc32 *al_anc = nullptr;
    const int32_t *al_len = nullptr;
    std::vector<Anc32> sorted_store;
    std::vector<double> S_store;
    std::vector<int32_t> P_store, A_store;
    void adopt_stores()
    {
        sorted = sorted_store.data(); S = S_store.data(); P = P_store.data(); S_arg = A_store.data();
    }
};

struct GuideJobRef { int32_t read; vmg::GuideJob job; };
// band: half-width k of the Ukkonen band (-1 = none); dist is exact when <= band, else only known to be > band
// seg_off/seg_n: exact-match segments of the job inside the array handed to Backend::edit_distance (seg_n = 0:
// none); with them a backend may return in `dist` any upper bound of the distance that is <= band
// segs_on_device: seg_off / seg_n address the sub-alignment's ANCHORS in the backend's device copy (ChainOut::al_anc)
// and the backend derives the segments there
struct EdJob { int32_t read; vmg::SeqRef a, b; int64_t dist = 0; int64_t band = -1; int64_t seg_off = 0; int32_t seg_n = 0;
               bool segs_on_device = false; };
struct ExtJobRef { int32_t read; vmg::ExtJob job; };
// CIGAR ops of a fill job: cig[cig_off .. cig_off + cig_len) of the array Backend::fill hands back
struct FillJobRef { int32_t read; vmg::FillJob job; int64_t cig_off = 0; int32_t cig_len = 0; };

// Records of a chunk as flat arrays (what a backend that runs the extension stage itself hands back): recs / cigar
// point into backend-owned host memory, valid until the backend's next extend_device call.
struct FlatRecords {
    std::vector<int64_t> rec_off;          // [n_reads + 1]
    const vmd::Rec *recs = nullptr;        // == vm_record; cigar_off counts inside `cigar`
    const uint32_t *cigar = nullptr;
    int64_t n_rec = 0, n_ops = 0;
    int64_t counters[vmd::CT_COUNT] = {0, 0, 0, 0, 0, 0, 0, 0};
};

// The hot loops.  Every method processes the jobs of a whole batch; consecutive hot loops whose
// hand-over needs no host decision are fused so the data never leaves the device in between.
struct Backend {
    virtual ~Backend() {}
    // minimizer seeding + cluster filter + majority-strand flip (index.map + :21202-21217), then
    // argsort by read position + global DP (exact / fast as hit2work_1 chooses, :23570-23579)
    // `accept`: primary-chain score threshold of hit2work_1 (:23650), for backends that extract the chains
    virtual void seed_chain(const ReadBatch &b, int check_num, int kmersize, double skipcost, int maxdiff, int maxgap, double accept,
                            std::vector<char> &need_reverse, ChainOut &out) = 0;
    // local 9-mer re-seeding (anchors of all guide jobs of a read concatenated in job order), then
    // argsort by read end + local DP variant[r] in {1, 2} (+ fast fall-back); variant 0 = skip the read
    virtual void reseed_chain(const ReadBatch &b, const std::vector<char> &need_reverse, const std::vector<GuideJobRef> &jobs,
                              const std::vector<int> &variant, const std::vector<double> &skipcost, int maxdiff, int maxgap,
                              ChainOut &out) = 0;
    // a = query, b = target of get_query_target_for_cigar; segs: exact-match segments the jobs refer to
    virtual void edit_distance(const ReadBatch &b, std::vector<EdJob> &jobs, const vmg::MatchSeg *segs, size_t n_segs) = 0;
    // optional: memory the backend wants the match segments written to (page-locked staging); nullptr = none
    virtual vmg::MatchSeg *seg_staging(size_t n_segs) { (void)n_segs; return nullptr; }
    virtual void extend(const ReadBatch &b, std::vector<ExtJobRef> &jobs) = 0;
    // sets cig_off / cig_len of every job; the returned array stays valid until the next fill()
    virtual const uint32_t *fill(const ReadBatch &b, bool eqx, std::vector<FillJobRef> &jobs) = 0;
    // Optional: the whole of extend_func (:19238-19303) and the second pass (:24079-24080) for the reads `ids` that came
    // out of reseed_chain, on the backend's side (nothing of the local stage has to visit the host then).  status: per
    // read, in / out (ReadStatus).  false: not supported, the driver runs the host glue.
    virtual bool extend_device(const ReadBatch &b, const std::vector<int32_t> &ids, const std::vector<char> &need_reverse,
                               const std::vector<int32_t> &mapq, const vmg::Options &opt, std::vector<int32_t> &status, FlatRecords &out)
    {
        (void)b; (void)ids; (void)need_reverse; (void)mapq; (void)opt; (void)status; (void)out;
        return false;
    }
    virtual bool has_device_extension() const { return false; }
    // Optional: hit2work_1's bookkeeping after the DP, guide-chain selection and the re-seeding jobs (:23581-23734,
    // :28482-28582, :23090-23191) on the backend's side, from the chains seed_chain extracted there.  fo[r]: what the
    // driver still needs of read r (status, number of guide chains, MAPQ inputs).  The jobs stay with the backend;
    // reseed_chain_front runs them.  false: not supported, the driver runs the host glue.
    virtual bool has_device_front() const { return false; }
    virtual bool front_device(const ReadBatch &b, const std::vector<char> &need_reverse, const ChainOut &g, int max_guides,
                              std::vector<vmd::FrontOut> &fo)
    {
        (void)b; (void)need_reverse; (void)g; (void)max_guides; (void)fo;
        return false;
    }
    virtual void reseed_chain_front(const ReadBatch &b, const std::vector<char> &need_reverse, const std::vector<vmd::FrontOut> &fo,
                                    const std::vector<int> &variant, const std::vector<double> &skipcost, int maxdiff, int maxgap,
                                    ChainOut &out)
    {
        (void)b; (void)need_reverse; (void)fo; (void)variant; (void)skipcost; (void)maxdiff; (void)maxgap; (void)out;
    }
};

// Thread time spent in the parts of the host glue, summed over the pool's threads ("t_<name>" stage entries):
// where the host cores go inside a phase.
enum SubPart { SP_H2W = 0, SP_GUIDES, SP_GUIDEJOB, SP_REBUILD, SP_QT, SP_SEGS, SP_MERGE, SP_FIXINV, SP_SPLIT, SP_FILLJOBS, SP_CIGCAT,
               SP_RECORDS, SP_TRACE, SP_COUNT };
static const char *const kSubPartName[SP_COUNT] = {"t_hit2work", "t_select_guides", "t_guide_job", "t_rebuild_break", "t_query_target",
                                                   "t_match_segments", "t_merge_conjacent", "t_fix_simple_inv", "t_split_alignment",
                                                   "t_fill_jobs", "t_cigar_concat", "t_make_records", "t_traceback_copy"};
struct SubTimes {
    std::atomic<int64_t> ns[SP_COUNT];
    SubTimes() { for (auto &x : ns) x.store(0); }
};
struct SubScope {
    SubTimes &st; int part; std::chrono::steady_clock::time_point t0;
    SubScope(SubTimes &s, int p) : st(s), part(p), t0(std::chrono::steady_clock::now()) {}
    ~SubScope() { st.ns[part].fetch_add(std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count(), std::memory_order_relaxed); }
};

struct ReadState {
    bool alive = false;
    bool dropped = false;           // the reference raises on this read (ReadDropped)
    bool need_reverse = false;
    int mapq = 0;
    std::vector<Path> guides;
    size_t n_guides_used = 0;
    Path asc;                       // local chain, ascending read order
    // extend_func state
    vmg::AlnList al;
    size_t n0 = 0;
    bool filtered = false;
    bool nofilter = false;
    std::vector<vmg::ExtJob> ext;
    vmg::AlnList kept;
    std::vector<vmg::FillJob> fills;
    std::vector<vmg::Record> recs;
};

// Why a read has the records it has (vm_result_read_status): the reference's "emit nothing" cases, told apart.
enum ReadStatus {
    RS_OK = 0,            // >= 1 record
    RS_FEW_ANCHORS = 1,   // <= 2 seed anchors after the cluster filter (decode_hit :23986-23988)
    RS_LOW_SCORE = 2,     // best global chain not above the mode's threshold (hit2work_1 :23650, :23711-23734)
    RS_SHORT_LOCAL = 3,   // local chain of <= 1 anchor (:24067-24068)
    RS_DROPPED = 4,       // the reference raises inside the read and its worker swallows it (:24116-24125):
                          // empty rebuild_chain_break, "Failed to compute CIGAR" (:21559-21569), Cigar length check (:20779-20786), ...
    RS_NO_RECORDS = 5,    // extend_func returned no record (:24075-24076)
    RS_FAILED = 6         // the library could not process the read (size limits); the reference has no counterpart
};

// Branches of the reference's per-read driver a batch went through (how often), for the parity tests' coverage claims
enum BranchCounter { BC_FAST_GLOBAL = 0, BC_MISMATCH_DP, BC_FAST_LOCAL, BC_DROP_MISPLACED, BC_MERGE_CONJACENT, BC_FIX_SIMPLE_INV,
                     BC_SECOND_PASS, BC_COUNT };
static const char *const kBranchName[BC_COUNT] = {"c_fast_global", "c_mismatch_dp", "c_fast_local", "c_drop_misplaced",
                                                  "c_merge_conjacent", "c_fix_simple_inv", "c_second_pass"};

struct BatchResult {
    bool flat = false;                               // true: the records are in `fr` (flat arrays), `records` is unused
    FlatRecords fr;
    std::vector<std::vector<vmg::Record>> records;   // per read, in the reference's emission order
    std::vector<int32_t> status;                     // ReadStatus per read
    int64_t branch[BC_COUNT] = {0, 0, 0, 0, 0, 0, 0};
};

// map the oriented sequences of the per-read driver onto the stored orientations:
// after need_reverse the reference swaps testseq / rc_testseq (:24063-24065)
static inline void orient(vmg::SeqRef &s, bool need_reverse)
{
    if (need_reverse && s.src != 0) s.src = 3 - s.src;
}

class Driver {
public:
    Driver(Backend &be, const vmg::Contigs &ctg, const vmg::Options &opt, int kmersize, int threads)
        : be_(be), ctg_(ctg), opt_(opt), k_(kmersize), threads_(threads) {}

    // optional wall-clock accounting of the host glue phases (name, milliseconds)
    std::function<void(const char *, double)> on_time;
    static double process_cpu_ms()
    {
        timespec ts;
        clock_gettime(CLOCK_PROCESS_CPUTIME_ID, &ts);
        return 1e3 * (double)ts.tv_sec + 1e-6 * (double)ts.tv_nsec;
    }
    std::function<void(const char *, double)> on_cpu;     // (phase, process CPU milliseconds burnt during it)
    struct Phase {
        Driver *d; const char *name; std::chrono::steady_clock::time_point t0; double c0;
        Phase(Driver *d_, const char *n) : d(d_), name(n), t0(std::chrono::steady_clock::now()), c0(process_cpu_ms()) {}
        ~Phase()
        {
            if (d->on_time) d->on_time(name, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
            if (d->on_cpu) d->on_cpu(name, process_cpu_ms() - c0);
        }
    };

    void align_batch(const ReadBatch &b, BatchResult &res)
    {
        const int64_t n = b.n;
        res.records.assign((size_t)n, {});
        res.status.assign((size_t)n, RS_OK);
        for (auto &x : bc_) x.store(0);
        std::vector<ReadState> st((size_t)n);
        std::vector<int64_t> read_len((size_t)n);
        for (int64_t r = 0; r < n; ++r) read_len[r] = b.len(r);

        // ---- 1-2. seeding + global chaining (fused on the device) ----
        std::vector<char> need_rev;
        ChainOut g;
        be_.seed_chain(b, opt_.check_num, k_, opt_.global_skipcost, opt_.global_maxdiff, 1000, opt_.mode.accept, need_rev, g);

        // ---- 3. hit2work bookkeeping + guide selection ----
        std::vector<int> variant((size_t)n, 0);
        std::vector<double> skip((size_t)n, opt_.local_skipcost);
        ChainOut lc;
        std::vector<vmd::FrontOut> fo;
        if (be_.has_device_front() && be_.front_device(b, need_rev, g, opt_.mode.max_guides, fo)) {
            // on the backend's side: the host only turns the per-read summary into flags
            Phase ph3(this, "g_hit2work");
            for (int64_t r = 0; r < n; ++r) {
                if (g.cnt[r] <= 2) { res.status[r] = RS_FEW_ANCHORS; continue; }       // decode_hit :23986
                if (!g.used_fast.empty() && g.used_fast[r]) bc_[BC_FAST_GLOBAL].fetch_add(1, std::memory_order_relaxed);
                const vmd::FrontOut &f = fo[(size_t)r];
                if (f.status != vmd::ST_OK) { res.status[r] = RS_LOW_SCORE; continue; }
                ReadState &s = st[r];
                s.alive = true;
                s.need_reverse = need_rev[r] != 0;
                // min(int(40*(1-f2/f1)*min(1, m/10)*np.log(f1)), 60)  (:23704); numba lowers np.log to libm log
                double v = 40 * (1 - f.f2 / f.f1);
                v = v * std::min(1.0, f.m / 10);
                v = v * std::log(f.f1);
                s.mapq = (int)std::min<int64_t>((int64_t)v, 60);
                if (f.n_guides > 1) {
                    variant[r] = 2;
                    bc_[BC_MISMATCH_DP].fetch_add(1, std::memory_order_relaxed);
                    if (opt_.mode.clamp40) skip[r] = std::min(skip[r], 40.0);
                } else variant[r] = 1;
            }
            be_.reseed_chain_front(b, need_rev, fo, variant, skip, opt_.local_maxdiff, opt_.mode.local_maxgap, lc);
        } else {
        std::vector<std::vector<GuideJobRef>> gjobs((size_t)n);
        Phase *ph = new Phase(this, "g_hit2work");
        parallel_for(n, threads_, [&](int64_t r) {
            const int64_t m = g.cnt[r];
            if (m <= 2) { res.status[r] = RS_FEW_ANCHORS; return; }     // decode_hit :23986 -- <= 2 anchors: unmapped
            if (!g.used_fast.empty() && g.used_fast[r]) bc_[BC_FAST_GLOBAL].fetch_add(1, std::memory_order_relaxed);
            const int64_t o = g.start[r];
            vmg::GlobalResult gr;
            {
                SubScope sc(sub_, SP_H2W);
                if (g.rec) {
                    const ExtractRec &x = g.rec[r];
                    vmg::hit2work_extracted(g.x_anc + x.anc_off, g.x_S + x.anc_off, g.x_len + x.meta_off, g.x_score + x.meta_off,
                                            x.n_chains, read_len[r], gr);
                } else vmg::hit2work(g.sorted + o, g.S + o, g.P + o, g.S_arg + o, m, g.gmax[r], read_len[r], opt_.mode.accept, gr);
            }
            if (!gr.ok) { res.status[r] = RS_LOW_SCORE; return; }
            ReadState &s = st[r];
            s.alive = true;
            s.need_reverse = need_rev[r] != 0;
            s.mapq = gr.mapq;
            s.guides.swap(gr.guides);
            {
                SubScope sc(sub_, SP_GUIDES);
                s.n_guides_used = vmg::select_guides(s.guides, opt_.mode);
            }
            SubScope sc(sub_, SP_GUIDEJOB);
            for (size_t gi = 0; gi < s.n_guides_used; ++gi) {
                GuideJobRef j;
                j.read = (int32_t)r;
                vmg::make_guide_job(s.guides[gi], read_len[r], 9, ctg_, j.job);
                gjobs[r].push_back(std::move(j));
            }
        });
        std::vector<GuideJobRef> all_gjobs;
        std::vector<int64_t> gj_start;
        parallel_concat(gjobs, threads_, all_gjobs, gj_start);
        for (int64_t r = 0; r < n; ++r) {
            if (!st[r].alive) continue;
            if (st[r].guides.size() > 1) {
                variant[r] = 2;
                bc_[BC_MISMATCH_DP].fetch_add(1, std::memory_order_relaxed);
                if (opt_.mode.clamp40) skip[r] = std::min(skip[r], 40.0);
            } else variant[r] = 1;
        }
        gjobs.clear();
        delete ph;

        // ---- 4-5. local re-seeding + local chaining (fused on the device) ----
        be_.reseed_chain(b, need_rev, all_gjobs, variant, skip, opt_.local_maxdiff, opt_.mode.local_maxgap, lc);
        }
        lc_ = &lc;
        if (be_.has_device_extension()) {
            // ---- 6'. extend_func and the second pass with their glue on the device: the host only launches ----
            std::vector<int32_t> ids, mapq((size_t)n, 0);
            for (int64_t r = 0; r < n; ++r)
                if (st[r].alive) { ids.push_back((int32_t)r); mapq[(size_t)r] = st[r].mapq; }
            for (int64_t r = 0; r < n; ++r)
                if (!lc.used_fast.empty() && lc.used_fast[r]) bc_[BC_FAST_LOCAL].fetch_add(1, std::memory_order_relaxed);
            if (be_.extend_device(b, ids, need_rev, mapq, opt_, res.status, res.fr)) {
                res.flat = true;
                res.records.clear();
                lc_ = nullptr;
                bc_[BC_DROP_MISPLACED].fetch_add(res.fr.counters[vmd::CT_DROP_MISPLACED]);
                bc_[BC_MERGE_CONJACENT].fetch_add(res.fr.counters[vmd::CT_MERGE_CONJACENT]);
                bc_[BC_FIX_SIMPLE_INV].fetch_add(res.fr.counters[vmd::CT_FIX_SIMPLE_INV]);
                bc_[BC_SECOND_PASS].fetch_add(res.fr.counters[vmd::CT_SECOND_PASS]);
                if (on_time)
                    for (int p = 0; p < SP_COUNT; ++p) on_time(kSubPartName[p], 1e-6 * (double)sub_.ns[p].exchange(0));
                Phase p2(this, "g_finish");
                parallel_for(n, threads_, [&](int64_t r) { st[r] = ReadState(); }, 64);
                st.clear();
                for (int i = 0; i < BC_COUNT; ++i) res.branch[i] = bc_[i].load();
                return;
            }
        }

        // ---- 6. traceback, then extend_func as a staged state machine ----
        Phase *ph = new Phase(this, "g_traceback");
        parallel_for(n, threads_, [&](int64_t r) {
            ReadState &s = st[r];
            if (!s.alive) return;
            if (!lc.used_fast.empty() && lc.used_fast[r]) bc_[BC_FAST_LOCAL].fetch_add(1, std::memory_order_relaxed);
            if (lc.cnt[r] == 0) { s.alive = false; res.status[r] = RS_DROPPED; return; }   // np.array([]) indexing raises in the reference
            const int64_t o = lc.start[r];
            SubScope sc(sub_, SP_TRACE);
            if (lc.rec) {
                // the path was extracted on the device; it is widened where it is consumed (rebuild_chain_break)
                if (lc.rec[r].n_anc <= 1) s.alive = false;
            } else {
                vmg::local_traceback(lc.sorted + o, lc.P + o, lc.gmax[r], s.asc);
                if (s.asc.size() <= 1) s.alive = false;
            }
            if (!s.alive) res.status[r] = RS_SHORT_LOCAL;
            s.nofilter = opt_.nodiscard;
        });
        delete ph;
        std::vector<int64_t> todo;
        for (int64_t r = 0; r < n; ++r)
            if (st[r].alive) todo.push_back(r);
        first_pass_ = true;
        extend_pass(b, read_len, st, todo);
        first_pass_ = false;     // the rare second pass rebuilds on the host (the device copy has served its jobs)
        // second pass (:24079-24080): paired large indels after a filtered sub-alignment
        std::vector<int64_t> again;
        for (int64_t r : todo) {
            ReadState &s = st[r];
            if (s.alive && !s.recs.empty() && !opt_.nodiscard && s.filtered && vmg::paired_indel(s.recs)) {
                s.nofilter = true;
                again.push_back(r);
                bc_[BC_SECOND_PASS].fetch_add(1, std::memory_order_relaxed);
            }
        }
        if (!again.empty()) extend_pass(b, read_len, st, again);
        lc_ = nullptr;
        if (on_time)
            for (int p = 0; p < SP_COUNT; ++p) on_time(kSubPartName[p], 1e-6 * (double)sub_.ns[p].exchange(0));
        {
            Phase p2(this, "g_finish");
            parallel_for(n, threads_, [&](int64_t r) {
                if (st[r].alive) res.records[r].swap(st[r].recs);
                else if (res.status[r] == RS_OK) res.status[r] = st[r].dropped ? RS_DROPPED : RS_NO_RECORDS;
                st[r] = ReadState();     // the per-read state is torn down by the pool, not serially
            }, 64);
            st.clear();
            for (int i = 0; i < BC_COUNT; ++i) res.branch[i] = bc_[i].load();
        }
    }

private:
    // one extend_func call (:19238-19303) for every read in `ids`
    void extend_pass(const ReadBatch &b, const std::vector<int64_t> &read_len, std::vector<ReadState> &st,
                     const std::vector<int64_t> &ids)
    {
        const int64_t m = (int64_t)ids.size();
        // a. rebuild_chain_break + divergence filter jobs
        Phase *ph = new Phase(this, "g_ext_rebuild");
        std::vector<std::vector<EdJob>> edj((size_t)m);
        std::vector<std::vector<vmg::MatchSeg>> segj((size_t)m);
        parallel_for(m, threads_, [&](int64_t t) {
            const int64_t r = ids[t];
            ReadState &s = st[r];
            s.recs.clear();
            s.filtered = false;
            try {
                if (first_pass_ && lc_ && lc_->al_rec) {
                    // sub-alignments rebuilt on the device: widen them, the match segments are derived there too
                    SubScope sc(sub_, SP_TRACE);
                    const RebuildRec &x = lc_->al_rec[r];
                    if (x.n_al == 0) throw vmg::ReadDropped("rebuild_chain_break: empty");
                    s.al.assign((size_t)x.n_al, Path());
                    long long ao = x.anc_off;
                    for (int32_t i = 0; i < x.n_al; ++i) {
                        const int32_t len = lc_->al_len[x.len_off + i];
                        Path &p = s.al[(size_t)i];
                        p.resize((size_t)len);
                        for (int32_t k = 0; k < len; ++k) p[(size_t)k] = vmg::widen(lc_->al_anc[ao + k]);
                        EdJob j;
                        j.read = (int32_t)r;
                        vmg::query_target(p.front(), p.back(), read_len[r], ctg_, j.b, j.a);
                        if (std::min(j.a.len(), j.b.len()) == 0) throw vmg::ReadDropped("division by zero");
                        j.seg_off = ao;
                        j.seg_n = len;
                        j.segs_on_device = true;
                        orient(j.a, s.need_reverse);
                        orient(j.b, s.need_reverse);
                        j.band = divergence_band(std::min(j.a.len(), j.b.len()));
                        edj[t].push_back(j);
                        ao += len;
                    }
                    return;
                }
                const Path *asc = &s.asc;
                thread_local Path widened;
                if (lc_ && lc_->rec) {
                    SubScope sc(sub_, SP_TRACE);
                    const ExtractRec &x = lc_->rec[r];
                    widened.resize((size_t)x.n_anc);
                    for (int32_t k = 0; k < x.n_anc; ++k) widened[(size_t)k] = vmg::widen(lc_->x_anc[x.anc_off + k]);
                    asc = &widened;
                }
                segj[t].reserve(asc->size());
                {
                    SubScope sc(sub_, SP_REBUILD);
                    vmg::rebuild_chain_break(ctg_, *asc, opt_.local_maxdiff, s.al);
                }
                for (size_t i = 0; i < s.al.size(); ++i) {
                    EdJob j;
                    j.read = (int32_t)r;
                    {
                        SubScope sc(sub_, SP_QT);
                        vmg::query_target(s.al[i].front(), s.al[i].back(), read_len[r], ctg_, j.b, j.a);
                    }
                    if (std::min(j.a.len(), j.b.len()) == 0) throw vmg::ReadDropped("division by zero");
                    j.seg_off = (int64_t)segj[t].size();
                    SubScope sc(sub_, SP_SEGS);
                    j.seg_n = vmg::match_segments(s.al[i], j.a.len(), j.b.len(), segj[t], 128) ? (int32_t)(segj[t].size() - (size_t)j.seg_off) : 0;
                    orient(j.a, s.need_reverse);
                    orient(j.b, s.need_reverse);
                    j.band = divergence_band(std::min(j.a.len(), j.b.len()));
                    edj[t].push_back(j);
                }
            } catch (const vmg::ReadDropped &) { s.alive = false; s.dropped = true; edj[t].clear(); segj[t].clear(); }
        });
        std::vector<EdJob> ed;
        std::vector<int64_t> ed_start, seg_start((size_t)m + 1, 0);
        std::vector<vmg::MatchSeg> segs_own;
        parallel_concat(edj, threads_, ed, ed_start);
        for (int64_t t = 0; t < m; ++t) seg_start[t + 1] = seg_start[t] + (int64_t)segj[t].size();
        const size_t n_segs = (size_t)seg_start[m];
        vmg::MatchSeg *segs = be_.seg_staging(n_segs);       // straight into the backend's staging when it has one
        if (!segs) { segs_own.resize(n_segs); segs = segs_own.data(); }
        parallel_for(m, threads_, [&](int64_t t) {
            if (!segj[t].empty()) memcpy(segs + seg_start[t], segj[t].data(), segj[t].size() * sizeof(vmg::MatchSeg));
            for (int64_t q = ed_start[t]; q < ed_start[t + 1]; ++q) ed[q].seg_off += seg_start[t];
        }, 64);
        delete ph;
        be_.edit_distance(b, ed, segs, n_segs);
        ph = new Phase(this, "g_ext_edges");
        parallel_for(m, threads_, [&](int64_t t) {
            ReadState &s = st[ids[t]];
            if (!s.alive) return;
            vmg::AlnList keep;
            for (int64_t q = ed_start[t]; q < ed_start[t + 1]; ++q) {
                const double ratio = (double)ed[q].dist / (double)std::min(ed[q].a.len(), ed[q].b.len());
                if (!(ratio > opt_.maxdivergence)) keep.push_back(std::move(s.al[q - ed_start[t]]));
            }
            s.al.swap(keep);
        });
        // b. edge extension, two dependency rounds
        extend_rounds(b, read_len, st, ids);
        // c. misplaced sub-alignments, second extension when something was dropped
        std::vector<int64_t> changed;
        for (int64_t t = 0; t < m; ++t) {
            ReadState &s = st[ids[t]];
            if (!s.alive) continue;
            s.n0 = s.al.size();
            if (s.al.size() > 2 && !s.nofilter) {
                size_t iloc = 0;
                while (iloc + 2 < s.al.size()) {
                    if (!vmg::drop_misplaced(s.al, iloc)) ++iloc;
                    else bc_[BC_DROP_MISPLACED].fetch_add(1, std::memory_order_relaxed);
                }
            }
            if (s.al.size() < s.n0) { s.filtered = true; changed.push_back(ids[t]); }
        }
        if (!changed.empty()) extend_rounds(b, read_len, st, changed);
        delete ph;
        ph = new Phase(this, "g_ext_split");
        // d. merge / inversion fix / fill jobs
        std::vector<std::vector<FillJobRef>> fj((size_t)m);
        parallel_for(m, threads_, [&](int64_t t) {
            const int64_t r = ids[t];
            ReadState &s = st[r];
            if (!s.alive) return;
            try {
                {
                    SubScope sc(sub_, SP_MERGE);
                    const size_t before = s.al.size();
                    vmg::merge_conjacent(s.al, ctg_);
                    if (s.al.size() != before) bc_[BC_MERGE_CONJACENT].fetch_add((int64_t)(before - s.al.size()), std::memory_order_relaxed);
                }
                if (s.al.size() > 2) {
                    SubScope sc(sub_, SP_FIXINV);
                    std::string oriented = oriented_read(b, r, s.need_reverse);
                    if (vmg::fix_simple_inv(s.al, ctg_, oriented.data(), read_len[r])) bc_[BC_FIX_SIMPLE_INV].fetch_add(1, std::memory_order_relaxed);
                }
                s.kept.assign(s.al.size(), Path());
                s.fills.clear();
                {
                    SubScope sc(sub_, SP_SPLIT);
                    for (size_t i = 0; i < s.al.size(); ++i)
                        vmg::split_alignment(s.al[i], (int)i, read_len[r], ctg_, s.kept[i], s.fills);
                }
                SubScope sc(sub_, SP_FILLJOBS);
                for (vmg::FillJob &f : s.fills) {
                    FillJobRef j;
                    j.read = (int32_t)r;
                    j.job = f;
                    orient(j.job.target, s.need_reverse);
                    orient(j.job.query, s.need_reverse);
                    fj[t].push_back(std::move(j));
                }
            } catch (const vmg::ReadDropped &) { s.alive = false; s.dropped = true; fj[t].clear(); }
        });
        std::vector<FillJobRef> fills;
        std::vector<int64_t> f_start;
        parallel_concat(fj, threads_, fills, f_start);
        delete ph;
        const uint32_t *cig_ops = be_.fill(b, opt_.eqx, fills);
        Phase p3(this, "g_ext_records");
        // e. records
        parallel_for(m, threads_, [&](int64_t t) {
            const int64_t r = ids[t];
            ReadState &s = st[r];
            if (!s.alive) return;
            std::vector<std::vector<vmg::OpSpan>> cig(s.al.size());
            {
                SubScope sc(sub_, SP_CIGCAT);
                for (int64_t q = f_start[t]; q < f_start[t + 1]; ++q)
                    cig[fills[q].job.aln].push_back(vmg::OpSpan{cig_ops + fills[q].cig_off, fills[q].cig_len});
            }
            SubScope sc(sub_, SP_RECORDS);
            try {
                vmg::make_records(s.kept, cig, s.mapq, read_len[r], ctg_, s.need_reverse, opt_.hardclip, s.recs);
            } catch (const vmg::ReadDropped &) { s.alive = false; s.dropped = true; s.recs.clear(); }
            if (s.recs.empty()) s.alive = false;
        });
    }

    void extend_rounds(const ReadBatch &b, const std::vector<int64_t> &read_len, std::vector<ReadState> &st,
                       const std::vector<int64_t> &ids)
    {
        const int64_t m = (int64_t)ids.size();
        for (int round = 0; round < 2; ++round) {
            parallel_for(m, threads_, [&](int64_t t) {
                ReadState &s = st[ids[t]];
                s.ext.clear();
                if (!s.alive || s.al.empty()) return;
                vmg::extend_prepare(round, read_len[ids[t]], s.al, ctg_, s.ext);
            });
            std::vector<ExtJobRef> jobs;
            std::vector<int64_t> start((size_t)m + 1, 0);
            for (int64_t t = 0; t < m; ++t) {
                start[t] = (int64_t)jobs.size();
                ReadState &s = st[ids[t]];
                for (vmg::ExtJob &e : s.ext) {
                    ExtJobRef j;
                    j.read = (int32_t)ids[t];
                    j.job = e;
                    orient(j.job.target, s.need_reverse);
                    orient(j.job.query, s.need_reverse);
                    jobs.push_back(j);
                }
            }
            start[m] = (int64_t)jobs.size();
            if (jobs.empty()) continue;
            be_.extend(b, jobs);
            parallel_for(m, threads_, [&](int64_t t) {
                ReadState &s = st[ids[t]];
                if (!s.alive) return;
                for (int64_t q = start[t]; q < start[t + 1]; ++q) {
                    s.ext[q - start[t]].q_e = jobs[q].job.q_e;
                    s.ext[q - start[t]].t_e = jobs[q].job.t_e;
                }
                vmg::extend_apply(s.al, s.ext);
            });
        }
    }

    // largest distance d that still passes the divergence filter `d / minlen > maxdivergence` (:19252-19254):
    // the filter only needs the exact distance up to there, so the kernel computes inside that band
    int64_t divergence_band(int64_t minlen) const
    {
        if (!(opt_.maxdivergence * (double)minlen < 1e9)) return -1;
        int64_t d = (int64_t)std::floor(opt_.maxdivergence * (double)minlen);
        while (d > 0 && (double)d / (double)minlen > opt_.maxdivergence) --d;
        while (!((double)(d + 1) / (double)minlen > opt_.maxdivergence)) ++d;
        return std::max<int64_t>(d, 0);
    }

    static std::string oriented_read(const ReadBatch &b, int64_t r, bool need_reverse)
    {
        std::string s(b.seq + b.off[r], (size_t)b.len(r));
        if (!need_reverse) return s;
        std::string rc(s.size(), 'N');
        for (size_t i = 0; i < s.size(); ++i) {
            const char c = s[s.size() - 1 - i];
            rc[i] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N';
        }
        return rc;
    }

    SubTimes sub_;
    std::atomic<int64_t> bc_[BC_COUNT];
    bool first_pass_ = false;
    const ChainOut *lc_ = nullptr;     // local-stage result of the batch in flight (valid during align_batch)
    Backend &be_;
    const vmg::Contigs &ctg_;
    vmg::Options opt_;
    int k_;
    int threads_;
};

} // namespace vmp
