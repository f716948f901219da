// The extension stage of a batch with its glue on the device: everything extend_func does for every read of a
// chunk (mammap_clrnano.py:19238-19303, plus the second pass :24079-24080) as a fixed sequence of launches.  The
// host only sizes buffers from a handful of scalars; no per-anchor or per-job work is left on it.
//
// Written against an execution policy `Exec` so that the very same sequence and the very same per-read functions
// (vm_dglue.hpp) run (a) as CUDA kernels in the product (vm_dglue.cu: CudaExec) and (b) as host loops over the
// oracle's C natives in the CPU test harness (tests/gluetest: OracleExec).
//
// Exec provides
//   Buf                                   growable arena in the executor's memory space; release(Buf &)
//   T *ensure<T>(Buf &, size_t n)         (contents undefined after growth)
//   HostBuf, T *host<T>(HostBuf &, n)     host memory the executor copies to fastest (page-locked for CUDA); release_host
//   zero(p, bytes), to_host(dst, src, bytes), to_exec(dst, src, bytes), sync()
//   per_item(n, functor)                  functor(int64_t t) once per item (one thread each)
//   per_item_warp(n, functor)             functor(int64_t t, Ext &ext) once per item by a whole warp (all lanes
//                                         run it redundantly; `ext` is the warp-collective z-drop extension)
//   add(counter *, value) inside functors through vmd::Atomic (see below)
//   ed_bounds(jobs, n, al_anc)            distance (or an upper bound) of every divergence-filter job -> result0
//   ed_exact(jobs, open_ids, n_open)      exact banded distance of the jobs the bound left open
//   fill(jobs, n, scratch_words, eqx, results, &ops)   global fill of every job: results[j] = (offset, length) in ops
#pragma once
#include "vm_dglue.hpp"
#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <vector>

namespace vmd {

#if defined(__CUDA_ARCH__)
#define VMD_ATOMIC_ADD(p, v) atomicAdd((unsigned long long *)(p), (unsigned long long)(v))
#else
#define VMD_ATOMIC_ADD(p, v) __atomic_fetch_add((unsigned long long *)(p), (unsigned long long)(v), __ATOMIC_RELAXED)
#endif

struct BackParams {
    double maxdivergence;
    int32_t eqx, hardclip, nodiscard;
};

// what the earlier stages hand over; every pointer lives in the executor's memory space
struct BackInput {
    int64_t n_reads;
    const int64_t *read_off;          // [n_reads + 1]
    const uint8_t *reads_fwd, *reads_rc, *ref;
    Ctg ctg;
    const int32_t *need_reverse;      // per read
    const int32_t *mapq;              // per read
    const int32_t *local_cnt;         // anchors that went into the read's local DP
    const ExtractRec *xrec;           // extracted local chain
    const RebuildRec *rrec;           // rebuilt sub-alignments
    const A32 *al_anc;
    const int32_t *al_len;
    int64_t NA, NT;                   // total sub-alignments / anchors in them (host-known scalars)
    int64_t total_bases;              // of the chunk's reads
};

// one pass of extend_func over a set of reads: device state
template <typename Exec>
struct PassState {
    typename Exec::Buf b_sub, b_jobs, b_alive, b_filtered, b_fa, b_seg, b_first, b_last, b_dup, b_fin, b_nfin, b_fill, b_res, b_out,
        b_open;
    Sub *sub = nullptr;
    Job *jobs = nullptr;
    int32_t *alive = nullptr, *filtered = nullptr, *nfin = nullptr, *open = nullptr;
    Anc *fa = nullptr, *first = nullptr, *last = nullptr;
    Seg *seg = nullptr;
    uint8_t *dup = nullptr;
    Fin *fin = nullptr;
    Job *fill = nullptr;
    U2 *res = nullptr;
    const uint32_t *ops = nullptr;
    ReadOut *out = nullptr;
    int64_t n_jobs = 0;
};

// ---- functors (one per launch) ----
struct FInit {
    BackInput in; BackParams p;
    const int32_t *ids; Sub *sub; Job *jobs; int32_t *alive, *status;
    VM_HD void operator()(int64_t t) const
    {
        const int32_t r = ids[t];
        ReadCtx rc;
        rc.read = r; rc.L = in.read_off[r + 1] - in.read_off[r]; rc.need_reverse = in.need_reverse[r] != 0; rc.ctg = in.ctg;
        const int st = init_read(rc, in.local_cnt[r], in.xrec[r], in.rrec[r], in.al_anc, in.al_len, p.maxdivergence, sub, jobs);
        alive[r] = st == ST_OK;
        status[r] = st;
    }
};

struct FOpen {      // jobs the upper bound leaves open (bound > band)
    const Job *jobs; int32_t *open; unsigned long long *counters;
    VM_HD void operator()(int64_t j) const
    {
        const Job &J = jobs[j];
        if (J.t.len > 0 && J.q.len > 0 && J.result0 > J.out_off) open[VMD_ATOMIC_ADD(counters + CT_OPEN_ED, 1)] = (int32_t)j;
    }
};

struct FExtend {
    BackInput in; BackParams p;
    const int32_t *ids; Sub *sub; const Job *jobs; const int32_t *alive; int32_t *filtered; const int32_t *nofilter;
    unsigned long long *counters;
    template <typename Ext>
    VM_HD void operator()(int64_t t, Ext &ext, bool leader) const
    {
        const int32_t r = ids[t];
        if (!alive[r]) return;
        ReadCtx rc;
        rc.read = r; rc.L = in.read_off[r + 1] - in.read_off[r]; rc.need_reverse = in.need_reverse[r] != 0; rc.ctg = in.ctg;
        const RebuildRec rr = in.rrec[r];
        bool f = false;
        int nd = 0;
        ext.read = r;
        extend_read(rc, sub + rr.len_off, rr.n_al, jobs + rr.len_off, in.al_anc, p.maxdivergence, p.nodiscard || (nofilter && nofilter[r]),
                    ext, &f, &nd);
        if (leader) {
            filtered[r] = f ? 1 : 0;
            if (nd) VMD_ATOMIC_ADD(counters + CT_DROP_MISPLACED, nd);
        }
    }
};

struct FFinalize {
    BackInput in; BackParams p;
    const int32_t *ids; const Sub *sub; int32_t *alive, *status, *nfin;
    Anc *fa, *first, *last; Seg *seg; uint8_t *dup; Fin *fin; Job *fill;
    unsigned long long *counters;
    int64_t job_cap;
    struct Alloc {
        unsigned long long *counters;
        int64_t job_cap;
        // false: the job arena is full (cannot happen within its bound; the read is then reported as failed)
        VM_HD bool operator()(int64_t nj, int64_t words, int64_t &job_base, int64_t &scratch_base) const
        {
            job_base = (int64_t)VMD_ATOMIC_ADD(counters + CT_N_JOBS, nj);
            scratch_base = (int64_t)VMD_ATOMIC_ADD(counters + CT_CIG_SCRATCH, words);
            return job_base + nj <= job_cap;
        }
    };
    VM_HD void operator()(int64_t t) const
    {
        const int32_t r = ids[t];
        nfin[r] = 0;
        if (!alive[r]) return;
        ReadCtx rc;
        rc.read = r; rc.L = in.read_off[r + 1] - in.read_off[r]; rc.need_reverse = in.need_reverse[r] != 0; rc.ctg = in.ctg;
        const RebuildRec rr = in.rrec[r];
        // the oriented read: after need_reverse the reference's testseq is the reverse complement
        const uint8_t *seq = (rc.need_reverse ? in.reads_rc : in.reads_fwd) + in.read_off[r];
        Alloc al{counters, job_cap};
        // room for the read's final anchor lists (its anchors + two spare slots per sub-alignment): own bump allocator --
        // anc_off and len_off come from two independent counters of the rebuild kernel, so their sum is not a layout
        const int64_t fa_off = (int64_t)VMD_ATOMIC_ADD(counters + CT_FA, (int64_t)rr.n_anc + 2 * (int64_t)rr.n_al);
        const FinalizeOut o = finalize_read(rc, sub + rr.len_off, rr.n_al, in.al_anc, fa + fa_off, seg + rr.len_off,
                                            first + rr.len_off, last + rr.len_off, dup + rr.len_off, in.ref, seq, fin + rr.len_off, fill, al);
        if (o.n_merged) VMD_ATOMIC_ADD(counters + CT_MERGE_CONJACENT, o.n_merged);
        if (o.n_fixinv) VMD_ATOMIC_ADD(counters + CT_FIX_SIMPLE_INV, 1);
        if (o.status != ST_OK) { alive[r] = 0; status[r] = o.status; return; }
        nfin[r] = o.n_fin;
    }
};

struct FCount {
    BackInput in; BackParams p;
    const int32_t *ids; int32_t *alive, *status; const int32_t *nfin, *filtered, *nofilter; const Fin *fin; const U2 *res; const uint32_t *ops;
    ReadOut *out;
    VM_HD void operator()(int64_t t) const
    {
        const int32_t r = ids[t];
        ReadOut o;
        o.n_rec = 0; o.second = 0; o.n_ops = 0;
        const RebuildRec rr = in.rrec[r];
        o.fin_lo = rr.len_off;
        if (alive[r]) {
            ReadCtx rc;
            rc.read = r; rc.L = in.read_off[r + 1] - in.read_off[r]; rc.need_reverse = in.need_reverse[r] != 0; rc.ctg = in.ctg;
            int64_t n_ops = 0;
            bool paired = false;
            const int st = count_read(rc, fin + rr.len_off, nfin[r], res, ops, p.hardclip != 0, &n_ops, &paired);
            if (st != ST_OK) { alive[r] = 0; status[r] = st; }
            else {
                o.n_rec = nfin[r];
                o.n_ops = n_ops;
                // second pass (:24079-24080): something was filtered and the CIGARs hold a pair of similar large indels
                const bool nf = p.nodiscard || (nofilter && nofilter[r]);
                o.second = (!nf && filtered[r] && paired) ? 1 : 0;
            }
        }
        out[r] = o;
    }
};

struct FWrite {
    BackInput in; BackParams p;
    int64_t n;
    const ReadOut *out1, *out2; const Fin *fin1, *fin2; const U2 *res1, *res2; const uint32_t *ops1, *ops2;
    const int64_t *rec_off, *cig_off;      // per read, exclusive prefix sums
    Rec *recs; uint32_t *cig;
    VM_HD void run(int64_t r, int lane, int nl) const
    {
        const bool second = out2 && out1[r].second;
        const ReadOut o = second ? out2[r] : out1[r];
        if (o.n_rec <= 0) return;
        const Fin *fin = (second ? fin2 : fin1) + o.fin_lo;
        const U2 *res = second ? res2 : res1;
        const uint32_t *ops = second ? ops2 : ops1;
        ReadCtx rc;
        rc.read = (int32_t)r; rc.L = in.read_off[r + 1] - in.read_off[r]; rc.need_reverse = in.need_reverse[r] != 0; rc.ctg = in.ctg;
        int64_t co = cig_off[r];
        for (int k = 0; k < o.n_rec; ++k) {
            const int i = rc.need_reverse ? o.n_rec - 1 - k : k;        // the reference reverses the list (:20836-20838)
            co += write_record(rc, fin[i], in.mapq[r], res, ops, p.hardclip != 0, recs + rec_off[r] + k, cig, co, lane, nl);
        }
    }
};

// host-side result of a chunk; recs / cigar point into the BackHalf's host buffers (valid until its next run)
struct BackResult {
    std::vector<int64_t> rec_off;     // [n_reads + 1]
    const Rec *recs = nullptr;
    const uint32_t *cigar = nullptr;
    int64_t n_rec = 0, n_ops = 0;
    std::vector<int32_t> status;      // per read (ST_*)
    int64_t counters[CT_COUNT] = {0, 0, 0, 0, 0, 0, 0, 0};
};

template <typename Exec>
class BackHalf {
public:
    explicit BackHalf(Exec &ex) : ex_(ex) {}
    ~BackHalf()
    {
        for (PassState<Exec> *P : {&P1_, &P2_}) {
            typename Exec::Buf *b[] = {&P->b_sub, &P->b_jobs, &P->b_alive, &P->b_filtered, &P->b_fa, &P->b_seg, &P->b_first, &P->b_last,
                                       &P->b_dup, &P->b_fin, &P->b_nfin, &P->b_fill, &P->b_res, &P->b_out, &P->b_open};
            for (auto *x : b) Exec::release(*x);
        }
        typename Exec::Buf *b[] = {&b_status_, &b_ids_, &b_ids2_, &b_counters_, &b_nofilter_, &b_recoff_, &b_cigoff_, &b_recs_, &b_cig_};
        for (auto *x : b) Exec::release(*x);
        Exec::release_host(h_recs_);
        Exec::release_host(h_cig_);
    }

    // ids_host: the reads that reached the extension stage; status_host: status of every read so far (ST_OK for ids)
    void run(const BackInput &in, const BackParams &p, const std::vector<int32_t> &ids_host, std::vector<int32_t> &status_host,
             BackResult &out)
    {
        const int64_t n = in.n_reads;
        int32_t *status = ex_.template ensure<int32_t>(b_status_, (size_t)n + 1);
        int32_t *ids = ex_.template ensure<int32_t>(b_ids_, ids_host.size() + 1);
        unsigned long long *counters = ex_.template ensure<unsigned long long>(b_counters_, 2 * CT_COUNT);
        ex_.to_exec(status, status_host.data(), (size_t)n * 4);
        ex_.to_exec(ids, ids_host.data(), ids_host.size() * 4);
        ex_.zero(counters, 2 * CT_COUNT * 8);
        run_pass(in, p, ids, (int64_t)ids_host.size(), nullptr, status, counters, P1_);
        // reads that ask for the second pass
        std::vector<ReadOut> ro((size_t)n);
        ex_.to_host(ro.data(), P1_.out, (size_t)n * sizeof(ReadOut));
        ex_.sync();
        std::vector<int32_t> ids2;
        for (int32_t r : ids_host)
            if (ro[(size_t)r].second) ids2.push_back(r);
        std::vector<ReadOut> ro2;
        bool have2 = false;
        if (!ids2.empty()) {
            have2 = true;
            int32_t *nofilter = ex_.template ensure<int32_t>(b_nofilter_, (size_t)n + 1);
            int32_t *ids2d = ex_.template ensure<int32_t>(b_ids2_, ids2.size() + 1);
            std::vector<int32_t> nf((size_t)n, 0);
            for (int32_t r : ids2) nf[(size_t)r] = 1;
            ex_.to_exec(nofilter, nf.data(), (size_t)n * 4);
            ex_.to_exec(ids2d, ids2.data(), ids2.size() * 4);
            ex_.fill_slot = 1;
            run_pass(in, p, ids2d, (int64_t)ids2.size(), nofilter, status, counters + CT_COUNT, P2_);
            ex_.fill_slot = 0;
            ro2.resize((size_t)n);
            ex_.to_host(ro2.data(), P2_.out, (size_t)n * sizeof(ReadOut));
            ex_.sync();
        }
        // offsets (O(reads) on the host), then one launch writes records and CIGARs in read order
        out.rec_off.assign((size_t)n + 1, 0);
        std::vector<int64_t> cig_off((size_t)n + 1, 0);
        std::vector<uint8_t> in_ids((size_t)n, 0);
        for (int32_t r : ids_host) in_ids[(size_t)r] = 1;
        for (int64_t r = 0; r < n; ++r) {
            int64_t nr = 0, no = 0;
            if (in_ids[(size_t)r]) {
                const ReadOut &o = (have2 && ro[(size_t)r].second) ? ro2[(size_t)r] : ro[(size_t)r];
                nr = o.n_rec; no = o.n_ops;
            }
            out.rec_off[(size_t)r + 1] = out.rec_off[(size_t)r] + nr;
            cig_off[(size_t)r + 1] = cig_off[(size_t)r] + no;
        }
        const int64_t n_rec = out.rec_off[(size_t)n], n_ops = cig_off[(size_t)n];
        int64_t *d_rec_off = ex_.template ensure<int64_t>(b_recoff_, (size_t)n + 1);
        int64_t *d_cig_off = ex_.template ensure<int64_t>(b_cigoff_, (size_t)n + 1);
        Rec *d_recs = ex_.template ensure<Rec>(b_recs_, (size_t)n_rec + 1);
        uint32_t *d_cig = ex_.template ensure<uint32_t>(b_cig_, (size_t)n_ops + 1);
        ex_.to_exec(d_rec_off, out.rec_off.data(), ((size_t)n + 1) * 8);
        ex_.to_exec(d_cig_off, cig_off.data(), ((size_t)n + 1) * 8);
        FWrite fw;
        fw.in = in; fw.p = p; fw.n = n;
        fw.out1 = P1_.out; fw.fin1 = P1_.fin; fw.res1 = P1_.res; fw.ops1 = P1_.ops;
        fw.out2 = have2 ? P2_.out : nullptr; fw.fin2 = P2_.fin; fw.res2 = P2_.res; fw.ops2 = P2_.ops;
        fw.rec_off = d_rec_off; fw.cig_off = d_cig_off; fw.recs = d_recs; fw.cig = d_cig;
        ex_.per_ids_write(ids, (int64_t)ids_host.size(), fw);
        Rec *h_recs = ex_.template host<Rec>(h_recs_, (size_t)n_rec + 1);
        uint32_t *h_cig = ex_.template host<uint32_t>(h_cig_, (size_t)n_ops + 1);
        out.recs = h_recs; out.cigar = h_cig; out.n_rec = n_rec; out.n_ops = n_ops;
        out.status.resize((size_t)n);
        unsigned long long hc[2 * CT_COUNT];
        if (n_rec) ex_.to_host(h_recs, d_recs, (size_t)n_rec * sizeof(Rec));
        if (n_ops) ex_.to_host(h_cig, d_cig, (size_t)n_ops * 4);
        ex_.to_host(out.status.data(), status, (size_t)n * 4);
        ex_.to_host(hc, counters, sizeof(hc));
        ex_.sync();
        for (int k = 0; k < CT_COUNT; ++k) out.counters[k] = (int64_t)(hc[k] + hc[CT_COUNT + k]);
        out.counters[CT_SECOND_PASS] = (int64_t)ids2.size();
        // a read with records is OK; one that reached the end without any keeps the status that ended it
        for (int32_t r : ids_host)
            if (out.rec_off[(size_t)r + 1] > out.rec_off[(size_t)r]) out.status[(size_t)r] = ST_OK;
            else if (out.status[(size_t)r] == ST_OK) out.status[(size_t)r] = ST_NO_RECORDS;
        status_host = out.status;
    }

private:
    void run_pass(const BackInput &in, const BackParams &p, const int32_t *ids, int64_t n_ids, const int32_t *nofilter, int32_t *status,
                  unsigned long long *counters, PassState<Exec> &P)
    {
        const int64_t n = in.n_reads, NA = in.NA, NT = in.NT;
        P.sub = ex_.template ensure<Sub>(P.b_sub, (size_t)NA + 1);
        P.jobs = ex_.template ensure<Job>(P.b_jobs, (size_t)NA + 1);
        P.alive = ex_.template ensure<int32_t>(P.b_alive, (size_t)n + 1);
        P.filtered = ex_.template ensure<int32_t>(P.b_filtered, (size_t)n + 1);
        P.nfin = ex_.template ensure<int32_t>(P.b_nfin, (size_t)n + 1);
        P.out = ex_.template ensure<ReadOut>(P.b_out, (size_t)n + 1);
        P.open = ex_.template ensure<int32_t>(P.b_open, (size_t)NA + 1);
        ex_.zero(P.alive, (size_t)n * 4);
        ex_.zero(P.filtered, (size_t)n * 4);
        ex_.zero(P.out, (size_t)n * sizeof(ReadOut));
        // neutral jobs everywhere first: sub-alignments of reads outside `ids` (second pass) must not be bounded / filled
        ex_.zero(P.jobs, (size_t)NA * sizeof(Job));
        FInit fi{in, p, ids, P.sub, P.jobs, P.alive, status};
        ex_.per_item(n_ids, fi);
        // ---- divergence filter: distance bounds, exact distances where the bound is not enough ----
        ex_.ed_bounds(P.jobs, NA, in.al_anc);
        if (Exec::kBoundsAreUpper) {
            FOpen fo{P.jobs, P.open, counters};
            ex_.per_item(NA, fo);
            unsigned long long n_open = 0;
            ex_.to_host(&n_open, counters + CT_OPEN_ED, 8);
            ex_.sync();
            if (n_open) ex_.ed_exact(P.jobs, P.open, (int64_t)n_open);
        }
        // ---- extensions + misplaced sub-alignments (warp per read: the z-drop DP is warp-collective) ----
        FExtend fe{in, p, ids, P.sub, P.jobs, P.alive, P.filtered, nofilter, counters};
        ex_.per_item_warp(n_ids, fe);
        // ---- merge / inversion fix / fill jobs ----
        P.fa = ex_.template ensure<Anc>(P.b_fa, (size_t)(NT + 2 * NA) + 1);
        P.seg = ex_.template ensure<Seg>(P.b_seg, (size_t)NA + 1);
        P.first = ex_.template ensure<Anc>(P.b_first, (size_t)NA + 1);
        P.last = ex_.template ensure<Anc>(P.b_last, (size_t)NA + 1);
        P.dup = ex_.template ensure<uint8_t>(P.b_dup, (size_t)NA + 1);
        P.fin = ex_.template ensure<Fin>(P.b_fin, (size_t)NA + 1);
        // a fill needs >= 200 read bases unless it closes a sub-alignment (:21544-21547): an upper bound of the job count
        const int64_t job_cap = in.total_bases / 200 + 2 * NA + 64;
        P.fill = ex_.template ensure<Job>(P.b_fill, (size_t)job_cap);
        ex_.zero(P.fill, (size_t)job_cap * sizeof(Job));      // slots nobody writes are empty jobs
        FFinalize ff{in, p, ids, P.sub, P.alive, status, P.nfin, P.fa, P.first, P.last, P.seg, P.dup, P.fin, P.fill, counters, job_cap};
        ex_.per_item(n_ids, ff);
        unsigned long long h2[2] = {0, 0};
        ex_.to_host(h2, counters + CT_N_JOBS, 16);
        ex_.sync();
        P.n_jobs = std::min<int64_t>((int64_t)h2[0], job_cap);
        P.res = ex_.template ensure<U2>(P.b_res, (size_t)P.n_jobs + 1);
        P.ops = nullptr;
        if (P.n_jobs > 0) ex_.fill(P.fill, P.n_jobs, (int64_t)h2[1], p.eqx != 0, P.res, &P.ops);
        // ---- records: counts, CIGAR length check, pairedindel ----
        FCount fc{in, p, ids, P.alive, status, P.nfin, P.filtered, nofilter, P.fin, P.res, P.ops, P.out};
        ex_.per_item(n_ids, fc);
        // the allocation counters start from zero again for the next pass
        ex_.zero(counters + CT_N_JOBS, 16);
        ex_.zero(counters + CT_FA, 8);
    }

    Exec &ex_;
    PassState<Exec> P1_, P2_;
    typename Exec::Buf b_status_, b_ids_, b_ids2_, b_counters_, b_nofilter_, b_recoff_, b_cigoff_, b_recs_, b_cig_;
    typename Exec::HostBuf h_recs_, h_cig_;
};

// ===============================================================================================================
// Front half on the executor: hit2work_1's bookkeeping, guide selection and the re-seeding jobs of every read
// (vm_dglue.hpp::front_read), one thread per read.  The job slots of a read are the slots of its extracted chains
// (xrec.meta_off ...), its windows / guide points the slots of its extracted anchors (xrec.anc_off ...).
// ===============================================================================================================
struct FrontInput {
    int64_t n_reads;
    const int64_t *read_off;
    Ctg ctg;
    const int32_t *need_reverse;
    const ExtractRec *xrec;          // global extraction of every read
    const A32 *anc;
    const double *S;
    const int32_t *chain_len;
    const double *chain_score;
    int64_t NA, NC;                  // extracted anchors / chains of the chunk (host-known scalars)
    int32_t max_guides, kmer;
};

struct FFront {
    FrontInput in;
    const int32_t *ids;
    // scratch, unsliced: nc-sized arrays are indexed at meta_off, na-sized ones at anc_off
    int32_t *order, *rest, *head, *tail, *nxt, *size, *cseg, *cpos, *cflat, *sc0, *sc1, *ordc;
    int64_t *dist, *k64c;
    double *kd;
    int32_t *coff, *pstart, *bins, *cur, *orda;
    int64_t *k64a;
    Anc *tmp;
    // outputs
    RJob *jobs;
    int64_t *wlo, *whi;
    int32_t *gx;
    int64_t *gy;
    FrontOut *out;
    VM_HD void operator()(int64_t t) const
    {
        const int32_t r = ids[t];
        const ExtractRec xr = in.xrec[r];
        ReadCtx rc;
        rc.read = r; rc.L = in.read_off[r + 1] - in.read_off[r]; rc.need_reverse = in.need_reverse[r] != 0; rc.ctg = in.ctg;
        FrontScratch W;
        const long long c = xr.meta_off, a = xr.anc_off;
        W.order = order + c; W.rest = rest + c; W.head = head + c; W.tail = tail + c; W.nxt = nxt + c; W.size = size + c;
        W.cseg = cseg + c; W.cpos = cpos + c; W.cflat = cflat + c; W.sc0 = sc0 + c; W.sc1 = sc1 + c; W.ordc = ordc + c;
        W.dist = dist + c; W.k64c = k64c + c; W.kd = kd + c;
        W.coff = coff + a; W.pstart = pstart + a; W.bins = bins + a; W.cur = cur + a; W.orda = orda + a; W.k64a = k64a + a; W.tmp = tmp + a;
        FrontOut o;
        front_read(rc, xr, in.anc, in.S, in.chain_len, in.chain_score, in.max_guides, in.kmer, W, jobs + c, wlo + a, whi + a, gx + a, gy + a, a, a, o);
        out[r] = o;
    }
};

template <typename Exec>
class FrontHalf {
public:
    explicit FrontHalf(Exec &ex) : ex_(ex) {}
    ~FrontHalf()
    {
        for (auto *x : {&b_ids_, &b_i32c_, &b_i64c_, &b_f64c_, &b_i32a_, &b_i64a_, &b_tmp_, &b_out_}) Exec::release(*x);
    }
    // jobs / wlo / whi / gx / gy: executor arrays with room for NC jobs and NA windows / guide points (jobs zeroed here).
    // fo: per read (only the entries of `ids_host` are meaningful); J: host copy of the NC job slots.
    void run(const FrontInput &in, const std::vector<int32_t> &ids_host, RJob *jobs, int64_t *wlo, int64_t *whi, int32_t *gx, int64_t *gy,
             std::vector<FrontOut> &fo, std::vector<RJob> &J, std::vector<ExtractRec> &xrec_host)
    {
        const int64_t n = in.n_reads;
        const size_t NC = (size_t)in.NC + 1, NA = (size_t)in.NA + 1;
        int32_t *ids = ex_.template ensure<int32_t>(b_ids_, ids_host.size() + 1);
        int32_t *i32c = ex_.template ensure<int32_t>(b_i32c_, 12 * NC);
        int64_t *i64c = ex_.template ensure<int64_t>(b_i64c_, 2 * NC);
        double *f64c = ex_.template ensure<double>(b_f64c_, NC);
        int32_t *i32a = ex_.template ensure<int32_t>(b_i32a_, 5 * NA);
        int64_t *i64a = ex_.template ensure<int64_t>(b_i64a_, NA);
        Anc *tmp = ex_.template ensure<Anc>(b_tmp_, NA);
        FrontOut *out = ex_.template ensure<FrontOut>(b_out_, (size_t)n + 1);
        ex_.to_exec(ids, ids_host.data(), ids_host.size() * 4);
        ex_.zero(jobs, NC * sizeof(RJob));
        ex_.zero(out, ((size_t)n + 1) * sizeof(FrontOut));
        FFront f;
        f.in = in; f.ids = ids;
        f.order = i32c; f.rest = i32c + NC; f.head = i32c + 2 * NC; f.tail = i32c + 3 * NC; f.nxt = i32c + 4 * NC; f.size = i32c + 5 * NC;
        f.cseg = i32c + 6 * NC; f.cpos = i32c + 7 * NC; f.cflat = i32c + 8 * NC; f.sc0 = i32c + 9 * NC; f.sc1 = i32c + 10 * NC; f.ordc = i32c + 11 * NC;
        f.dist = i64c; f.k64c = i64c + NC; f.kd = f64c;
        f.coff = i32a; f.pstart = i32a + NA; f.bins = i32a + 2 * NA; f.cur = i32a + 3 * NA; f.orda = i32a + 4 * NA;
        f.k64a = i64a; f.tmp = tmp;
        f.jobs = jobs; f.wlo = wlo; f.whi = whi; f.gx = gx; f.gy = gy; f.out = out;
        ex_.per_item((int64_t)ids_host.size(), f);
        fo.resize((size_t)n);
        J.resize((size_t)in.NC);
        xrec_host.resize((size_t)n);
        ex_.to_host(fo.data(), out, (size_t)n * sizeof(FrontOut));
        if (in.NC) ex_.to_host(J.data(), jobs, (size_t)in.NC * sizeof(RJob));
        ex_.to_host(xrec_host.data(), in.xrec, (size_t)n * sizeof(ExtractRec));
        ex_.sync();
    }

private:
    Exec &ex_;
    typename Exec::Buf b_ids_, b_i32c_, b_i64c_, b_f64c_, b_i32a_, b_i64a_, b_tmp_, b_out_;
};

} // namespace vmd
