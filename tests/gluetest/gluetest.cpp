// TEST HARNESS (not product): runs the product's host driver + glue (vm_pipeline.hpp, vm_glue.hpp)
// on a CPU-only box with the ORACLE's C stage functions standing in for the CUDA kernels, so the
// glue can be checked against the reference-generated golden records without a GPU.
#include "../../vacmap_b200/csrc/vm_pipeline.hpp"
#include "../../vacmap_b200/csrc/vm_dgrun.hpp"
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>

extern "C" {
// oracle/orc_*.c
typedef struct { const float *extra; int64_t extra_size; const float *readgapcost; const double *log2cache; int64_t log2cache_size; } orc_tables;
void orc_argsort_i64(const int64_t *A, int64_t n, int64_t *R);
int64_t orc_chain_global_d_all(const int64_t *a, int64_t n, int kmersize, double skipcost_in, int64_t maxdiff_in, int64_t maxgap,
                               const orc_tables *tb, int64_t max_factor, double *S, int32_t *P, int32_t *S_arg, int64_t *opcount_out);
int64_t orc_chain_fast(const int64_t *a, int64_t n, int kmersize, int variant, double skipcost_in, int64_t maxdiff_in, int64_t maxgap,
                       int64_t fast_t, const orc_tables *tb, const float *rgcost, double *S, int32_t *P, int32_t *S_arg_i);
int64_t orc_chain_local(const int64_t *a, int64_t n, int kmersize, int variant, double skipcost, int64_t maxdiff, int64_t maxgap,
                        const orc_tables *tb, const float *rgcost, double *S, int64_t *P, int64_t *S_arg, int64_t *opcount_out);
void orc_large_readgap_table(int maxgap, int large_readgap, float *out);
int64_t orc_map(void *h, const char *seq, int64_t len, int32_t check_num, int32_t mid_occ, int64_t *rows, int64_t cap);
int64_t orc_local_reseed(const char *ref, const int64_t *win_lo, const int64_t *win_hi, int32_t n_win, const int32_t *gx,
                         const int64_t *gy, int64_t n_guide, const char *seq, const char *rc_seq, int64_t L, int32_t k,
                         int64_t readstart, int64_t readend, int64_t **rows_out);
void orc_free(void *p);
typedef struct { int32_t score, max_t, max_q, zdropped, n_cigar, q_e, t_e, ndel, nins; } orc_kc_result;
int orc_k_cigar(const char *target, int32_t tlen, const char *query, int32_t qlen, int32_t match, int32_t mismatch, int32_t q1,
                int32_t e1, int32_t q2, int32_t e2, int32_t bw, int32_t zdrop, int32_t eqx, uint32_t *cigar, int32_t cigar_cap,
                orc_kc_result *res);
int64_t orc_edit_distance(const char *a, int64_t n, const char *b, int64_t m);
}

using namespace vmp;

static std::string revcomp(const std::string &s)
{
    std::string r(s.size(), 'N');
    for (size_t i = 0; i < s.size(); ++i) {
        const char c = s[s.size() - 1 - i];
        r[i] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N';
    }
    return r;
}

struct OracleBackend;

// Execution policy of vm_dgrun.hpp for the CPU harness: the per-read functors run in plain loops, the hot loops are
// the oracle's C natives (exact edit distance, k_cigar extension and fill).
struct OracleExec {
    typedef std::vector<char> Buf;
    typedef std::vector<char> HostBuf;
    static constexpr bool kBoundsAreUpper = false;
    OracleBackend *be = nullptr;
    int fill_slot = 0;
    std::vector<uint32_t> ops_[2];
    static void release(Buf &b) { Buf().swap(b); }
    static void release_host(HostBuf &b) { HostBuf().swap(b); }
    template <typename T> T *ensure(Buf &b, size_t n) { if (b.size() < n * sizeof(T) + 64) b.resize(n * sizeof(T) + 64); return (T *)b.data(); }
    template <typename T> T *host(HostBuf &b, size_t n) { return ensure<T>(b, n); }
    void zero(void *p, size_t bytes) { memset(p, 0, bytes); }
    void to_host(void *dst, const void *src, size_t bytes) { memcpy(dst, src, bytes); }
    void to_exec(void *dst, const void *src, size_t bytes) { memcpy(dst, src, bytes); }
    void sync() {}
    template <typename F> void per_item(int64_t n, const F &f) { for (int64_t t = 0; t < n; ++t) f(t); }
    struct Ext {
        OracleBackend *be;
        int32_t read;
        void operator()(const vmd::Spec &t, const vmd::Spec &q, int32_t &q_e, int32_t &t_e);
    };
    template <typename F> void per_item_warp(int64_t n, const F &f)
    {
        Ext ext{be, 0};
        for (int64_t t = 0; t < n; ++t) f(t, ext, true);
    }
    template <typename F> void per_ids_write(const int32_t *ids, int64_t n, const F &f) { for (int64_t t = 0; t < n; ++t) f.run(ids[t], 0, 1); }
    void ed_bounds(vmd::Job *jobs, int64_t n, const vmd::A32 *);
    void ed_exact(vmd::Job *, const int32_t *, int64_t) {}
    void fill(vmd::Job *jobs, int64_t nj, int64_t, bool eqx, vmd::U2 *res, const uint32_t **ops);
};

struct OracleBackend : public Backend {
    void *index;
    const orc_tables *tb;
    const char *ref;
    bool device_ext = false;             // GT_DEVICE_GLUE: run extend_func through vm_dgrun.hpp / vm_dglue.hpp
    const vmg::Contigs *ctg = nullptr;
    const ChainOut *last_lc = nullptr;
    const ReadBatch *cur = nullptr;
    std::string fwd_all, rc_all;
    OracleExec exec;
    std::unique_ptr<vmd::BackHalf<OracleExec>> back;
    OracleBackend(void *ix, const orc_tables *t, const char *r) : index(ix), tb(t), ref(r) {}

    std::string materialize_spec(const vmd::Spec &s, int read) const
    {
        std::string out;
        if (s.src == 0) out.assign(ref + s.lo, (size_t)s.len);
        else out = (s.src == 1 ? fwd_all : rc_all).substr((size_t)(cur->off[read] - cur->off[0] + s.lo), (size_t)s.len);
        if (s.comp)
            for (char &c : out) c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N';
        if (s.reverse) std::reverse(out.begin(), out.end());
        return out;
    }

    // ---- front half through vm_dgrun.hpp / vm_dglue.hpp (what the CUDA kernels run), over a host copy of what the
    // extraction kernel leaves behind ----
    std::unique_ptr<vmd::FrontHalf<OracleExec>> front;
    std::vector<vmd::RJob> fJ;
    std::vector<vmd::ExtractRec> fx;
    std::vector<int64_t> f_wlo, f_whi, f_gy;
    std::vector<int32_t> f_gx;
    bool has_device_front() const override { return device_ext; }

    // vm_extract_global_kernel on the host: primary chain + residual chains with score > 40, discovery order
    static void extract_global_host(const ChainOut &g, int64_t r, double accept, std::vector<vmd::A32> &anc, std::vector<double> &S_out,
                                    std::vector<int32_t> &len, std::vector<double> &score, vmd::ExtractRec &rec)
    {
        memset(&rec, 0, sizeof(rec));
        const int64_t n = g.cnt[(size_t)r], gm = g.gmax[(size_t)r];
        if (n <= 0 || gm < 0) return;
        const Anc32 *a = g.sorted + g.start[(size_t)r];
        const double *S = g.S + g.start[(size_t)r];
        const int32_t *P = g.P + g.start[(size_t)r], *A = g.S_arg + g.start[(size_t)r];
        std::vector<char> used((size_t)n, 0);
        std::vector<vmd::A32> ta;
        std::vector<double> tS, tscore;
        std::vector<int32_t> tlen;
        bool hit = false;
        {
            int64_t take = gm;
            used[(size_t)take] = 1;
            const double sc = S[take];
            for (;;) {
                ta.push_back(vmd::A32{a[take].x, a[take].y, a[take].s, a[take].l});
                tS.push_back(S[take]);
                if (P[take] == vmg::kNoPre) break;
                take = P[take];
                used[(size_t)take] = 1;
            }
            if (sc > 40) { hit = true; tlen.push_back((int32_t)ta.size()); tscore.push_back(sc); }
            else { ta.clear(); tS.clear(); }
        }
        const double scores = S[gm], max_scores = scores > 0 ? scores : 0;
        if (!(hit && max_scores > accept)) return;
        for (int64_t q = n - 1; q >= 0; --q) {
            int64_t take = A[q];
            if (used[(size_t)take]) continue;
            const size_t k0 = ta.size();
            used[(size_t)take] = 1;
            double sc = S[take];
            for (;;) {
                ta.push_back(vmd::A32{a[take].x, a[take].y, a[take].s, a[take].l});
                tS.push_back(0.0);
                if (P[take] == vmg::kNoPre) break;
                take = P[take];
                if (used[(size_t)take]) { sc = sc - S[take]; break; }
                used[(size_t)take] = 1;
            }
            if (sc > 40) { tlen.push_back((int32_t)(ta.size() - k0)); tscore.push_back(sc); }
            else { ta.resize(k0); tS.resize(k0); }
        }
        rec.n_anc = (int32_t)ta.size();
        rec.n_chains = (int32_t)tlen.size();
        rec.anc_off = (long long)anc.size();
        rec.meta_off = (long long)len.size();
        anc.insert(anc.end(), ta.begin(), ta.end());
        S_out.insert(S_out.end(), tS.begin(), tS.end());
        len.insert(len.end(), tlen.begin(), tlen.end());
        score.insert(score.end(), tscore.begin(), tscore.end());
    }

    double accept_ = 60.0;
    bool front_device(const ReadBatch &b, const std::vector<char> &need_reverse, const ChainOut &g, int max_guides,
                      std::vector<vmd::FrontOut> &fo) override
    {
        if (!device_ext) return false;
        const int64_t n = b.n;
        std::vector<vmd::ExtractRec> xrec((size_t)n);
        std::vector<vmd::A32> anc;
        std::vector<double> S, score;
        std::vector<int32_t> len, ids, nrev((size_t)n);
        std::vector<int64_t> off((size_t)n + 1);
        // reads in REVERSE order: meta_off / anc_off must not be assumed to grow with the read index
        for (int64_t r = n - 1; r >= 0; --r) extract_global_host(g, r, accept_, anc, S, len, score, xrec[(size_t)r]);
        for (int64_t r = 0; r <= n; ++r) off[(size_t)r] = b.off[r] - b.off[0];
        for (int64_t r = 0; r < n; ++r) {
            nrev[(size_t)r] = need_reverse[(size_t)r] ? 1 : 0;
            if (g.cnt[(size_t)r] > 2) ids.push_back((int32_t)r);
        }
        anc.push_back(vmd::A32{0, 0, 0, 0}); S.push_back(0); len.push_back(0); score.push_back(0);
        vmd::FrontInput in;
        in.n_reads = n; in.read_off = off.data();
        in.ctg.start = ctg->start.data(); in.ctg.len = ctg->len.data(); in.ctg.n = (int32_t)ctg->start.size();
        in.need_reverse = nrev.data();
        in.xrec = xrec.data(); in.anc = anc.data(); in.S = S.data(); in.chain_len = len.data(); in.chain_score = score.data();
        in.NA = (int64_t)anc.size() - 1; in.NC = (int64_t)len.size() - 1;
        in.max_guides = max_guides; in.kmer = 9;
        std::vector<vmd::RJob> jobs((size_t)in.NC + 1);
        f_wlo.assign((size_t)in.NA + 1, 0); f_whi.assign((size_t)in.NA + 1, 0); f_gx.assign((size_t)in.NA + 1, 0); f_gy.assign((size_t)in.NA + 1, 0);
        exec.be = this;
        if (!front) front.reset(new vmd::FrontHalf<OracleExec>(exec));
        front->run(in, ids, jobs.data(), f_wlo.data(), f_whi.data(), f_gx.data(), f_gy.data(), fo, fJ, fx);
        return true;
    }

    void reseed_chain_front(const ReadBatch &b, const std::vector<char> &need_reverse, const std::vector<vmd::FrontOut> &fo,
                            const std::vector<int> &variant, const std::vector<double> &skipcost, int maxdiff, int maxgap,
                            ChainOut &out) override
    {
        // back to the job form of the host path, in read order, for the oracle's re-seeding
        std::vector<GuideJobRef> jobs;
        for (int64_t r = 0; r < b.n; ++r) {
            if (variant[(size_t)r] == 0) continue;
            for (int q = 0; q < fo[(size_t)r].n_jobs; ++q) {
                const vmd::RJob &J = fJ[(size_t)(fx[(size_t)r].meta_off + q)];
                GuideJobRef gj;
                gj.read = (int32_t)r;
                gj.job.readstart = J.readstart; gj.job.readend = J.readend;
                gj.job.win_lo.assign(f_wlo.begin() + J.win_off, f_wlo.begin() + J.win_off + J.n_win);
                gj.job.win_hi.assign(f_whi.begin() + J.win_off, f_whi.begin() + J.win_off + J.n_win);
                gj.job.gx.assign(f_gx.begin() + J.g_off, f_gx.begin() + J.g_off + J.n_guide);
                gj.job.gy.assign(f_gy.begin() + J.g_off, f_gy.begin() + J.g_off + J.n_guide);
                jobs.push_back(std::move(gj));
            }
        }
        reseed_chain(b, need_reverse, jobs, variant, skipcost, maxdiff, maxgap, out);
    }

    bool has_device_extension() const override { return device_ext; }
    bool extend_device(const ReadBatch &b, const std::vector<int32_t> &ids, const std::vector<char> &need_reverse,
                       const std::vector<int32_t> &mapq, const vmg::Options &opt, std::vector<int32_t> &status, FlatRecords &out) override
    {
        if (!device_ext) return false;
        const int64_t n = b.n;
        cur = &b;
        fwd_all.assign(b.seq + b.off[0], (size_t)(b.off[n] - b.off[0]));
        rc_all.resize(fwd_all.size());
        std::vector<int64_t> off((size_t)n + 1);
        for (int64_t r = 0; r <= n; ++r) off[(size_t)r] = b.off[r] - b.off[0];
        for (int64_t r = 0; r < n; ++r) {
            const std::string rc = revcomp(fwd_all.substr((size_t)off[(size_t)r], (size_t)(off[(size_t)r + 1] - off[(size_t)r])));
            memcpy(&rc_all[(size_t)off[(size_t)r]], rc.data(), rc.size());
        }
        // what the device kernels would have left behind: the extracted local chain and its rebuilt sub-alignments
        const ChainOut &lc = *last_lc;
        std::vector<vmd::ExtractRec> xrec((size_t)n);
        std::vector<vmd::RebuildRec> rrec((size_t)n);
        std::vector<vmd::A32> al_anc;
        std::vector<int32_t> al_len, cnt((size_t)n), nrev((size_t)n), mq(mapq);
        std::vector<vmg::AlnList> als((size_t)n);
        for (int64_t r = 0; r < n; ++r) {
            memset(&xrec[(size_t)r], 0, sizeof(vmd::ExtractRec));
            memset(&rrec[(size_t)r], 0, sizeof(vmd::RebuildRec));
            cnt[(size_t)r] = lc.cnt[(size_t)r];
            nrev[(size_t)r] = need_reverse[(size_t)r] ? 1 : 0;
            if (lc.cnt[(size_t)r] <= 0 || lc.gmax[(size_t)r] < 0) continue;
            Path asc;
            vmg::local_traceback(lc.sorted + lc.start[(size_t)r], lc.P + lc.start[(size_t)r], lc.gmax[(size_t)r], asc);
            xrec[(size_t)r].n_anc = (int32_t)asc.size();
            xrec[(size_t)r].n_chains = 1;
            if (asc.size() <= 1) continue;
            try { vmg::rebuild_chain_break(*ctg, asc, opt.local_maxdiff, als[(size_t)r]); } catch (const vmg::ReadDropped &) { als[(size_t)r].clear(); }
        }
        // the rebuild kernel claims its anchor room and its length room from two independent counters, in whatever order the
        // threads arrive: here the lengths are laid out in read order and the anchors in REVERSE read order, so that nothing
        // downstream can rely on the two offsets growing together
        for (int64_t r = 0; r < n; ++r) {
            rrec[(size_t)r].len_off = (long long)al_len.size();
            rrec[(size_t)r].n_al = (int32_t)als[(size_t)r].size();
            for (const Path &p : als[(size_t)r]) al_len.push_back((int32_t)p.size());
        }
        for (int64_t r = n - 1; r >= 0; --r) {
            rrec[(size_t)r].anc_off = (long long)al_anc.size();
            for (const Path &p : als[(size_t)r])
                for (const Anc &a : p) al_anc.push_back(vmd::A32{(int32_t)a.x, (uint32_t)a.y, a.s, a.l});
            rrec[(size_t)r].n_anc = (int32_t)(al_anc.size() - (size_t)rrec[(size_t)r].anc_off);
        }
        al_anc.push_back(vmd::A32{0, 0, 0, 0});
        al_len.push_back(0);
        vmd::BackInput in;
        in.n_reads = n;
        in.read_off = off.data();
        in.reads_fwd = (const uint8_t *)fwd_all.data();
        in.reads_rc = (const uint8_t *)rc_all.data();
        in.ref = (const uint8_t *)ref;
        in.ctg.start = ctg->start.data(); in.ctg.len = ctg->len.data(); in.ctg.n = (int32_t)ctg->start.size();
        in.need_reverse = nrev.data();
        in.mapq = mq.data();
        in.local_cnt = cnt.data();
        in.xrec = xrec.data(); in.rrec = rrec.data(); in.al_anc = al_anc.data(); in.al_len = al_len.data();
        in.NA = (int64_t)al_len.size() - 1; in.NT = (int64_t)al_anc.size() - 1;
        in.total_bases = off[(size_t)n];
        vmd::BackParams p;
        p.maxdivergence = opt.maxdivergence; p.eqx = opt.eqx; p.hardclip = opt.hardclip; p.nodiscard = opt.nodiscard;
        exec.be = this;
        if (!back) back.reset(new vmd::BackHalf<OracleExec>(exec));
        vmd::BackResult br;
        back->run(in, p, ids, status, br);
        out.rec_off.swap(br.rec_off);
        out.recs = br.recs; out.cigar = br.cigar; out.n_rec = br.n_rec; out.n_ops = br.n_ops;
        for (int k = 0; k < vmd::CT_COUNT; ++k) out.counters[k] = br.counters[k];
        return true;
    }

    std::string materialize(const ReadBatch &b, int read, const vmg::SeqRef &s) const
    {
        std::string out;
        if (s.src == 0) out.assign(ref + s.lo, (size_t)(s.hi - s.lo));
        else {
            std::string fwd(b.seq + b.off[read], (size_t)b.len(read));
            if (s.src == 2) fwd = revcomp(fwd);
            out = fwd.substr((size_t)s.lo, (size_t)(s.hi - s.lo));
        }
        if (s.comp)
            for (char &c : out) c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N';
        if (s.reverse) std::reverse(out.begin(), out.end());
        return out;
    }

    static void put(ChainOut &out, int64_t r, const std::vector<Anc> &srt, const std::vector<double> &S,
                    const std::vector<int32_t> &P, const std::vector<int32_t> &A, int64_t g)
    {
        out.start[r] = (int64_t)out.sorted_store.size();
        out.cnt[r] = (int32_t)srt.size();
        out.gmax[r] = g;
        for (size_t t = 0; t < srt.size(); ++t) {
            out.sorted_store.push_back(vmg::Anc32{(int32_t)srt[t].x, (uint32_t)srt[t].y, srt[t].s, srt[t].l});
            out.S_store.push_back(S[t]);
            out.P_store.push_back(P[t]);
            out.A_store.push_back(A[t]);
        }
    }

    static void to_rows(const Anc *a, int64_t n, std::vector<int64_t> &rows)
    {
        rows.resize((size_t)n * 4);
        for (int64_t t = 0; t < n; ++t) { rows[t * 4] = a[t].x; rows[t * 4 + 1] = a[t].y; rows[t * 4 + 2] = a[t].s; rows[t * 4 + 3] = a[t].l; }
    }

    void seed_chain(const ReadBatch &b, int check_num, int kmersize, double skipcost, int maxdiff, int maxgap, double,
                    std::vector<char> &need_reverse, ChainOut &out) override
    {
        out = ChainOut();
        out.start.assign((size_t)b.n, 0); out.cnt.assign((size_t)b.n, 0); out.gmax.assign((size_t)b.n, -1);
        out.used_fast.assign((size_t)b.n, 0);
        need_reverse.assign((size_t)b.n, 0);
        for (int64_t r = 0; r < b.n; ++r) {
            const int64_t L = b.len(r);
            std::vector<int64_t> rows((size_t)(4 * (4 * L + 4096)));
            int64_t m = orc_map(index, b.seq + b.off[r], L, check_num, -1, rows.data(), (int64_t)rows.size() / 4);
            if (m < 0) { rows.resize((size_t)(-m * 4 + 64)); m = orc_map(index, b.seq + b.off[r], L, check_num, -1, rows.data(), -m + 16); }
            // get_reversed_chain_numpy_rough :21202-21217
            int64_t neg = 0, pos = 0;
            for (int64_t t = 0; t < m; ++t) (rows[t * 4 + 2] == 1 ? pos : neg)++;
            const bool flip = m >= 3 && neg > pos;
            need_reverse[r] = flip;
            std::vector<Anc> anch;
            for (int64_t t = 0; t < m; ++t) {
                const int64_t q = flip ? m - 1 - t : t;
                Anc a{rows[q * 4], rows[q * 4 + 1], (int32_t)rows[q * 4 + 2], (int32_t)rows[q * 4 + 3]};
                if (flip) { a.x = L - a.x - a.l; a.s = -a.s; }
                anch.push_back(a);
            }
            const int64_t n = (int64_t)anch.size();
            std::vector<int64_t> keys((size_t)n), perm((size_t)n), srows;
            for (int64_t t = 0; t < n; ++t) keys[t] = anch[t].x;
            orc_argsort_i64(keys.data(), n, perm.data());
            std::vector<Anc> srt((size_t)n);
            for (int64_t t = 0; t < n; ++t) srt[t] = anch[perm[t]];
            to_rows(srt.data(), n, srows);
            std::vector<double> S((size_t)n);
            std::vector<int32_t> P((size_t)n), A((size_t)n);
            int64_t g = -1;
            if (n > 0) {
                const bool fast = (double)n / (double)L > 5.0;
                if (!fast) g = orc_chain_global_d_all(srows.data(), n, kmersize, skipcost, maxdiff, maxgap, tb, 1000, S.data(), P.data(), A.data(), nullptr);
                if (fast || g == -1) {
                    g = orc_chain_fast(srows.data(), n, kmersize, 0, skipcost, maxdiff, maxgap, 5, tb, nullptr, S.data(), P.data(), A.data());
                    if (n > 2) out.used_fast[(size_t)r] = 1;
                }
            }
            put(out, r, srt, S, P, A, g);
        }
        out.adopt_stores();
    }

    void reseed_chain(const ReadBatch &b, const std::vector<char> &need_reverse, const std::vector<GuideJobRef> &jobs,
                      const std::vector<int> &variant, const std::vector<double> &skipcost, int maxdiff, int maxgap,
                      ChainOut &out) override
    {
        out = ChainOut();
        out.start.assign((size_t)b.n, 0); out.cnt.assign((size_t)b.n, 0); out.gmax.assign((size_t)b.n, -1);
        std::vector<float> lrg((size_t)maxgap + 1);
        orc_large_readgap_table(maxgap, 30, lrg.data());
        size_t q = 0;
        for (int64_t r = 0; r < b.n; ++r) {
            std::vector<Anc> anch;
            while (q < jobs.size() && jobs[q].read == r) {
                const vmg::GuideJob &j = jobs[q].job;
                std::string fwd(b.seq + b.off[r], (size_t)b.len(r)), rc = revcomp(fwd);
                if (need_reverse[r]) std::swap(fwd, rc);
                int64_t *rows = nullptr;
                const int64_t m = orc_local_reseed(ref, j.win_lo.data(), j.win_hi.data(), (int32_t)j.win_lo.size(), j.gx.data(),
                                                   j.gy.data(), (int64_t)j.gx.size(), fwd.c_str(), rc.c_str(), b.len(r), 9,
                                                   j.readstart, j.readend, &rows);
                for (int64_t t = 0; t < m; ++t)
                    anch.push_back(Anc{rows[t * 4], rows[t * 4 + 1], (int32_t)rows[t * 4 + 2], (int32_t)rows[t * 4 + 3]});
                if (rows) orc_free(rows);
                ++q;
            }
            const int64_t n = variant[r] ? (int64_t)anch.size() : 0;
            std::vector<int64_t> keys((size_t)n), perm((size_t)n), srows;
            for (int64_t t = 0; t < n; ++t) keys[t] = anch[t].x + anch[t].l;
            orc_argsort_i64(keys.data(), n, perm.data());
            std::vector<Anc> srt((size_t)n);
            for (int64_t t = 0; t < n; ++t) srt[t] = anch[perm[t]];
            to_rows(srt.data(), n, srows);
            std::vector<double> S((size_t)n);
            std::vector<int32_t> P((size_t)n), A32((size_t)n);
            int64_t g = -1;
            if (n > 0) {
                std::vector<int64_t> P64((size_t)n), A64((size_t)n);
                const float *rg = variant[r] == 1 ? tb->readgapcost : lrg.data();
                g = orc_chain_local(srows.data(), n, 9, variant[r], skipcost[r], maxdiff, maxgap, tb, rg, S.data(), P64.data(), A64.data(), nullptr);
                if (g == -2) g = orc_chain_fast(srows.data(), n, 9, variant[r], skipcost[r], maxdiff, maxgap, 5, tb, rg, S.data(), P.data(), A32.data());
                else for (int64_t t = 0; t < n; ++t) P[t] = (int32_t)P64[t];
            }
            put(out, r, srt, S, P, A32, g);
        }
        out.adopt_stores();
        last_lc = &out;
    }

    void edit_distance(const ReadBatch &b, std::vector<EdJob> &jobs, const vmg::MatchSeg *, size_t) override
    {
        for (EdJob &j : jobs) {
            const std::string x = materialize(b, j.read, j.a), y = materialize(b, j.read, j.b);
            j.dist = orc_edit_distance(x.data(), (int64_t)x.size(), y.data(), (int64_t)y.size());
        }
    }

    void extend(const ReadBatch &b, std::vector<ExtJobRef> &jobs) override
    {
        for (ExtJobRef &j : jobs) {
            const std::string t = materialize(b, j.read, j.job.target), q = materialize(b, j.read, j.job.query);
            orc_kc_result res;
            std::vector<uint32_t> cig(t.size() + q.size() + 4);
            orc_k_cigar(t.data(), (int32_t)t.size(), q.data(), (int32_t)q.size(), 2, -4, 4, 4, 4, 4, 100, 50, 0, cig.data(), (int32_t)cig.size(), &res);
            j.job.q_e = res.q_e;
            j.job.t_e = res.t_e;
        }
    }

    std::vector<uint32_t> fill_ops_;
    const uint32_t *fill(const ReadBatch &b, bool eqx, std::vector<FillJobRef> &jobs) override
    {
        fill_ops_.clear();
        for (FillJobRef &j : jobs) {
            const std::string t = materialize(b, j.read, j.job.target), q = materialize(b, j.read, j.job.query);
            orc_kc_result res;
            std::vector<uint32_t> cig(t.size() + q.size() + 4);
            orc_k_cigar(t.data(), (int32_t)t.size(), q.data(), (int32_t)q.size(), 2, -4, 4, 2, 24, 1, -1, -1, eqx ? 1 : 0, cig.data(), (int32_t)cig.size(), &res);
            j.cig_off = (int64_t)fill_ops_.size();
            j.cig_len = res.n_cigar;
            fill_ops_.insert(fill_ops_.end(), cig.begin(), cig.begin() + res.n_cigar);
        }
        return fill_ops_.data();
    }
};

void OracleExec::Ext::operator()(const vmd::Spec &t, const vmd::Spec &q, int32_t &q_e, int32_t &t_e)
{
    const std::string ts = be->materialize_spec(t, read), qs = be->materialize_spec(q, read);
    orc_kc_result res;
    std::vector<uint32_t> cig(ts.size() + qs.size() + 4);
    orc_k_cigar(ts.data(), (int32_t)ts.size(), qs.data(), (int32_t)qs.size(), 2, -4, 4, 4, 4, 4, 100, 50, 0, cig.data(), (int32_t)cig.size(), &res);
    q_e = res.q_e;
    t_e = res.t_e;
}

void OracleExec::ed_bounds(vmd::Job *jobs, int64_t n, const vmd::A32 *)
{
    for (int64_t j = 0; j < n; ++j) {
        vmd::Job &J = jobs[j];
        if (J.t.len <= 0 || J.q.len <= 0) continue;
        const std::string x = be->materialize_spec(J.q, J.read), y = be->materialize_spec(J.t, J.read);
        J.result0 = orc_edit_distance(x.data(), (int64_t)x.size(), y.data(), (int64_t)y.size());
    }
}

void OracleExec::fill(vmd::Job *jobs, int64_t nj, int64_t, bool eqx, vmd::U2 *res, const uint32_t **ops)
{
    std::vector<uint32_t> &O = ops_[fill_slot];
    O.clear();
    for (int64_t j = 0; j < nj; ++j) {
        vmd::Job &J = jobs[j];
        res[j].x = (uint32_t)O.size();
        res[j].y = 0;
        if (J.t.len <= 0 || J.q.len <= 0) continue;
        const std::string t = be->materialize_spec(J.t, J.read), q = be->materialize_spec(J.q, J.read);
        orc_kc_result r;
        std::vector<uint32_t> cig(t.size() + q.size() + 4);
        orc_k_cigar(t.data(), (int32_t)t.size(), q.data(), (int32_t)q.size(), 2, -4, 4, 2, 24, 1, -1, -1, eqx ? 1 : 0, cig.data(), (int32_t)cig.size(), &r);
        res[j].y = (uint32_t)r.n_cigar;
        O.insert(O.end(), cig.begin(), cig.begin() + r.n_cigar);
    }
    O.push_back(0);
    *ops = O.data();
}

extern "C" {

struct gt_options {
    double global_skipcost, local_skipcost, maxdivergence, accept;
    int32_t global_maxdiff, local_maxdiff, check_num, eqx, hardclip, nodiscard, max_guides, local_maxgap, clamp40, kmersize, threads;
};

// returns number of records; flat outputs sized by the caller (rec_cap rows of 9 int64, cigar_cap uint32)
int64_t gt_align_batch(void *orc_index, const orc_tables *tb, const char *ref, const int64_t *ctg_start, const int64_t *ctg_len,
                       int32_t n_ctg, const char *reads, const int64_t *read_off, int64_t n_reads, const gt_options *o,
                       int64_t *rec_rows, int64_t rec_cap, uint32_t *cigar, int64_t cigar_cap, int64_t *n_cigar_out)
{
    vmg::Contigs ctg;
    for (int c = 0; c < n_ctg; ++c) { ctg.names.push_back("c" + std::to_string(c)); ctg.start.push_back(ctg_start[c]); ctg.len.push_back(ctg_len[c]); }
    ctg.seq = ref;
    ctg.total = n_ctg ? ctg_start[n_ctg - 1] + ctg_len[n_ctg - 1] : 0;
    vmg::Options opt;
    opt.global_skipcost = o->global_skipcost; opt.local_skipcost = o->local_skipcost; opt.maxdivergence = o->maxdivergence;
    opt.global_maxdiff = o->global_maxdiff; opt.local_maxdiff = o->local_maxdiff; opt.check_num = o->check_num;
    opt.eqx = o->eqx; opt.hardclip = o->hardclip; opt.nodiscard = o->nodiscard;
    opt.mode = vmg::ModeConst{o->accept, o->max_guides, o->local_maxgap, o->clamp40 != 0};
    OracleBackend be(orc_index, tb, ref);
    be.ctg = &ctg;
    be.accept_ = o->accept;
    be.device_ext = getenv("GT_DEVICE_GLUE") != nullptr;
    Driver drv(be, ctg, opt, o->kmersize, o->threads);
    std::map<std::string, double> phase_ms;
    if (getenv("GT_TIMES")) drv.on_time = [&](const char *nm, double ms) { phase_ms[nm] += ms; };
    ReadBatch b;
    b.n = n_reads; b.seq = reads; b.off = read_off;
    BatchResult res;
    drv.align_batch(b, res);
    for (auto &kv : phase_ms) fprintf(stderr, "GT_TIME %s %.3f\n", kv.first.c_str(), kv.second);
    int64_t nrec = 0, ncig = 0;
    if (const char *cf = getenv("GT_COUNTERS")) {
        FILE *f = fopen(cf, "w");
        if (f) {
            for (int k = 0; k < BC_COUNT; ++k) fprintf(f, "%s %lld\n", kBranchName[k], (long long)res.branch[k]);
            for (int64_t r = 0; r < n_reads; ++r) fprintf(f, "status %lld %d\n", (long long)r, (int)res.status[(size_t)r]);
            fclose(f);
        }
    }
    if (res.flat) {
        for (int64_t r = 0; r < n_reads; ++r)
            for (int64_t q = res.fr.rec_off[(size_t)r]; q < res.fr.rec_off[(size_t)r + 1]; ++q) {
                const vmd::Rec &rec = res.fr.recs[q];
                if (nrec < rec_cap && ncig + rec.cigar_len <= cigar_cap) {
                    int64_t *row = rec_rows + nrec * 9;
                    row[0] = r; row[1] = rec.contig; row[2] = rec.strand; row[3] = rec.q_st; row[4] = rec.q_en;
                    row[5] = rec.r_st; row[6] = rec.r_en; row[7] = rec.mapq; row[8] = rec.cigar_len;
                    memcpy(cigar + ncig, res.fr.cigar + rec.cigar_off, (size_t)rec.cigar_len * 4);
                }
                ++nrec;
                ncig += rec.cigar_len;
            }
        *n_cigar_out = ncig;
        return nrec;
    }
    for (int64_t r = 0; r < n_reads; ++r)
        for (const vmg::Record &rec : res.records[r]) {
            if (nrec < rec_cap && ncig + (int64_t)rec.cigar.size() <= cigar_cap) {
                int64_t *row = rec_rows + nrec * 9;
                row[0] = r; row[1] = rec.contig; row[2] = rec.strand; row[3] = rec.q_st; row[4] = rec.q_en;
                row[5] = rec.r_st; row[6] = rec.r_en; row[7] = rec.mapq; row[8] = (int64_t)rec.cigar.size();
                memcpy(cigar + ncig, rec.cigar.data(), rec.cigar.size() * 4);
            }
            ++nrec;
            ncig += (int64_t)rec.cigar.size();
        }
    *n_cigar_out = ncig;
    return nrec;
}
}
