// Per-read glue of the extension stage as allocation-free functions over flat arrays, callable from CUDA kernels
// (vm_dglue.cu: the product) and from plain C++ (tests/gluetest: the same functions run in host loops with the
// oracle's C natives standing in for the kernels, so the logic is checked on a CPU-only box).
//
// Mirrors, for one read, everything extend_func does between its hot loops (mammap_clrnano.py:19238-19303):
// the divergence filter's bookkeeping (:19246-19256), extend_edge_test (:2302-2525, literally sequential here:
// left then right extension of every sub-alignment in order), drop_misplaced_alignment_test (:726-786),
// merge_conjacent_alignment (:16736-16780) with getdupiloc_numba (:16680-16734), fix_simple_inv (:24226-24312),
// split_alignment_test's segment selection (:21505-21617), get_onemapinfolist (:20731-20838) and pairedindel
// (:5604-5650).  vm_glue.hpp holds the older vector-based host versions (still used by the front half).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define VM_HD __host__ __device__ __forceinline__
#else
#define VM_HD inline
#endif

namespace vmd {

struct A32 { int32_t x; uint32_t y; int32_t s, l; };       // == VmAnchor (kernel storage)
struct Anc { int64_t x, y; int32_t s, l; };                // working form: 64-bit coordinates

VM_HD Anc widen(const A32 &a) { Anc r; r.x = a.x; r.y = (int64_t)a.y; r.s = a.s; r.l = a.l; return r; }
VM_HD Anc mk(int64_t x, int64_t y, int32_t s, int32_t l) { Anc r; r.x = x; r.y = y; r.s = s; r.l = l; return r; }
VM_HD int64_t iabs(int64_t v) { return v < 0 ? -v : v; }
VM_HD int64_t imin(int64_t a, int64_t b) { return a < b ? a : b; }
VM_HD int64_t imax(int64_t a, int64_t b) { return a > b ? a : b; }

// one side of an alignment job (== VmSeqSpec)
struct Spec { int64_t lo; int32_t len, src, reverse, comp; };
// == VmAlnJobDev
struct Job {
    Spec t, q;
    int32_t read, n_out;
    int64_t out_off, dir_off, sc_off, result0, result1;
};

struct RebuildRec { long long anc_off, len_off; int32_t n_anc, n_al; };    // == VmRebuildRec
struct ExtractRec { long long anc_off, meta_off; int32_t n_anc, n_chains; }; // == VmExtractRec

// one sub-alignment of a read while extend_func works on it
struct Sub {
    int64_t anc_off;     // its anchors in the rebuilt anchor array
    int32_t n_anc;
    int32_t alive;
    Anc first, last;     // current boundary anchors (rewritten to zero-length points by the extensions)
};

// one final sub-alignment (after merge / inversion fix / split): what a record is made of
struct Fin {
    Anc front, back;     // kept.front(), kept.back() of split_alignment
    int64_t job_lo;      // its fill jobs, in CIGAR order: [job_lo, job_lo + n_jobs)
    int32_t n_jobs;
    int32_t pad;
};

// per-read result header after the record stage
struct ReadOut {
    int32_t n_rec, second;       // records; 1: the read asks for the second pass (:24079-24080)
    int64_t n_ops;               // CIGAR ops of all its records (clips and tail M included)
    int64_t fin_lo;              // its Fin entries: [fin_lo, fin_lo + n_rec)
};

enum { ST_OK = 0, ST_FEW_ANCHORS = 1, ST_LOW_SCORE = 2, ST_SHORT_LOCAL = 3, ST_DROPPED = 4, ST_NO_RECORDS = 5, ST_FAILED = 6 };
enum { CT_DROP_MISPLACED = 0, CT_MERGE_CONJACENT, CT_FIX_SIMPLE_INV, CT_SECOND_PASS, CT_N_JOBS, CT_CIG_SCRATCH, CT_OPEN_ED, CT_FA, CT_COUNT = 8 };

struct Ctg {
    const int64_t *start, *len;
    int32_t n;
    // pos2contig :51-59 -- last contig whose start <= pos (the first one if pos precedes all)
    VM_HD int cid(int64_t pos) const
    {
        int lo = 0, hi = n;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (start[mid] <= pos) lo = mid + 1; else hi = mid;
        }
        return lo > 0 ? lo - 1 : 0;
    }
    // Python slice contig[a:b] -> [lo, hi) in GLOBAL coordinates
    VM_HD void slice(int c, int64_t a, int64_t b, int64_t &lo, int64_t &hi) const
    {
        const int64_t n_ = len[c];
        if (a < 0) a = imax(a + n_, 0);
        if (b < 0) b = imax(b + n_, 0);
        a = imin(a, n_);
        b = imin(b, n_);
        if (b < a) b = a;
        lo = start[c] + a;
        hi = start[c] + b;
    }
};

VM_HD void pyslice(int64_t n, int64_t a, int64_t b, int64_t &lo, int64_t &hi)
{
    if (a < 0) a = imax(a + n, 0);
    if (b < 0) b = imax(b + n, 0);
    a = imin(a, n);
    b = imin(b, n);
    if (b < a) b = a;
    lo = a; hi = b;
}

VM_HD Spec spec(int src, int64_t lo, int64_t hi, int reverse, int comp, bool need_reverse)
{
    Spec s;
    // after need_reverse the per-read driver swaps testseq / rc_testseq (:24063-24065)
    if (need_reverse && src != 0) src = 3 - src;
    s.lo = lo; s.len = (int32_t)(hi - lo); s.src = src; s.reverse = reverse; s.comp = comp;
    return s;
}

// get_query_target_for_cigar :5802-5818
VM_HD void query_target(const Anc &pre, const Anc &now, int64_t L, const Ctg &ctg, bool need_reverse, Spec &target, Spec &query)
{
    int64_t tlo, thi, qlo, qhi;
    if (pre.s == 1) {
        const int c = ctg.cid(pre.y);
        const int64_t b = ctg.start[c];
        ctg.slice(c, pre.y - b, now.y - b, tlo, thi);
        pyslice(L, pre.x, now.x, qlo, qhi);
        query = spec(1, qlo, qhi, 0, 0, need_reverse);
    } else {
        const int c = ctg.cid(now.y);
        const int64_t b = ctg.start[c];
        ctg.slice(c, now.y + now.l - b, pre.y + pre.l - b, tlo, thi);
        pyslice(L, L - now.x, L - pre.x, qlo, qhi);
        query = spec(2, qlo, qhi, 0, 0, need_reverse);
    }
    target = spec(0, tlo, thi, 0, 0, need_reverse);
}

// largest distance d that still passes `d / minlen > maxdivergence` (:19252-19254)
VM_HD int64_t divergence_band(double maxdivergence, int64_t minlen)
{
    if (!(maxdivergence * (double)minlen < 1e9)) return -1;
    int64_t d = (int64_t)(maxdivergence * (double)minlen);      // floor for non-negative values
    while (d > 0 && (double)d / (double)minlen > maxdivergence) --d;
    while (!((double)(d + 1) / (double)minlen > maxdivergence)) ++d;
    return d > 0 ? d : 0;
}

struct ReadCtx {
    int32_t read;
    int64_t L;
    bool need_reverse;
    Ctg ctg;
};

// ---------------------------------------------------------------------------------------------------------------
// stage A: sub-alignment table + divergence-filter jobs of one read.  Returns the read's status (ST_OK: goes on).
// sub / jobs are indexed like the rebuilt sub-alignments (rec.len_off + i).
// ---------------------------------------------------------------------------------------------------------------
VM_HD int init_read(const ReadCtx &rc, int32_t local_cnt, const ExtractRec &xr, const RebuildRec &rec, const A32 *al_anc,
                    const int32_t *al_len, double maxdivergence, Sub *sub, Job *jobs)
{
    if (local_cnt <= 0) return ST_DROPPED;          // np.array([]) indexing raises in the reference
    if (xr.n_anc <= 1) return ST_SHORT_LOCAL;       // :24067-24068
    if (rec.n_al <= 0) return ST_DROPPED;           // rebuild_chain_break leaves nothing: the reference raises
    int64_t ao = rec.anc_off;
    int status = ST_OK;
    for (int i = 0; i < rec.n_al; ++i) {
        const int32_t len = al_len[rec.len_off + i];
        Sub s;
        s.anc_off = ao; s.n_anc = len; s.alive = 1;
        s.first = widen(al_anc[ao]);
        s.last = widen(al_anc[ao + len - 1]);
        sub[rec.len_off + i] = s;
        Job j;
        query_target(s.first, s.last, rc.L, rc.ctg, rc.need_reverse, j.t, j.q);
        const int64_t m = imin(j.t.len, j.q.len);
        j.read = rc.read;
        j.n_out = len;
        j.dir_off = ao;
        j.sc_off = 0; j.result0 = 0; j.result1 = 0;
        j.out_off = m > 0 ? divergence_band(maxdivergence, m) : -1;
        if (m <= 0) status = ST_DROPPED;            // division by zero in the divergence filter (:19251)
        jobs[rec.len_off + i] = j;
        ao += len;
    }
    if (status != ST_OK)
        for (int i = 0; i < rec.n_al; ++i) {        // neutral jobs: nothing to bound, nothing to filter
            Job &j = jobs[rec.len_off + i];
            j.t.len = 0; j.q.len = 0; j.n_out = 0; j.out_off = 1;
        }
    return status;
}

// ---------------------------------------------------------------------------------------------------------------
// stage B: divergence filter, edge extension, misplaced sub-alignments
// ---------------------------------------------------------------------------------------------------------------
VM_HD int next_alive(const Sub *s, int n, int i) { while (i < n && !s[i].alive) ++i; return i; }
VM_HD int prev_alive(const Sub *s, int i) { while (i >= 0 && !s[i].alive) --i; return i; }
VM_HD int count_alive(const Sub *s, int n) { int c = 0; for (int i = 0; i < n; ++i) c += s[i].alive != 0; return c; }

// extend_edge_test :2302-2525.  `ext(target, query, q_e, t_e)` runs mp.k_cigar(2,-4,4,4,4,4,bw=100,zdropvalue=50).
template <typename ExtFn>
VM_HD void extend_edge(const ReadCtx &rc, Sub *s, int n, const A32 *al_anc, ExtFn &ext)
{
    const int64_t max_extend = 20000, L = rc.L;
    const Ctg &ctg = rc.ctg;
    int prev = -1;
    for (int idx = next_alive(s, n, 0); idx < n; prev = idx, idx = next_alive(s, n, idx + 1)) {
        Sub &one = s[idx];
        const int nxt = next_alive(s, n, idx + 1);
        // ---- towards the read start ----
        if (one.first.x > 0) {
            int64_t looksize = prev < 0 ? one.first.x : one.first.x - (s[prev].last.x + s[prev].last.l);
            const Anc pre = one.first;
            const int c = ctg.cid(pre.y);
            const int64_t cs = ctg.start[c], clen = ctg.len[c];
            int64_t qlo, qhi, tlo, thi;
            int32_t q_e = 0, t_e = 0;
            if (pre.s == 1) {
                const int64_t target_st = pre.y, query_st = pre.x;
                looksize = imin(looksize, target_st - cs);
                if (looksize > max_extend) looksize = max_extend;
                if (looksize != 0) {
                    pyslice(L, imax(query_st - looksize, 0), query_st, qlo, qhi);
                    ctg.slice(c, target_st - cs - (qhi - qlo), target_st - cs, tlo, thi);
                    ext(spec(0, tlo, thi, 1, 0, rc.need_reverse), spec(1, qlo, qhi, 1, 0, rc.need_reverse), q_e, t_e);
                    one.first = mk(query_st - q_e, target_st - t_e, 1, 0);
                }
            } else {
                const int64_t target_en = pre.y + pre.l, query_st = pre.x;
                looksize = imin(looksize, cs + clen - (target_en - 1));
                if (looksize > max_extend) looksize = max_extend;
                if (looksize != 0) {
                    pyslice(L, imax(query_st - looksize, 0), query_st, qlo, qhi);
                    // revcomp(ref[target_en : target_en + len])[::-1] == the complement, in forward order
                    ctg.slice(c, target_en - cs, target_en + (qhi - qlo) - cs, tlo, thi);
                    ext(spec(0, tlo, thi, 0, 1, rc.need_reverse), spec(1, qlo, qhi, 1, 0, rc.need_reverse), q_e, t_e);
                    one.first = mk(query_st - q_e, target_en + t_e, -1, 0);
                }
            }
        } else {
            const Anc t = one.first;
            one.first = t.s == 1 ? mk(t.x, t.y, 1, 0) : mk(t.x, t.y + t.l, -1, 0);
        }
        // ---- towards the read end ----
        if (one.last.x + one.last.l < L) {
            int64_t looksize = nxt >= n ? L - (one.last.x + one.last.l) : s[nxt].first.x - (one.last.x + one.last.l);
            // preitem = onealignment[-2]: of a two-anchor sub-alignment that is the first anchor as the extension
            // towards the read start has just rewritten it (only its contig and strand are read)
            const Anc pre = one.n_anc == 2 ? one.first : widen(al_anc[one.anc_off + one.n_anc - 2]);
            const Anc now = one.last;
            const int c = ctg.cid(pre.y);
            const int64_t cs = ctg.start[c], clen = ctg.len[c];
            int64_t qlo, qhi, tlo, thi;
            int32_t q_e = 0, t_e = 0;
            if (pre.s == 1) {
                const int64_t target_en = now.y + now.l, query_en = now.x + now.l;
                looksize = imin(looksize, cs + clen - (target_en - 1));
                if (looksize > max_extend) looksize = max_extend;
                if (looksize != 0) {
                    pyslice(L, query_en, query_en + looksize, qlo, qhi);
                    ctg.slice(c, target_en - cs, target_en + (qhi - qlo) - cs, tlo, thi);
                    ext(spec(0, tlo, thi, 0, 0, rc.need_reverse), spec(1, qlo, qhi, 0, 0, rc.need_reverse), q_e, t_e);
                    one.last = mk(query_en + q_e, target_en + t_e, 1, 0);
                }
            } else {
                const int64_t target_st = now.y, query_en = now.x + now.l;
                looksize = imin(looksize, target_st - cs);
                if (looksize > max_extend) looksize = max_extend;
                if (looksize != 0) {
                    pyslice(L, query_en, query_en + looksize, qlo, qhi);
                    // revcomp(ref[target_st - len : target_st]): reversed and complemented
                    ctg.slice(c, target_st - cs - (qhi - qlo), target_st - cs, tlo, thi);
                    ext(spec(0, tlo, thi, 1, 1, rc.need_reverse), spec(1, qlo, qhi, 0, 0, rc.need_reverse), q_e, t_e);
                    one.last = mk(query_en + q_e, target_st - t_e, -1, 0);
                }
            }
        } else {
            const Anc t = one.last;
            one.last = t.s == 1 ? mk(t.x + t.l, t.y + t.l, 1, 0) : mk(t.x + t.l, t.y, -1, 0);
        }
    }
}

VM_HD void gaps_of(const Anc &pre, const Anc &now, int64_t &readgap, int64_t &refgap)
{
    readgap = now.x - pre.x - pre.l;
    refgap = pre.s == 1 ? now.y - pre.y - pre.l : pre.y - now.y - now.l;
}

// drop_misplaced_alignment_test :726-786 on the alive sub-alignments a < b < c; true: b is to be removed
VM_HD bool misplaced(const Sub &a, const Sub &b, const Sub &c)
{
    if (!(a.first.s == b.first.s && a.first.s == c.first.s)) return false;
    const int64_t mid = b.last.x + b.last.l - b.first.x;
    if (mid > 1000) return false;
    int64_t readgap, refgap;
    gaps_of(a.last, b.first, readgap, refgap);
    if (!(iabs(refgap) < 100000)) return false;
    int DEL = 0, INS = 0;
    if (readgap - refgap < -30) ++DEL;
    else if (readgap - refgap > 30) ++INS;
    else return false;
    const int64_t gap_1 = iabs(readgap - refgap);
    gaps_of(b.last, c.first, readgap, refgap);
    if (!(iabs(refgap) < 100000)) return false;
    if (readgap - refgap < -30) ++DEL;
    else if (readgap - refgap > 30) ++INS;
    else return false;
    const int64_t gap_2 = iabs(readgap - refgap);
    return DEL == 1 && INS == 1 && (mid < 500 || (double)imax(gap_1, gap_2) / (double)mid > 0.5);
}

// The extension part of extend_func for one read: `s` = its n sub-alignments with the divergence filter's
// distances in dist[] (exact, or an upper bound that is within the job's band).  Returns ST_OK or the status
// that ends the read; *filtered / *n_dropped report what drop_misplaced did.
template <typename ExtFn>
VM_HD int extend_read(const ReadCtx &rc, Sub *s, int n, const Job *jobs, const A32 *al_anc, double maxdivergence, bool nofilter,
                      ExtFn &ext, bool *filtered, int *n_dropped)
{
    *filtered = false;
    *n_dropped = 0;
    for (int i = 0; i < n; ++i) {
        const double ratio = (double)jobs[i].result0 / (double)imin(jobs[i].t.len, jobs[i].q.len);
        if (ratio > maxdivergence) s[i].alive = 0;
    }
    extend_edge(rc, s, n, al_anc, ext);
    const int n0 = count_alive(s, n);
    if (n0 > 2 && !nofilter) {
        int a = next_alive(s, n, 0);
        for (;;) {
            const int b = next_alive(s, n, a + 1);
            const int c = b < n ? next_alive(s, n, b + 1) : n;
            if (c >= n) break;
            if (misplaced(s[a], s[b], s[c])) { s[b].alive = 0; ++*n_dropped; }
            else a = b;
        }
    }
    if (*n_dropped > 0) {
        *filtered = true;
        extend_edge(rc, s, n, al_anc, ext);
    }
    return ST_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// stage C: merge_conjacent, fix_simple_inv, split_alignment -> final sub-alignments and their fill jobs
// ---------------------------------------------------------------------------------------------------------------
// A final sub-alignment while it is assembled: anchors fa[lo .. lo + n), with one spare slot before and after.
struct Seg { int64_t lo; int32_t n; int32_t pad; };

// getdupiloc_numba :16680-16734 on `m` sub-alignments given by their first / last anchors (keeps the
// `[0][2]` strand-for-length quirk); dup[i] = 1 when index i is in the returned list
VM_HD void getdupiloc(const Anc *first, const Anc *last, int m, uint8_t *dup)
{
    for (int i = 0; i < m; ++i) dup[i] = 0;
    if (m < 2) return;
    int iloc = 0;
    while (iloc + 1 < m) {
        const Anc la = last[iloc];
        const int64_t readpos_1 = la.x + la.l;
        const int64_t refpos_1 = la.s == 1 ? la.y + la.l : la.y;
        const int strand_1 = la.s == 1 ? 1 : -1;
        int jloc = iloc, new_iloc = 0;
        bool hit = false;
        int64_t dupsize = 0, readpos_2 = 0;
        while (jloc + 1 < m) {
            ++jloc;
            int64_t refpos_2;
            int strand_2;
            if (last[jloc].s == 1) { refpos_2 = first[jloc].y; strand_2 = 1; }
            else { refpos_2 = first[jloc].y + first[jloc].s; strand_2 = -1; }
            if (strand_1 != strand_2) continue;
            const int64_t d = strand_1 == 1 ? refpos_2 - refpos_1 : refpos_1 - refpos_2;
            if (d < 50) { new_iloc = jloc; dupsize = d; readpos_2 = first[jloc].x; hit = true; }
        }
        if (hit) {
            const int64_t readgap = readpos_2 - readpos_1;
            if ((iloc + 1) < new_iloc || ((dupsize - readgap) < -30 && readgap < 30))
                for (int q = iloc; q < new_iloc; ++q) dup[q] = 1;
            iloc = new_iloc;
        } else ++iloc;
    }
}

struct FinalizeOut {
    int32_t n_fin;        // final sub-alignments of the read
    int32_t status;       // ST_OK or what ended the read
    int32_t n_merged, n_fixinv;
};

// Job allocator: returns the base index of `n` consecutive fill-job slots and of `ops` CIGAR scratch words.
// merge_conjacent_alignment + fix_simple_inv + split_alignment_test for one read.
//   s[n]        the read's sub-alignments after stage B
//   al_anc      rebuilt anchors (A32)
//   fa          the read's slice of the final-anchor arena: room for n_anc + 2 * n entries
//   seg, first, last, dup   per-read scratch of n entries each
//   fin         output, n entries at most
//   read_seq    the ORIENTED read (what the reference calls testseq at that point), ref = concatenated reference
// alloc(n_jobs, scratch_words, &job_base, &scratch_base) claims room for the read's fill jobs.
template <typename AllocFn>
VM_HD FinalizeOut finalize_read(const ReadCtx &rc, const Sub *s, int n, const A32 *al_anc, Anc *fa, Seg *seg, Anc *first, Anc *last,
                                uint8_t *dup, const uint8_t *ref, const uint8_t *read_seq, Fin *fin, Job *jobs_out, AllocFn &alloc)
{
    FinalizeOut out;
    out.n_fin = 0; out.status = ST_OK; out.n_merged = 0; out.n_fixinv = 0;
    const Ctg &ctg = rc.ctg;
    const int64_t L = rc.L;
    // ---- the alive sub-alignments, in order ----
    int m = 0;
    for (int i = 0; i < n; ++i)
        if (s[i].alive) { first[m] = s[i].first; last[m] = s[i].last; ++m; }
    if (m == 0) { out.status = ST_NO_RECORDS; return out; }
    // ---- merge_conjacent_alignment :16736-16780 on (first, last) ----
    // The reference walks a shrinking list with index iloc; `dup` (computed once, on the list before any merge) is
    // asked about that CURRENT index.  Groups of consecutive sources are what its merges amount to:
    // seg[g].n = number of sources of group g.
    getdupiloc(first, last, m, dup);
    int n_grp = 1;
    {
        int iloc = 0;                // index of the open group in the reference's list
        Anc cur_last = last[0];
        seg[0].n = 1;
        for (int nxt = 1; nxt < m; ++nxt) {
            bool merge = false;
            if (!dup[iloc]) {
                const Anc pre = cur_last, now = first[nxt];
                if (pre.s == now.s && ctg.cid(pre.y) == ctg.cid(now.y)) {
                    int64_t readgap, refgap;
                    gaps_of(pre, now, readgap, refgap);
                    if (refgap >= 0 && imin(readgap, refgap) < 50 && iabs(readgap - refgap) < 10000) merge = true;
                }
            }
            if (merge) {
                seg[iloc].n += 1;
                ++out.n_merged;
            } else {
                ++iloc;
                seg[iloc].n = 1;
            }
            cur_last = last[nxt];
        }
        n_grp = iloc + 1;
    }
    // ---- materialise every group as [spare][anchors of its sources, back to back][spare] ----
    {
        int64_t w = 0;
        int sub_i = next_alive(s, n, 0);
        for (int g = 0; g < n_grp; ++g) {
            const int cnt = seg[g].n;
            const int64_t lo = w + 1;          // one spare slot in front
            int64_t k = lo;
            for (int q = 0; q < cnt; ++q) {
                const Sub &one = s[sub_i];
                fa[k++] = one.first;
                for (int t = 1; t + 1 < one.n_anc; ++t) fa[k++] = widen(al_anc[one.anc_off + t]);
                fa[k++] = one.last;
                sub_i = next_alive(s, n, sub_i + 1);
            }
            seg[g].lo = lo;
            seg[g].n = (int32_t)(k - lo);
            w = k + 1;                          // one spare slot behind
        }
    }
    // ---- fix_simple_inv :24226-24312 ----
    if (n_grp > 2) {
        for (int iloc = 0; iloc + 2 < n_grp; ++iloc) {
            Seg &A = seg[iloc], &B = seg[iloc + 1], &C = seg[iloc + 2];
            const Anc A0 = fa[A.lo], B0 = fa[B.lo], C0 = fa[C.lo];
            if (!(A0.s == C0.s && A0.s != B0.s && A0.s == 1)) continue;
            const Anc Ab = fa[A.lo + A.n - 1], Bb = fa[B.lo + B.n - 1];
            const int c = ctg.cid(A0.y);
            const int64_t bias0 = ctg.start[c];
            const int64_t refen_0 = Ab.y + Ab.l - bias0;
            const int64_t readen_0 = Ab.x + Ab.l;
            const int64_t refst_1 = Bb.y - bias0;
            const int64_t readst_1 = B0.x;
            const int64_t refen_1 = B0.y + B0.l - bias0;
            const int64_t readen_1 = Bb.x + Bb.l;
            const int64_t refst_2 = C0.y - bias0;
            const int64_t readst_2 = C0.x;
            if (!(refst_2 - refen_0 == refen_1 - refst_1 && readst_1 - readen_0 + readst_2 - readen_1 == 0)) continue;
            if (!(refst_1 - refen_0 != 0 && refst_1 - refen_0 + refst_2 - refen_1 == 0)) continue;
            int64_t rlo, rhi, qlo, qhi;
            if (refen_0 > refst_1) {
                ctg.slice(c, refen_1, refen_1 + refen_0 - refst_1, rlo, rhi);
                pyslice(L, readen_0 - refen_0 + refst_1, readen_0, qlo, qhi);
                bool same = (rhi - rlo) == (qhi - qlo);
                for (int64_t t = 0; same && t < rhi - rlo; ++t) {
                    const uint8_t r_ = ref[rhi - 1 - t];
                    const uint8_t cc = r_ == 'A' ? 'T' : r_ == 'T' ? 'A' : r_ == 'G' ? 'C' : r_ == 'C' ? 'G' : 'N';
                    if (cc != read_seq[qlo + t]) same = false;
                }
                if (same) {
                    ++out.n_fixinv;
                    const int64_t bias = refen_0 - refst_1;
                    fa[C.lo] = mk(readst_2 - bias, refst_2 - bias + bias0, 1, 0);
                    const Anc ins = mk(readst_2 - bias, refen_0 + bias0, -1, 0);
                    for (;;) {
                        if (B.n == 0) { out.status = ST_DROPPED; return out; }      // the reference raises (pop from empty list)
                        const Anc bb = fa[B.lo + B.n - 1];
                        if (ins.x <= bb.x + bb.l) --B.n;
                        else break;
                    }
                    fa[B.lo + B.n] = ins;
                    ++B.n;
                }
            } else {
                ctg.slice(c, refen_0, refst_1, rlo, rhi);
                pyslice(L, readen_0, readen_0 - refen_0 + refst_1, qlo, qhi);
                bool same = (rhi - rlo) == (qhi - qlo);
                for (int64_t t = 0; same && t < rhi - rlo; ++t)
                    if (ref[rlo + t] != read_seq[qlo + t]) same = false;
                if (same) {
                    ++out.n_fixinv;
                    fa[A.lo + A.n - 1] = mk(readen_0 - refen_0 + refst_1, refst_1 + bias0, 1, 0);
                    const Anc ins = mk(readen_0 - refen_0 + refst_1, refen_1 + refen_0 - refst_1 + bias0, -1, 0);
                    for (;;) {
                        if (B.n == 0) { out.status = ST_DROPPED; return out; }
                        if (ins.x >= fa[B.lo].x) { ++B.lo; --B.n; }
                        else break;
                    }
                    --B.lo;
                    fa[B.lo] = ins;
                    ++B.n;
                }
            }
        }
    }
    // ---- split_alignment_test :21505-21617: count the fill jobs, claim room, write them ----
    // pass 0 counts (jobs, scratch words), pass 1 writes
    int64_t job_base = 0, scratch_base = 0;
    for (int pass = 0; pass < 2; ++pass) {
        int64_t nj = 0, words = 0;
        for (int g = 0; g < n_grp; ++g) {
            const Seg &S = seg[g];
            const bool fwd = fa[S.lo].s == 1;
            // boundary rewrites (:21512-21516, :21575-21583) -- idempotent, done in pass 0
            if (pass == 0) {
                if (fwd) {
                    Anc t = fa[S.lo + S.n - 1];
                    if (t.l != 0) fa[S.lo + S.n - 1] = mk(t.x + t.l, t.y + t.l, 1, 0);
                } else {
                    Anc t = fa[S.lo];
                    if (t.l != 0) fa[S.lo] = mk(t.x, t.y + t.l, -1, 0);
                    t = fa[S.lo + S.n - 1];
                    if (t.l != 0) fa[S.lo + S.n - 1] = mk(t.x + t.l, t.y, -1, 0);
                }
            }
            const int64_t before = nj;
            Anc pre = fwd ? fa[S.lo] : fa[S.lo + S.n - 1];
            Anc kept_back = pre;
            for (int iloc = 1; iloc < S.n; ++iloc) {
                const Anc now = fwd ? fa[S.lo + iloc] : fa[S.lo + S.n - 1 - iloc];
                int64_t readgap, refgap;
                if (fwd) { readgap = now.x - pre.x - pre.l; refgap = now.y - pre.y - pre.l; }
                else { readgap = pre.x - now.x - now.l; refgap = now.y - pre.y - pre.l; }
                if ((now.l < 19 || imin(readgap, refgap) < 200) && iloc + 1 != S.n) continue;
                Spec t_, q_;
                if (fwd) query_target(pre, now, L, ctg, rc.need_reverse, t_, q_);
                else query_target(now, pre, L, ctg, rc.need_reverse, t_, q_);
                if (!(t_.len > 0 && q_.len > 0)) { out.status = ST_DROPPED; return out; }      // "Failed to compute CIGAR" :21559-21569
                if (pass == 1) {
                    Job j;
                    j.t = t_; j.q = q_; j.read = rc.read; j.n_out = 0;
                    j.out_off = scratch_base + words;
                    j.dir_off = 0; j.sc_off = 0; j.result0 = 0; j.result1 = 0;
                    jobs_out[job_base + nj] = j;
                }
                ++nj;
                words += (int64_t)t_.len + q_.len + 2;
                kept_back = now;
                pre = now;
            }
            if (nj == before) { out.status = ST_DROPPED; return out; }      // cigarlist[-1] == [] -> the record assembly raises
            if (pass == 1) {
                Fin f;
                f.front = fwd ? fa[S.lo] : fa[S.lo + S.n - 1];
                f.back = kept_back;
                f.job_lo = job_base + before;
                f.n_jobs = (int32_t)(nj - before);
                f.pad = 0;
                fin[g] = f;
            }
        }
        if (pass == 0 && !alloc(nj, words, job_base, scratch_base)) { out.status = ST_FAILED; return out; }
    }
    out.n_fin = n_grp;
    return out;
}

// ---------------------------------------------------------------------------------------------------------------
// stage D: get_onemapinfolist :20731-20838 -- record fields, CIGAR length check, pairedindel
// ---------------------------------------------------------------------------------------------------------------
struct Rec {            // == vm_record
    int32_t contig, strand;
    int64_t q_st, q_en, r_st, r_en;
    int32_t mapq, cigar_len;
    int64_t cigar_off;
};

VM_HD void record_fields(const ReadCtx &rc, const Fin &f, int mapq, Rec &r, int64_t &tailM)
{
    const Anc a0 = f.front, ab = f.back;
    r.contig = rc.ctg.cid(a0.y);
    const int64_t bias = rc.ctg.start[r.contig];
    r.mapq = mapq;
    tailM = 0;
    if (a0.s == 1) {
        r.q_st = a0.x;
        r.q_en = ab.x + ab.l;
        r.r_st = a0.y - bias;
        r.r_en = ab.y + ab.l - bias;
        if (ab.l > 0) tailM = ab.l;
        r.strand = rc.need_reverse ? -1 : 1;
    } else {
        r.q_st = rc.L - a0.x - a0.l;
        r.q_en = rc.L - ab.x;
        r.r_st = a0.y - bias;
        r.r_en = ab.y + ab.l - bias;
        r.strand = rc.need_reverse ? 1 : -1;
    }
}

// results[j] = (offset, length) of fill job j's ops in the dense op arena `ops`
struct U2 { uint32_t x, y; };

// Counts the read's records and CIGAR ops, checks every record's query length (:20777-20786) and evaluates
// pairedindel (:5604-5650) over all its CIGARs.  Returns ST_OK / ST_DROPPED / ST_NO_RECORDS.
VM_HD int count_read(const ReadCtx &rc, const Fin *fin, int n_fin, const U2 *results, const uint32_t *ops, bool hardclip,
                     int64_t *n_ops_out, bool *paired)
{
    int64_t total = 0;
    *paired = false;
    // indels > 30 of all records; pairedindel is true iff two of them have min / max > 0.7 (adjacent values of the
    // sorted list have the largest ratios, so "some adjacent pair" == "some pair")
    const int CAP = 48;
    double big[CAP];
    int nbig = 0;
    bool overflow = false;
    for (int i = 0; i < n_fin; ++i) {
        Rec r;
        int64_t tailM;
        record_fields(rc, fin[i], 0, r, tailM);
        int64_t qlen = 0, nops = 0;
        if (r.q_st > 0) { ++nops; if (!hardclip) qlen += r.q_st; }
        for (int j = 0; j < fin[i].n_jobs; ++j) {
            const U2 res = results[fin[i].job_lo + j];
            const uint32_t *p = ops + res.x;
            for (uint32_t t = 0; t < res.y; ++t) {
                const uint32_t o = p[t], op = o & 0xf, ln = o >> 4;
                if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) qlen += ln;
                if ((op == 1 || op == 2) && (double)ln > 30.0) {
                    if (nbig < CAP) big[nbig++] = (double)ln; else overflow = true;
                }
            }
            nops += res.y;
        }
        if (tailM > 0) { ++nops; qlen += tailM; }
        if (rc.L - r.q_en > 0) { ++nops; if (!hardclip) qlen += rc.L - r.q_en; }
        const int64_t want = hardclip ? (r.q_en - r.q_st) : rc.L;
        if (want != qlen) return ST_DROPPED;                    // Cigar length mismatch: the reference raises
        total += nops;
    }
    if (n_fin == 0) return ST_NO_RECORDS;
    for (int a = 0; a < nbig && !*paired; ++a)
        for (int b = a + 1; b < nbig; ++b) {
            const double lo = big[a] < big[b] ? big[a] : big[b], hi = big[a] < big[b] ? big[b] : big[a];
            if (hi > 0 && lo / hi > 0.7) { *paired = true; break; }
        }
    if (overflow && !*paired) {
        // more large indels than the local buffer holds: compare every op against every later one (rare)
        for (int i = 0; i < n_fin && !*paired; ++i)
            for (int j = 0; j < fin[i].n_jobs && !*paired; ++j) {
                const U2 ra = results[fin[i].job_lo + j];
                for (uint32_t t = 0; t < ra.y && !*paired; ++t) {
                    const uint32_t o = ops[ra.x + t], op = o & 0xf;
                    if (!((op == 1 || op == 2) && (double)(o >> 4) > 30.0)) continue;
                    const double va = (double)(o >> 4);
                    for (int i2 = i; i2 < n_fin && !*paired; ++i2)
                        for (int j2 = (i2 == i ? j : 0); j2 < fin[i2].n_jobs && !*paired; ++j2) {
                            const U2 rb = results[fin[i2].job_lo + j2];
                            for (uint32_t t2 = (i2 == i && j2 == j ? t + 1 : 0); t2 < rb.y; ++t2) {
                                const uint32_t o2 = ops[rb.x + t2], op2 = o2 & 0xf;
                                if (!((op2 == 1 || op2 == 2) && (double)(o2 >> 4) > 30.0)) continue;
                                const double vb = (double)(o2 >> 4);
                                const double lo = va < vb ? va : vb, hi = va < vb ? vb : va;
                                if (hi > 0 && lo / hi > 0.7) { *paired = true; break; }
                            }
                        }
                }
            }
    }
    *n_ops_out = total;
    return ST_OK;
}

// Writes one record of a read (fields + CIGAR ops at cig[cig_off ..)); returns the number of ops.  `lane` of `nl`
// cooperating threads copies every nl-th op (1 thread: lane 0 of 1); all of them must call with the same arguments.
// With need_reverse the reference reverses the record list (:20836-20838): the caller picks the slot.
VM_HD int64_t write_record(const ReadCtx &rc, const Fin &f, int mapq, const U2 *results, const uint32_t *ops, bool hardclip,
                           Rec *out, uint32_t *cig, int64_t cig_off, int lane, int nl)
{
    Rec r;
    int64_t tailM;
    record_fields(rc, f, mapq, r, tailM);
    const uint32_t clip = hardclip ? 5u : 4u;
    int64_t k = cig_off;
    if (r.q_st > 0) { if (lane == 0) cig[k] = (uint32_t)r.q_st << 4 | clip; ++k; }
    for (int j = 0; j < f.n_jobs; ++j) {
        const U2 res = results[f.job_lo + j];
        for (uint32_t t = (uint32_t)lane; t < res.y; t += (uint32_t)nl) cig[k + t] = ops[res.x + t];
        k += res.y;
    }
    if (tailM > 0) { if (lane == 0) cig[k] = (uint32_t)tailM << 4 | 0u; ++k; }
    if (rc.L - r.q_en > 0) { if (lane == 0) cig[k] = (uint32_t)(rc.L - r.q_en) << 4 | clip; ++k; }
    r.cigar_off = cig_off;
    r.cigar_len = (int32_t)(k - cig_off);
    if (lane == 0) *out = r;
    return k - cig_off;
}

} // namespace vmd
