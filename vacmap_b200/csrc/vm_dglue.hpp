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

// ===============================================================================================================
// Front half: what hit2work_1 does after its DP (:23581-23734: chain grouping, MAPQ inputs, secondary chains),
// guide-chain selection (merge_chain / drop_somechains / sort, :28482-28582) and the inputs of the re-seeding
// kernel for every guide chain (:23090-23191: windows, guide points) -- for one read, from the chains the
// extraction kernel left in HBM.  vm_glue.hpp::hit2work_extracted / select_guides / make_guide_job are the
// vector-based host versions of the same.
// ===============================================================================================================
namespace vmd {

// numba's argsort (numba/misc/quicksort.py) replayed on R[0..n): the permutation, ties and all.
template <typename K>
VM_HD void argsort_replay(const K *A, int32_t n, int32_t *R)
{
    for (int32_t t = 0; t < n; ++t) R[t] = t;
    if (n < 2) return;
    {
        bool asc = true, desc = true;
        for (int32_t t = 1; t < n && (asc || desc); ++t) {
            if (!(A[t - 1] < A[t])) asc = false;
            if (!(A[t] < A[t - 1])) desc = false;
        }
        if (asc) return;
        if (desc) { for (int32_t t = 0; t < n / 2; ++t) { const int32_t x = R[t]; R[t] = R[n - 1 - t]; R[n - 1 - t] = x; } return; }
    }
    int32_t slo[64], shi[64];
    int sp = 1;
    slo[0] = 0; shi[0] = n - 1;
    while (sp > 0) {
        --sp;
        int32_t low = slo[sp], high = shi[sp];
        while (high - low >= 15) {
            const int32_t mid = (low + high) >> 1;
            int32_t x;
            if (A[R[mid]] < A[R[low]]) { x = R[low]; R[low] = R[mid]; R[mid] = x; }
            if (A[R[high]] < A[R[mid]]) { x = R[high]; R[high] = R[mid]; R[mid] = x; }
            if (A[R[mid]] < A[R[low]]) { x = R[low]; R[low] = R[mid]; R[mid] = x; }
            const K pivot = A[R[mid]];
            x = R[high]; R[high] = R[mid]; R[mid] = x;
            int32_t i = low, j = high - 1;
            for (;;) {
                while (i < high && A[R[i]] < pivot) ++i;
                while (j >= low && pivot < A[R[j]]) --j;
                if (i >= j) break;
                x = R[i]; R[i] = R[j]; R[j] = x;
                ++i; --j;
            }
            x = R[i]; R[i] = R[high]; R[high] = x;
            if (high - i > i - low) {
                if (high > i) { slo[sp] = i + 1; shi[sp] = high; ++sp; }
                high = i - 1;
            } else {
                if (i > low) { slo[sp] = low; shi[sp] = i - 1; ++sp; }
                low = i + 1;
            }
        }
        for (int32_t i = low + 1; i <= high; ++i) {
            const int32_t k = R[i];
            const K v = A[k];
            int32_t j = i;
            while (j > low && v < A[R[j - 1]]) { R[j] = R[j - 1]; --j; }
            R[j] = k;
        }
    }
}

// == VmReseedJobDev
struct RJob {
    int32_t read, need_reverse, readstart, readend, n_win, n_guide;
    int64_t win_off, g_off, hit_off;
    int32_t hit_cap, pad0;
    int64_t dense_off, tab_off;
    int32_t tab_size, pad;
};

struct FrontOut {
    int32_t status;        // ST_OK: the read goes on to the local stage
    int32_t n_guides;      // guide chains after merge_chain / drop_somechains (> 1: the multi-chain local DP)
    int32_t n_jobs;        // of them re-seeded (the mode's cap)
    int32_t pad;
    double f1, f2, m;      // MAPQ inputs (:23697-23704); the host takes the logarithm (libm, as numba does)
};

// per-read scratch, every pointer already offset to the read's slice
struct FrontScratch {
    // sized by the read's extracted chains (nc)
    int32_t *order, *coff /* nc + 1 */, *pstart /* nc + 1 */, *rest, *head, *tail, *nxt, *size, *cseg, *cpos, *cflat, *sc0, *sc1, *ordc;
    int64_t *dist, *k64c;
    double *kd;
    // sized by the read's extracted anchors (na)
    int32_t *bins, *cur, *orda;
    int64_t *k64a;
    Anc *tmp;
};

// a (possibly merged) guide chain = original chains linked head -> nxt -> ... -> tail, each in descending read order
struct ChainView {
    const A32 *anc;            // the read's extracted anchors
    const int32_t *coff, *nxt;
    VM_HD Anc front(const int32_t *head, int c) const { return widen(anc[coff[head[c]]]); }
    VM_HD Anc back(const int32_t *tail, int c) const { return widen(anc[coff[tail[c] + 1] - 1]); }
};

// anchor number `idx` of the chain starting at original chain `seg` -- sequential cursor
struct Cursor {
    int32_t seg, pos;          // original chain, offset inside it
};

VM_HD int64_t front_read(const ReadCtx &rc, const ExtractRec &xr, const A32 *anc_all, const double *S_all, const int32_t *len_all,
                         const double *score_all, int max_guides, int kmer, const FrontScratch &W, RJob *jobs, int64_t *wlo, int64_t *whi,
                         int32_t *gx, int64_t *gy, int64_t g_base, int64_t w_base, FrontOut &out)
{
    out.status = ST_LOW_SCORE; out.n_guides = 0; out.n_jobs = 0; out.pad = 0; out.f1 = 0; out.f2 = 0; out.m = 0;
    const int nc = xr.n_chains;
    if (nc <= 0) return 0;
    const A32 *anc = anc_all + xr.anc_off;
    const double *S_prim = S_all + xr.anc_off;
    const int32_t *len = len_all + xr.meta_off;
    const double *score = score_all + xr.meta_off;
    int32_t *coff = W.coff;
    coff[0] = 0;
    for (int c = 0; c < nc; ++c) coff[c + 1] = coff[c] + len[c];
    // ---- order of the chains: descending score, the primary chain first (:23655-23662) ----
    int32_t *order = W.order;
    argsort_replay<double>(score, nc, order);
    for (int t = 0; t < nc / 2; ++t) { const int32_t x = order[t]; order[t] = order[nc - 1 - t]; order[nc - 1 - t] = x; }
    if (order[0] != 0)
        for (int i = 0; i < nc; ++i)
            if (order[i] == 0) { order[i] = order[0]; order[0] = 0; break; }
    // ---- grouping by 100-bp read-bin overlap (:23672-23694): only the second score of the primary group is used ----
    // bin set of chain c: its anchors walked backwards (ascending read position)
    int32_t *bins = W.bins, *pstart = W.pstart, *cur = W.cur;
    int n_sets = 0;
    double f2 = 0.0;
    bool have_f2 = false;
    auto binset = [&](int c, int32_t *dst) -> int {
        int n = 0;
        bool sorted = true;
        for (int t = coff[c + 1] - 1; t >= coff[c]; --t) {
            const int32_t b = anc[t].x / 100;
            if (n > 0 && b < dst[n - 1]) sorted = false;
            if (n == 0 || b != dst[n - 1]) dst[n++] = b;
        }
        if (!sorted) {          // never for chains out of the DP (strictly descending read positions); kept for safety
            for (int i = 1; i < n; ++i) { const int32_t v = dst[i]; int j = i; while (j > 0 && dst[j - 1] > v) { dst[j] = dst[j - 1]; --j; } dst[j] = v; }
            int u = 0;
            for (int i = 0; i < n; ++i) if (u == 0 || dst[i] != dst[u - 1]) dst[u++] = dst[i];
            n = u;
        }
        return n;
    };
    pstart[0] = 0;
    if (nc > 1) pstart[1] = binset(order[0], bins);
    else pstart[1] = 0;
    n_sets = 1;
    for (int oi = 1; oi < nc; ++oi) {
        const int c = order[oi];
        const int nb = binset(c, cur);
        double best = 0.0;
        int pref = 0;
        for (int p = 0; p < n_sets; ++p) {
            const int32_t *ps = bins + pstart[p];
            const int np = pstart[p + 1] - pstart[p];
            int inter = 0;
            if (nb > 0 && np > 0 && cur[0] <= ps[np - 1] && ps[0] <= cur[nb - 1]) {
                int i = 0, j = 0;
                while (i < nb && j < np) {
                    if (cur[i] < ps[j]) ++i;
                    else if (ps[j] < cur[i]) ++j;
                    else { ++inter; ++i; ++j; }
                }
            }
            const double ov = (double)inter / (double)(np < nb ? np : nb);
            if (ov > best) { best = ov; pref = p; }
        }
        if (best < 0.5) {
            for (int i = 0; i < nb; ++i) bins[pstart[n_sets] + i] = cur[i];
            pstart[n_sets + 1] = pstart[n_sets] + nb;
            ++n_sets;
        } else if (pref == 0 && !have_f2) { f2 = score[c]; have_f2 = true; }
    }
    out.f1 = score[order[0]];
    out.f2 = f2;
    out.m = (double)len[order[0]];
    // ---- select_secondary_alignment :23505-23538 ----
    int32_t *rest = W.rest;      // chain ids: the secondary chains in selection order
    int n_sec = 0;
    if (nc > 1) {
        const int np = len[0];
        auto loc2score_at = [&](int64_t q) -> double {
            int lo = 0, hi = np;          // first t with prim[t].x <= q (descending read order)
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if ((int64_t)anc[mid].x <= q) hi = mid; else lo = mid + 1;
            }
            return lo < np ? S_prim[lo] : 0.0;
        };
        for (int oi = 1; oi < nc; ++oi) {
            const int c = order[oi];
            const double f2s = score[c];
            const int64_t en_loc = anc[coff[c]].x, st_loc = anc[coff[c + 1] - 1].x;
            if (en_loc - st_loc < 50) continue;
            double f1s = loc2score_at(en_loc) - loc2score_at(st_loc);
            if (f1s < 1.0) f1s = 1.0;
            const double df = f1s - f2s;
            if (f2s / f1s > 0.9 || (df < 0 ? -df : df) < 40) {
                bool skip = false;
                for (int q = 0; q < n_sec; ++q) {
                    const int pc = rest[q];
                    const int64_t pe = anc[coff[pc]].x, ps = anc[coff[pc + 1] - 1].x;
                    const int64_t ovs = imax(imin(en_loc, pe) - imax(ps, st_loc), 0);
                    if ((double)ovs / (double)(en_loc - st_loc) > 0.5) { skip = true; break; }
                }
                if (!skip) rest[n_sec++] = c;
            }
        }
    }
    // ---- merge_chain :28529-28569 on the secondary chains ----
    int32_t *head = W.head, *tail = W.tail, *nxt = W.nxt, *size = W.size;
    for (int c = 0; c < nc; ++c) { head[c] = c; tail[c] = c; nxt[c] = -1; size[c] = len[c]; }
    ChainView cv{anc, coff, nxt};
    int n_rest = n_sec;
    if (n_rest > 0) {
        // sort by the read position of the last anchor (numba argsort on int64 keys)
        for (int q = 0; q < n_rest; ++q) W.k64c[q] = cv.back(tail, rest[q]).x;
        argsort_replay<int64_t>(W.k64c, n_rest, W.ordc);
        for (int q = 0; q < n_rest; ++q) W.cseg[q] = rest[W.ordc[q]];
        for (int q = 0; q < n_rest; ++q) rest[q] = W.cseg[q];
        int iloc = 0;
        while (iloc + 1 < n_rest) {
            int jloc = iloc + 1;
            while (jloc < n_rest) {
                const int ci = rest[iloc], cj = rest[jloc];
                const Anc a0 = cv.front(head, ci), bl = cv.back(tail, cj);
                if (a0.x + a0.l <= bl.x && a0.s == bl.s) {
                    const int64_t readgap = bl.x - a0.x - a0.l;
                    const int64_t refgap = a0.s == 1 ? bl.y - a0.y - a0.l : a0.y - bl.y - bl.l;
                    if (iabs(readgap - refgap) < 500) {
                        // merged = rest[jloc] followed by rest[iloc]; it takes rest[iloc]'s place
                        nxt[tail[cj]] = head[ci];
                        head[ci] = head[cj];
                        size[ci] += size[cj];
                        for (int q = jloc; q + 1 < n_rest; ++q) rest[q] = rest[q + 1];
                        --n_rest;
                        continue;
                    }
                }
                ++jloc;
            }
            ++iloc;
        }
        // sort by length
        for (int q = 0; q < n_rest; ++q) W.k64c[q] = size[rest[q]];
        argsort_replay<int64_t>(W.k64c, n_rest, W.ordc);
        for (int q = 0; q < n_rest; ++q) W.cseg[q] = rest[W.ordc[q]];
        for (int q = 0; q < n_rest; ++q) rest[q] = W.cseg[q];
    }
    // ---- drop_somechains :28482-28528 (chains[0] = the primary chain) ----
    // Note: tail[] of a merged chain is the tail of its LAST part, kept in tail[ci] only for unmerged chains; a merged
    // chain's last part is found by walking nxt (merges are rare)
    auto last_part = [&](int c) { int s = head[c]; while (nxt[s] >= 0) s = nxt[s]; return s; };
    if (n_rest > 0) {
        for (int q = 0; q < n_rest; ++q) {
            W.cseg[q] = head[rest[q]]; W.cpos[q] = 0; W.cflat[q] = 0; W.sc0[q] = 0; W.sc1[q] = 0; W.dist[q] = INT64_MAX;
        }
        for (int t = coff[0]; t < coff[1]; ++t) {
            const Anc item = widen(anc[t]);
            for (int q = 0; q < n_rest; ++q) {
                const int c = rest[q];
                const int lp = last_part(c);
                const int64_t back_x = anc[coff[lp + 1] - 1].x, front_x = anc[coff[head[c]]].x;
                if (item.x >= back_x && item.x <= front_x) { if (item.s == 1) W.sc0[q]++; else W.sc1[q]++; }
                for (;;) {
                    const A32 ca = anc[coff[W.cseg[q]] + W.cpos[q]];
                    if (!((int64_t)ca.x > item.x)) break;
                    if (W.cflat[q] < size[c] - 1) {
                        ++W.cflat[q];
                        if (++W.cpos[q] >= len[W.cseg[q]]) { W.cseg[q] = nxt[W.cseg[q]]; W.cpos[q] = 0; }
                    } else break;
                }
                const int64_t d = iabs(item.y - (int64_t)anc[coff[W.cseg[q]] + W.cpos[q]].y);
                if (d < W.dist[q]) W.dist[q] = d;
            }
        }
        int kept = 0;
        for (int q = 0; q < n_rest; ++q) {
            const int c = rest[q];
            int64_t c0 = 0, c1 = 0;
            for (int s = head[c]; s >= 0; s = nxt[s])
                for (int t = coff[s]; t < coff[s + 1]; ++t) { if (anc[t].s == 1) ++c0; else ++c1; }
            const bool keep = (W.sc0[q] > W.sc1[q] && c0 > c1) || (W.sc0[q] < W.sc1[q] && c0 < c1);
            const int lp = last_part(c);
            const int64_t span = (int64_t)anc[coff[head[c]]].x - (int64_t)anc[coff[lp + 1] - 1].x;
            if ((!keep && W.dist[q] < 500) || span < 100) continue;
            rest[kept++] = c;
        }
        n_rest = kept;
    }
    // ---- order of re-seeding: sort all chains by 1 / length (:28574), cap (:28575-28582) ----
    const int n_all = 1 + n_rest;
    // chain list: index 0 = primary (original chain 0), then rest[]
    for (int q = 0; q < n_all; ++q) W.kd[q] = 1.0 / (double)(q == 0 ? size[0] : size[rest[q - 1]]);
    argsort_replay<double>(W.kd, n_all, W.ordc);
    int used = n_all;
    if (max_guides > 0 && used > max_guides) used = max_guides;
    out.status = ST_OK;
    out.n_guides = n_all;
    out.n_jobs = used;
    // ---- one re-seeding job per used chain (:23090-23191) ----
    const int64_t look_span = 7000;
    int64_t g_used = 0, w_used = 0;
    for (int ji = 0; ji < used; ++ji) {
        const int li = W.ordc[ji];
        const int c = li == 0 ? 0 : rest[li - 1];
        const int n = size[c];
        // materialise the chain (descending read order) into tmp
        Anc *ch = W.tmp;
        {
            int k = 0;
            for (int s = head[c]; s >= 0; s = nxt[s])
                for (int t = coff[s]; t < coff[s + 1]; ++t) ch[k++] = widen(anc[t]);
        }
        int64_t readgap = 0;
        bool x_desc = true, y_asc = true, y_desc = true;
        for (int i = 1; i < n; ++i) {
            const int64_t d = iabs(ch[i].x - ch[i - 1].x);
            if (d > readgap) readgap = d;
            if (!(ch[i].x < ch[i - 1].x)) x_desc = false;
            if (!(ch[i - 1].y < ch[i].y)) y_asc = false;
            if (!(ch[i].y < ch[i - 1].y)) y_desc = false;
        }
        readgap = imax(readgap + 1000, 5000);
        int32_t *jgx = gx + g_used;
        int64_t *jgy = gy + g_used;
        int64_t *ys = W.k64a;             // reference positions ascending (the :23103 argsort)
        if (n >= 1 && x_desc && (y_asc || y_desc)) {
            for (int i = 0; i < n; ++i) {
                ys[i] = y_asc ? ch[i].y : ch[n - 1 - i].y;
                jgx[i] = (int32_t)ch[n - 1 - i].x;
                jgy[i] = ch[n - 1 - i].y;
            }
        } else {
            // general case: argsort by reference position, then by read position (:23183), both numba's
            int64_t *keys = (int64_t *)jgy;          // the job's gy room doubles as key scratch until it is written
            for (int i = 0; i < n; ++i) keys[i] = ch[i].y;
            argsort_replay<int64_t>(keys, n, W.orda);
            // by_y into the bins / cur scratch is too small for anchors: permute through ys + a second pass
            for (int i = 0; i < n; ++i) ys[i] = ch[W.orda[i]].y;
            // keys for the second sort: read positions of by_y
            for (int i = 0; i < n; ++i) keys[i] = ch[W.orda[i]].x;
            int32_t *ord2 = W.cur;               // na-sized int32 scratch (the bin sets are done)
            argsort_replay<int64_t>(keys, n, ord2);
            for (int i = 0; i < n; ++i) jgx[i] = (int32_t)ch[W.orda[ord2[i]]].x;
            for (int i = 0; i < n; ++i) jgy[i] = ch[W.orda[ord2[i]]].y;      // overwrites keys only after both reads of it are over
        }
        // windows_of + windows_to_ranges, with the retry that splits windows at contig borders
        int64_t *jlo = wlo + w_used, *jhi = whi + w_used;
        int n_win = 0;
        for (int attempt = 0; attempt < 2; ++attempt) {
            const bool split = attempt == 1;
            // the (first, second) pairs are produced one at a time; a pair is final once the next one starts
            n_win = 0;
            bool retry = false;
            int64_t first = ys[0], second = ys[0];
            int curc = rc.ctg.cid(ys[0]);
            bool open = true;
            auto emit = [&](int64_t a, int64_t b) -> bool {      // false: windows span two contigs -> retry
                int64_t min_ref = a, max_ref = b;
                const int cc = rc.ctg.cid(min_ref);
                if (cc != rc.ctg.cid(max_ref)) return false;
                const int64_t cs = rc.ctg.start[cc];
                const int64_t lookfurther = imin(look_span, min_ref - cs);
                min_ref -= lookfurther;
                max_ref += look_span;
                int64_t lo, hi;
                rc.ctg.slice(cc, min_ref - cs, max_ref - cs, lo, hi);
                jlo[n_win] = lo; jhi[n_win] = hi; ++n_win;
                return true;
            };
            for (int i = 1; i < n && !retry; ++i) {
                const int64_t y = ys[i];
                if ((y - second) < readgap && (!split || curc == rc.ctg.cid(y))) second = y;
                else {
                    if (first != second && !emit(first, second)) retry = true;
                    first = y; second = y;
                    curc = rc.ctg.cid(y);
                }
            }
            if (!retry && open && first != second && !emit(first, second)) retry = true;
            if (!retry || split) break;
            // the reference keeps the windows added before the cross-contig one and then ADDS those of the second attempt
            // (windows_to_ranges clears its lists first in the host version: job.win_lo.clear()) -- the host version clears
        }
        RJob J;
        J.read = rc.read; J.need_reverse = rc.need_reverse ? 1 : 0;
        J.readstart = (int32_t)imax(0, (int64_t)jgx[0] - look_span);
        J.readend = (int32_t)imin(rc.L - kmer + 1, (int64_t)jgx[n - 1] + look_span);
        J.n_win = n_win; J.n_guide = n;
        J.win_off = w_base + w_used; J.g_off = g_base + g_used;
        J.hit_off = 0; J.hit_cap = 0; J.pad0 = 0; J.dense_off = 0; J.tab_off = 0; J.tab_size = 0; J.pad = 0;
        jobs[ji] = J;
        g_used += n;
        w_used += n_win;
    }
    return g_used;
}

} // namespace vmd
