// Host-side per-read glue of the alignment pipeline (pure C++17, no CUDA).
//
// The reference keeps this logic in Python/numba between its hot loops
// (mammap_clrnano.py: hit2work_1 bookkeeping :23588-23734, guide-chain selection
// :28479-28574, rebuild_chain_break :23437, extend_edge_test :2302, drop_misplaced
// :726, merge_conjacent :16736, fix_simple_inv :24226, split_alignment_test :21505,
// get_onemapinfolist :20731, pairedindel :5604).  Here it is host C++ that prepares
// job lists for, and consumes results of, the CUDA kernels; it never computes a
// DP, a sketch, a k-mer join or an alignment itself.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <set>
#include <stdexcept>
#include <string>
#include <unordered_set>
#include <vector>

namespace vmg {

static const int kNoPre = -9999999;

struct Anc {
    int64_t x, y;   // read position, GLOBAL reference position
    int32_t s, l;   // strand +1/-1, length
};
typedef std::vector<Anc> Path;

// anchor as the kernels store it (16 bytes; == VmAnchor): what the chaining stages hand back
struct Anc32 {
    int32_t x;
    uint32_t y;
    int32_t s, l;
};
static inline Anc widen(const Anc32 &a) { return Anc{a.x, (int64_t)a.y, a.s, a.l}; }

struct ReadDropped : public std::runtime_error {
    explicit ReadDropped(const char *m) : std::runtime_error(m) {}
};

// ---- numba quicksort permutation (numba/misc/quicksort.py), host replay ----
template <typename K>
static void argsort_replay(const K *A, int64_t n, std::vector<int64_t> &R)
{
    R.resize((size_t)n);
    for (int64_t t = 0; t < n; ++t) R[t] = t;
    if (n < 2) return;
    {
        // strictly monotone keys have a single sorted order, whatever the algorithm: no replay needed
        bool asc = true, desc = true;
        for (int64_t t = 1; t < n && (asc || desc); ++t) {
            if (!(A[t - 1] < A[t])) asc = false;
            if (!(A[t] < A[t - 1])) desc = false;
        }
        if (asc) return;
        if (desc) { std::reverse(R.begin(), R.end()); return; }
    }
    int64_t slo[100], shi[100];
    int sp = 1;
    slo[0] = 0; shi[0] = n - 1;
    while (sp > 0) {
        --sp;
        int64_t low = slo[sp], high = shi[sp];
        while (high - low >= 15) {
            int64_t mid = (low + high) >> 1;
            if (A[R[mid]] < A[R[low]]) std::swap(R[low], R[mid]);
            if (A[R[high]] < A[R[mid]]) std::swap(R[high], R[mid]);
            if (A[R[mid]] < A[R[low]]) std::swap(R[low], R[mid]);
            const K pivot = A[R[mid]];
            std::swap(R[high], R[mid]);
            int64_t i = low, j = high - 1;
            for (;;) {
                while (i < high && A[R[i]] < pivot) ++i;
                while (j >= low && pivot < A[R[j]]) --j;
                if (i >= j) break;
                std::swap(R[i], R[j]);
                ++i; --j;
            }
            std::swap(R[i], R[high]);
            if (high - i > i - low) {
                if (high > i) { slo[sp] = i + 1; shi[sp] = high; ++sp; }
                high = i - 1;
            } else {
                if (i > low) { slo[sp] = low; shi[sp] = i - 1; ++sp; }
                low = i + 1;
            }
        }
        for (int64_t i = low + 1; i <= high; ++i) {
            const int64_t k = R[i];
            const K v = A[k];
            int64_t j = i;
            while (j > low && v < A[R[j - 1]]) { R[j] = R[j - 1]; --j; }
            R[j] = k;
        }
    }
}

// ---- reference sequence access ----
struct Contigs {
    std::vector<std::string> names;
    std::vector<int64_t> start;   // global offset of each contig
    std::vector<int64_t> len;
    const char *seq = nullptr;    // concatenated, upper-case, host copy
    int64_t total = 0;

    // pos2contig :51-59 -- last contig whose start <= pos (first contig if pos precedes all)
    int cid(int64_t pos) const
    {
        int c = 0;
        for (size_t i = 0; i < start.size(); ++i) {
            if (pos < start[i]) break;
            c = (int)i;
        }
        return c;
    }
    // Python slice contig[a:b] -> [lo, hi) in GLOBAL coordinates
    void slice(int c, int64_t a, int64_t b, int64_t &lo, int64_t &hi) const
    {
        const int64_t n = len[c];
        if (a < 0) a = std::max<int64_t>(a + n, 0);
        if (b < 0) b = std::max<int64_t>(b + n, 0);
        a = std::min(a, n);
        b = std::min(b, n);
        if (b < a) b = a;
        lo = start[c] + a;
        hi = start[c] + b;
    }
};

static inline void pyslice(int64_t n, int64_t a, int64_t b, int64_t &lo, int64_t &hi)
{
    if (a < 0) a = std::max<int64_t>(a + n, 0);
    if (b < 0) b = std::max<int64_t>(b + n, 0);
    a = std::min(a, n);
    b = std::min(b, n);
    if (b < a) b = a;
    lo = a; hi = b;
}

struct ModeConst {
    double accept;      // primary-chain acceptance threshold (:23650)
    int max_guides;     // <=0: unlimited (:28581)
    int local_maxgap;   // :24061
    bool clamp40;       // mode L: skipcost = min(skipcost, 40) for the multi-chain local DP
};

struct Options {
    double global_skipcost = 40, local_skipcost = 40, maxdivergence = 0.2;
    int global_maxdiff = 50, local_maxdiff = 30;
    int check_num = 100;
    bool eqx = false, hardclip = false, nodiscard = false;
    ModeConst mode{60.0, 5, 99, false};
};

// ---------------------------------------------------------------------------
// hit2work_1 bookkeeping after the DP (:23581-23734)
// ---------------------------------------------------------------------------
struct GlobalResult {
    bool ok = false;
    int mapq = 0;
    double score = 0;          // score of the primary chain
    std::vector<Path> guides;  // [0] primary chain, then secondary chains; each DESCENDING read order
};

static void hit2work_chains(std::vector<Path> &path_list, std::vector<double> &scores_list, const std::vector<double> &S_arr,
                            int64_t L, GlobalResult &out, int bin_size, double overlap);

static void hit2work(const Anc32 *a, const double *S, const int32_t *P, const int32_t *S_arg, int64_t n, int64_t g,
                     int64_t L, double accept, GlobalResult &out, int bin_size = 100, double overlap = 0.5)
{
    out = GlobalResult();
    std::vector<char> used((size_t)n, 0);
    std::vector<Path> path_list;
    std::vector<double> scores_list;
    std::vector<double> S_arr;
    bool hit = false;
    {
        Path path;
        int64_t take = g;
        used[take] = 1;
        const double score = S[take];
        for (;;) {
            path.push_back(widen(a[take]));
            S_arr.push_back(S[take]);
            if (P[take] == kNoPre) break;
            take = P[take];
            used[take] = 1;
        }
        if (score > 40) {
            hit = true;
            scores_list.push_back(score);
            path_list.push_back(std::move(path));
        }
    }
    const double scores = S[g];
    const double max_scores = scores > 0 ? scores : 0;
    if (!(hit && max_scores > accept)) return;   // nothing below can change the verdict
    Path path;
    for (int64_t q = n - 1; q >= 0; --q) {
        int64_t take = S_arg[q];
        if (used[take]) continue;
        path.clear();
        used[take] = 1;
        double score = S[take];
        for (;;) {
            path.push_back(widen(a[take]));
            if (P[take] == kNoPre) break;
            take = P[take];
            if (used[take]) { score = score - S[take]; break; }
            used[take] = 1;
        }
        if (score > 40) {
            scores_list.push_back(score);
            path_list.push_back(path);
        }
    }
    hit2work_chains(path_list, scores_list, S_arr, L, out, bin_size, overlap);
}

// The same bookkeeping from chains that were already extracted (on the device, vm_extract.cu): chain c is
// anc[sum(len[0..c)) .. + len[c]) in descending read order, chain 0 the primary one with its S values in S_prim.
static void hit2work_extracted(const Anc32 *anc, const double *S_prim, const int32_t *len, const double *score, int n_chains,
                               int64_t L, GlobalResult &out, int bin_size = 100, double overlap = 0.5)
{
    out = GlobalResult();
    if (n_chains <= 0) return;
    std::vector<Path> path_list((size_t)n_chains);
    std::vector<double> scores_list(score, score + n_chains);
    int64_t o = 0;
    for (int c = 0; c < n_chains; ++c) {
        Path &p = path_list[(size_t)c];
        p.resize((size_t)len[c]);
        for (int32_t t = 0; t < len[c]; ++t) p[(size_t)t] = widen(anc[o + t]);
        o += len[c];
    }
    std::vector<double> S_arr(S_prim, S_prim + len[0]);
    hit2work_chains(path_list, scores_list, S_arr, L, out, bin_size, overlap);
}

static void hit2work_chains(std::vector<Path> &path_list, std::vector<double> &scores_list, const std::vector<double> &S_arr,
                            int64_t L, GlobalResult &out, int bin_size, double overlap)
{
    (void)L;
    std::vector<int64_t> order;
    argsort_replay<double>(scores_list.data(), (int64_t)scores_list.size(), order);
    std::reverse(order.begin(), order.end());
    if (order[0] != 0) {
        for (size_t i = 0; i < order.size(); ++i)
            if (order[i] == 0) { order[i] = order[0]; order[0] = 0; break; }
    }
    // 100-bp read-bin sets as sorted unique vectors (the reference's Python sets are only ever
    // intersected and counted, :23672-23694)
    auto binset = [&](const Path &p, std::vector<int64_t> &s) {
        // chains come in descending read order: walking them backwards gives the bins already sorted
        s.clear();
        s.reserve(p.size());
        bool sorted = true;
        for (size_t i = p.size(); i-- > 0;) {
            const int64_t b = p[i].x / bin_size;
            if (!s.empty() && b < s.back()) sorted = false;
            if (s.empty() || b != s.back()) s.push_back(b);
        }
        if (!sorted) {
            std::sort(s.begin(), s.end());
            s.erase(std::unique(s.begin(), s.end()), s.end());
        }
    };
    std::vector<std::vector<int64_t>> prim_sets(1);
    std::vector<std::vector<double>> prim_scores;
    if (order.size() > 1) binset(path_list[order[0]], prim_sets[0]);      // only ever compared with other chains
    prim_scores.push_back({scores_list[order[0]]});
    std::vector<int64_t> b;
    for (size_t oi = 1; oi < order.size(); ++oi) {
        const int64_t iloc = order[oi];
        binset(path_list[iloc], b);
        double best = 0.0;
        size_t pref = 0;
        for (size_t p = 0; p < prim_sets.size(); ++p) {
            const std::vector<int64_t> &ps = prim_sets[p];
            size_t inter = 0, i = 0, j = 0;
            if (!b.empty() && !ps.empty() && b.front() <= ps.back() && ps.front() <= b.back())
                while (i < b.size() && j < ps.size()) {
                    if (b[i] < ps[j]) ++i;
                    else if (ps[j] < b[i]) ++j;
                    else { ++inter; ++i; ++j; }
                }
            const double ov = (double)inter / (double)std::min(ps.size(), b.size());
            if (ov > best) { best = ov; pref = p; }
        }
        if (best < overlap) {
            prim_sets.push_back(b);
            prim_scores.push_back({scores_list[iloc]});
        } else prim_scores[pref].push_back(scores_list[iloc]);
    }
    const double m = (double)path_list[order[0]].size();
    const double f1 = prim_scores[0][0];
    const double f2 = prim_scores[0].size() >= 2 ? prim_scores[0][1] : 0.0;
    // min(int(40*(1-f2/f1)*min(1, m/10)*np.log(f1)), 60)  (:23704); numba lowers np.log to libm log
    double v = 40 * (1 - f2 / f1);
    v = v * std::min(1.0, m / 10);
    v = v * std::log(f1);
    out.mapq = (int)std::min<int64_t>((int64_t)v, 60);
    // select_secondary_alignment :23505-23538
    std::vector<const Path *> secondary;
    if (path_list.size() > 1) {
        // loc2score[q] (:23508-23515) = S of the primary-chain anchor with the largest start <= q, 0 below the
        // chain; the chain is in descending read order, so a lookup is one binary search
        const Path &prim = path_list[0];
        const size_t np = std::min(prim.size(), S_arr.size());
        auto loc2score_at = [&](int64_t q) -> double {
            size_t lo = 0, hi = np;          // first t with prim[t].x <= q
            while (lo < hi) {
                const size_t mid = (lo + hi) >> 1;
                if (prim[mid].x <= q) hi = mid; else lo = mid + 1;
            }
            return lo < np ? S_arr[lo] : 0.0;
        };
        for (size_t oi = 1; oi < order.size(); ++oi) {
            const Path &one = path_list[order[oi]];
            const double f2s = scores_list[order[oi]];
            const int64_t en_loc = one.front().x, st_loc = one.back().x;
            if (en_loc - st_loc < 50) continue;
            const double f1s = std::max(loc2score_at(en_loc) - loc2score_at(st_loc), 1.0);
            if (f2s / f1s > 0.9 || std::fabs(f1s - f2s) < 40) {
                bool skip = false;
                for (const Path *pri : secondary) {
                    const int64_t pe = pri->front().x, ps = pri->back().x;
                    const int64_t ovs = std::max<int64_t>(std::min(en_loc, pe) - std::max(ps, st_loc), 0);
                    if ((double)ovs / (double)(en_loc - st_loc) > 0.5) { skip = true; break; }
                }
                if (!skip) secondary.push_back(&one);
            }
        }
    }
    out.ok = true;
    out.score = scores_list[0];
    out.guides.reserve(1 + secondary.size());
    {   // path_list is not used after this: the chains move into the result
        std::vector<Path> sec;
        sec.reserve(secondary.size());
        for (const Path *p : secondary) sec.push_back(std::move(*const_cast<Path *>(p)));
        out.guides.push_back(std::move(path_list[0]));
        for (Path &p : sec) out.guides.push_back(std::move(p));
    }
}

// ---------------------------------------------------------------------------
// guide-chain selection for the local stage (:28529-28582)
// ---------------------------------------------------------------------------
static void merge_chain(std::vector<Path> &chains)
{
    if (chains.size() <= 1) return;
    std::vector<Path> rest(std::make_move_iterator(chains.begin() + 1), std::make_move_iterator(chains.end()));
    if (!rest.empty()) {
        std::vector<int64_t> keys, order;
        for (const Path &c : rest) keys.push_back(c.back().x);
        argsort_replay<int64_t>(keys.data(), (int64_t)keys.size(), order);
        std::vector<Path> t;
        for (int64_t i : order) t.push_back(rest[i]);
        rest.swap(t);
    }
    size_t iloc = 0;
    while (iloc + 1 < rest.size()) {
        size_t jloc = iloc + 1;
        while (jloc < rest.size()) {
            const Anc &a0 = rest[iloc].front();
            const Anc &bl = rest[jloc].back();
            if (a0.x + a0.l <= bl.x && a0.s == bl.s) {
                const int64_t readgap = bl.x - a0.x - a0.l;
                const int64_t refgap = a0.s == 1 ? bl.y - a0.y - a0.l : a0.y - bl.y - bl.l;
                if (std::llabs(readgap - refgap) < 500) {
                    Path merged = rest[jloc];
                    merged.insert(merged.end(), rest[iloc].begin(), rest[iloc].end());
                    rest[iloc].swap(merged);
                    rest.erase(rest.begin() + jloc);
                    continue;
                }
            }
            ++jloc;
        }
        ++iloc;
    }
    if (!rest.empty()) {
        std::vector<int64_t> keys, order;
        for (const Path &c : rest) keys.push_back((int64_t)c.size());
        argsort_replay<int64_t>(keys.data(), (int64_t)keys.size(), order);
        std::vector<Path> t;
        for (int64_t i : order) t.push_back(rest[i]);
        rest.swap(t);
    }
    chains.resize(1);
    for (Path &p : rest) chains.push_back(std::move(p));
}

static void drop_somechains(std::vector<Path> &chains)
{
    const size_t m = chains.size() - 1;
    std::vector<size_t> iloclist(m, 0);
    std::vector<int64_t> distance(m, INT64_MAX);
    std::vector<int64_t> sc0(m, 0), sc1(m, 0), c0(m, 0), c1(m, 0);
    for (const Anc &item : chains[0]) {
        for (size_t ci = 0; ci < m; ++ci) {
            const Path &chain = chains[ci + 1];
            if (item.x >= chain.back().x && item.x <= chain.front().x) {
                if (item.s == 1) sc0[ci]++; else sc1[ci]++;
            }
            while (chain[iloclist[ci]].x > item.x) {
                if (iloclist[ci] < chain.size() - 1) iloclist[ci]++;
                else break;
            }
            const int64_t d = std::llabs(item.y - chain[iloclist[ci]].y);
            if (d < distance[ci]) distance[ci] = d;
        }
    }
    for (size_t ci = 0; ci < m; ++ci)
        for (const Anc &item : chains[ci + 1]) {
            if (item.s == 1) c0[ci]++; else c1[ci]++;
        }
    std::vector<Path> out;
    out.reserve(chains.size());
    out.push_back(std::move(chains[0]));
    for (size_t ci = 0; ci < m; ++ci) {
        const bool keep = (sc0[ci] > sc1[ci] && c0[ci] > c1[ci]) || (sc0[ci] < sc1[ci] && c0[ci] < c1[ci]);
        Path &ch = chains[ci + 1];
        if ((!keep && distance[ci] < 500) || (ch.front().x - ch.back().x) < 100) continue;
        out.push_back(std::move(ch));
    }
    chains.swap(out);
}

// chains in the order the reference re-seeds them; returns how many are re-seeded
static size_t select_guides(std::vector<Path> &chains, const ModeConst &mc)
{
    merge_chain(chains);
    drop_somechains(chains);
    std::vector<double> keys;
    std::vector<int64_t> order;
    for (const Path &c : chains) keys.push_back(1.0 / (double)c.size());
    argsort_replay<double>(keys.data(), (int64_t)keys.size(), order);
    std::vector<Path> t;
    t.reserve(chains.size());
    for (int64_t i : order) t.push_back(std::move(chains[i]));
    chains.swap(t);
    size_t used = 1;
    int count = 2;
    for (size_t i = 1; i < chains.size(); ++i) {
        ++used;
        ++count;
        if (mc.max_guides > 0 && count > mc.max_guides) break;
    }
    return used;
}

// ---------------------------------------------------------------------------
// one guide chain -> inputs of the re-seeding kernel (:23090-23191)
// ---------------------------------------------------------------------------
struct GuideJob {
    std::vector<int64_t> win_lo, win_hi;   // reference windows, GLOBAL [lo, hi), in insertion order
    std::vector<int32_t> gx;               // guide read positions, ascending (after the :23183 argsort)
    std::vector<int64_t> gy;               // matching guide reference positions
    int32_t readstart = 0, readend = 0;
};

static void windows_of(const int64_t *ys, size_t n, int64_t readgap, const Contigs &ctg, bool split_contigs,
                       std::vector<std::pair<int64_t, int64_t>> &se)
{
    se.clear();
    se.push_back({ys[0], ys[0]});
    int cur = ctg.cid(ys[0]);
    for (size_t i = 1; i < n; ++i) {
        const int64_t y = ys[i];
        if ((y - se.back().second) < readgap && (!split_contigs || cur == ctg.cid(y))) se.back().second = y;
        else {
            if (se.back().first == se.back().second) se.pop_back();
            se.push_back({y, y});
            cur = ctg.cid(y);
        }
    }
    if (!se.empty() && se.back().first == se.back().second) se.pop_back();
}

static bool windows_to_ranges(const std::vector<std::pair<int64_t, int64_t>> &se, const Contigs &ctg, int64_t look_span,
                              GuideJob &job)
{
    job.win_lo.clear();
    job.win_hi.clear();
    for (const auto &w : se) {
        int64_t min_ref = w.first, max_ref = w.second;
        const int c = ctg.cid(min_ref);
        if (c != ctg.cid(max_ref)) return true;   // retry_diffcontig; windows added so far stay (as in the reference)
        const int64_t cs = ctg.start[c];
        const int64_t lookfurther = std::min(look_span, min_ref - cs);
        min_ref -= lookfurther;
        max_ref += look_span;
        int64_t lo, hi;
        ctg.slice(c, min_ref - cs, max_ref - cs, lo, hi);
        job.win_lo.push_back(lo);
        job.win_hi.push_back(hi);
    }
    return false;
}

static void make_guide_job(const Path &chain, int64_t L, int k, const Contigs &ctg, GuideJob &job)
{
    const int64_t look_span = 7000;
    int64_t readgap = 0;
    for (size_t i = 1; i < chain.size(); ++i) {
        const int64_t d = std::llabs(chain[i].x - chain[i - 1].x);
        if (d > readgap) readgap = d;
    }
    readgap = std::max<int64_t>(readgap + 1000, 5000);
    const size_t n = chain.size();
    // Common case: read positions strictly descending (every chain out of the DP) and reference positions strictly
    // monotone.  Then both argsorts (:23103 by reference position, :23183 by read position) have a single possible
    // result and the guide points are the chain in ascending read order -- no permutation to replay.
    bool x_desc = true, y_asc = true, y_desc = true;
    for (size_t i = 1; i < n; ++i) {
        if (!(chain[i].x < chain[i - 1].x)) x_desc = false;
        if (!(chain[i - 1].y < chain[i].y)) y_asc = false;
        if (!(chain[i].y < chain[i - 1].y)) y_desc = false;
    }
    std::vector<int64_t> ys(n);
    std::vector<std::pair<int64_t, int64_t>> se;
    job.gx.resize(n);
    job.gy.resize(n);
    if (n >= 1 && x_desc && (y_asc || y_desc)) {
        for (size_t i = 0; i < n; ++i) {
            ys[i] = y_asc ? chain[i].y : chain[n - 1 - i].y;
            job.gx[i] = (int32_t)chain[n - 1 - i].x;
            job.gy[i] = chain[n - 1 - i].y;
        }
    } else {
        std::vector<int64_t> keys(n), order;
        for (size_t i = 0; i < n; ++i) keys[i] = chain[i].y;
        argsort_replay<int64_t>(keys.data(), (int64_t)n, order);
        std::vector<Anc> by_y(n);
        for (size_t i = 0; i < n; ++i) { by_y[i] = chain[order[i]]; ys[i] = by_y[i].y; }
        for (size_t i = 0; i < n; ++i) keys[i] = by_y[i].x;
        argsort_replay<int64_t>(keys.data(), (int64_t)n, order);
        for (size_t i = 0; i < n; ++i) {
            job.gx[i] = (int32_t)by_y[order[i]].x;
            job.gy[i] = by_y[order[i]].y;
        }
    }
    windows_of(ys.data(), n, readgap, ctg, false, se);
    if (windows_to_ranges(se, ctg, look_span, job)) {
        windows_of(ys.data(), n, readgap, ctg, true, se);
        windows_to_ranges(se, ctg, look_span, job);
    }
    job.readstart = (int32_t)std::max<int64_t>(0, (int64_t)job.gx.front() - look_span);
    job.readend = (int32_t)std::min<int64_t>(L - k + 1, (int64_t)job.gx.back() + look_span);
}

// ---------------------------------------------------------------------------
// local traceback with overlap trimming (:27508-27527); result ASCENDING read order
// ---------------------------------------------------------------------------
static void local_traceback(const Anc32 *a, const int32_t *P, int64_t g, Path &asc)
{
    Path path;
    int64_t take = g;
    path.push_back(widen(a[take]));
    Anc pre = widen(a[take]);
    for (;;) {
        if (P[take] == kNoPre) break;
        take = P[take];
        const Anc now = widen(a[take]);
        if (pre.x < now.x + now.l) {
            const int64_t ov = now.x + now.l - pre.x;
            Anc t;
            t.x = pre.x + ov;
            t.y = pre.s == 1 ? pre.y + ov : pre.y;
            t.s = pre.s;
            t.l = (int32_t)(pre.l - ov);
            path.back() = t;
        }
        path.push_back(now);
        pre = now;
    }
    asc.assign(path.rbegin(), path.rend());
}

// ---------------------------------------------------------------------------
// sub-alignment surgery
// ---------------------------------------------------------------------------
typedef std::vector<Path> AlnList;

static void rebuild_chain_break(const Contigs &ctg, const Path &raw, int64_t large_cost, AlnList &al,
                                int64_t small_alignment = 50)
{
    al.clear();
    Anc pre = raw[0];
    al.push_back(Path{pre});
    al.back().reserve(raw.size());
    for (size_t i = 1; i < raw.size(); ++i) {
        const Anc &now = raw[i];
        if (pre.s == now.s) {
            const int64_t readgap = now.x - pre.x - pre.l;
            const int64_t refgap = pre.s == 1 ? now.y - pre.y - pre.l : pre.y - now.y - now.l;
            if (std::llabs(readgap - refgap) <= large_cost && refgap >= -20 && readgap < 100) {
                if (ctg.cid(pre.y) == ctg.cid(now.y)) {
                    if (refgap >= 0) { al.back().push_back(now); pre = now; continue; }
                    if (readgap <= 20) continue;
                    al.back().push_back(now);
                    pre = now;
                    continue;
                }
            }
        }
        if (al.back().size() == 1) al.pop_back();
        if (!al.empty()) {
            const Path &b = al.back();
            if ((b.back().x + b.back().l - b.front().x) < small_alignment) al.pop_back();
        }
        al.push_back(Path{now});
        al.back().reserve(raw.size() - i);
        pre = now;
    }
    if (al.back().size() == 1) al.pop_back();
    if (al.empty()) throw ReadDropped("rebuild_chain_break: empty");
    {
        const Path &b = al.back();
        if ((b.back().x + b.back().l - b.front().x) < small_alignment) al.pop_back();
    }
}

// A pair of sequences handed to an alignment kernel.  Each side is a slice of either the
// concatenated reference or of the read / its reverse complement, optionally reversed and/or
// complemented on the fly by the kernel.
struct SeqRef {
    int32_t src = 0;      // 0 reference (global coords), 1 read forward, 2 read reverse-complement
    int64_t lo = 0, hi = 0;
    int32_t reverse = 0;  // read the slice back to front
    int32_t comp = 0;     // complement each base
    int64_t len() const { return hi - lo; }
};

// get_query_target_for_cigar :5802-5818
static void query_target(const Anc &pre, const Anc &now, int64_t L, const Contigs &ctg, SeqRef &target, SeqRef &query)
{
    target = SeqRef();
    query = SeqRef();
    if (pre.s == 1) {
        const int c = ctg.cid(pre.y);
        const int64_t b = ctg.start[c];
        ctg.slice(c, pre.y - b, now.y - b, target.lo, target.hi);
        query.src = 1;
        pyslice(L, pre.x, now.x, query.lo, query.hi);
    } else {
        const int c = ctg.cid(now.y);
        const int64_t b = ctg.start[c];
        ctg.slice(c, now.y + now.l - b, pre.y + pre.l - b, target.lo, target.hi);
        query.src = 2;
        pyslice(L, L - now.x, L - pre.x, query.lo, query.hi);
    }
}

// Exact-match segments of one sub-alignment in the coordinates of its divergence-filter job
// (query_target above): (query offset, target offset, length), ascending, clipped to the two slices and
// trimmed so that consecutive segments overlap in neither sequence.  They let the device bound the edit
// distance from above by an alignment that runs through the chain's anchors (vm_ed_upper_kernel).
struct MatchSeg { int32_t q, t, l; };
static inline bool match_segments(const Path &aln, int64_t qlen, int64_t tlen, std::vector<MatchSeg> &out, int64_t max_qgap)
{
    const size_t first = out.size();
    if (aln.size() < 2) return false;
    const Anc &pre = aln.front(), &now = aln.back();
    const bool fwd = pre.s == 1;
    int64_t cq = 0, ct = 0;
    const size_t n = aln.size();
    for (size_t k = 0; k < n; ++k) {
        const Anc &a = fwd ? aln[k] : aln[n - 1 - k];     // ascending in the job's query coordinate
        if (a.s != pre.s) { out.resize(first); return false; }
        int64_t q = fwd ? a.x - pre.x : now.x - a.x - a.l;
        int64_t t = fwd ? a.y - pre.y : a.y - (now.y + now.l);
        int64_t l = a.l;
        int64_t d = std::max<int64_t>(std::max<int64_t>(cq - q, ct - t), 0);
        q += d; t += d; l -= d;
        l = std::min(l, std::min(qlen - q, tlen - t));
        if (l <= 0) continue;
        if (q - cq > max_qgap) { out.resize(first); return false; }
        out.push_back(MatchSeg{(int32_t)q, (int32_t)t, (int32_t)l});
        cq = q + l;
        ct = t + l;
    }
    if (qlen - cq > max_qgap) { out.resize(first); return false; }
    return true;
}

struct ExtJob {      // one z-drop edge extension (k_cigar 2,-4,4,4,4,4 bw 100 zdrop 50)
    int32_t aln = 0;
    int32_t side = 0;    // 0: left of the first anchor, 1: right of the last anchor
    int32_t strand = 1;
    int64_t q_anchor = 0, t_anchor = 0;   // query_st/target_st (left) or query_en/target_en|target_st (right)
    SeqRef target, query;
    int32_t q_e = 0, t_e = 0;             // result
};

// extend_edge_test :2302-2525, split in the two dependency rounds described in DESIGN.md:
// round 0 = every right-hand extension plus the left-hand extension of sub-alignment 0
// (none of them depends on another extension of this call); round 1 = the remaining
// left-hand extensions, which start from the (already extended) end of their predecessor.
// `prepare` emits the jobs of a round (and applies the job-free boundary rewrites),
// `apply` writes the kernel results back.
static void extend_prepare(int round, int64_t L, AlnList &al, const Contigs &ctg, std::vector<ExtJob> &jobs)
{
    const int64_t max_extend = 20000;
    jobs.clear();
    for (size_t idx = 0; idx < al.size(); ++idx) {
        Path &one = al[idx];
        const bool left_now = (round == 0) == (idx == 0);
        if (left_now) {
            if (one[0].x > 0) {
                int64_t looksize;
                if (idx == 0) looksize = one[0].x;
                else looksize = one[0].x - (al[idx - 1].back().x + al[idx - 1].back().l);
                const Anc pre = one[0];
                const int c = ctg.cid(pre.y);
                const int64_t cs = ctg.start[c], clen = ctg.len[c];
                ExtJob j;
                j.aln = (int32_t)idx; j.side = 0; j.strand = pre.s;
                if (pre.s == 1) {
                    const int64_t target_st = pre.y, query_st = pre.x;
                    looksize = std::min(looksize, target_st - cs);
                    if (looksize > max_extend) looksize = max_extend;
                    if (looksize != 0) {
                        j.q_anchor = query_st; j.t_anchor = target_st;
                        j.query.src = 1; j.query.reverse = 1;
                        pyslice(L, std::max<int64_t>(query_st - looksize, 0), query_st, j.query.lo, j.query.hi);
                        j.target.src = 0; j.target.reverse = 1;
                        ctg.slice(c, target_st - cs - j.query.len(), target_st - cs, j.target.lo, j.target.hi);
                        jobs.push_back(j);
                    }
                } else {
                    const int64_t target_en = pre.y + pre.l, query_st = pre.x;
                    looksize = std::min(looksize, cs + clen - (target_en - 1));
                    if (looksize > max_extend) looksize = max_extend;
                    if (looksize != 0) {
                        j.q_anchor = query_st; j.t_anchor = target_en;
                        j.query.src = 1; j.query.reverse = 1;
                        pyslice(L, std::max<int64_t>(query_st - looksize, 0), query_st, j.query.lo, j.query.hi);
                        // revcomp(ref[target_en : target_en+len])[::-1] == complement, forward order
                        j.target.src = 0; j.target.reverse = 0; j.target.comp = 1;
                        ctg.slice(c, target_en - cs, target_en + j.query.len() - cs, j.target.lo, j.target.hi);
                        jobs.push_back(j);
                    }
                }
            } else {
                const Anc t = one[0];
                if (t.s == 1) one[0] = Anc{t.x, t.y, 1, 0};
                else one[0] = Anc{t.x, t.y + t.l, -1, 0};
            }
        }
        if (round == 0) {
            if ((one.back().x + one.back().l) < L) {
                int64_t looksize;
                if (idx + 1 == al.size()) looksize = L - (one.back().x + one.back().l);
                else looksize = al[idx + 1][0].x - (one.back().x + one.back().l);
                const Anc pre = one[one.size() - 2], now = one.back();
                const int c = ctg.cid(pre.y);
                const int64_t cs = ctg.start[c], clen = ctg.len[c];
                ExtJob j;
                j.aln = (int32_t)idx; j.side = 1; j.strand = pre.s;
                if (pre.s == 1) {
                    const int64_t target_en = now.y + now.l, query_en = now.x + now.l;
                    looksize = std::min(looksize, cs + clen - (target_en - 1));
                    if (looksize > max_extend) looksize = max_extend;
                    if (looksize != 0) {
                        j.q_anchor = query_en; j.t_anchor = target_en;
                        j.query.src = 1;
                        pyslice(L, query_en, query_en + looksize, j.query.lo, j.query.hi);
                        j.target.src = 0;
                        ctg.slice(c, target_en - cs, target_en + j.query.len() - cs, j.target.lo, j.target.hi);
                        jobs.push_back(j);
                    }
                } else {
                    const int64_t target_st = now.y, query_en = now.x + now.l;
                    looksize = std::min(looksize, target_st - cs);
                    if (looksize > max_extend) looksize = max_extend;
                    if (looksize != 0) {
                        j.q_anchor = query_en; j.t_anchor = target_st;
                        j.query.src = 1;
                        pyslice(L, query_en, query_en + looksize, j.query.lo, j.query.hi);
                        // revcomp(ref[target_st-len : target_st]) : reversed and complemented
                        j.target.src = 0; j.target.reverse = 1; j.target.comp = 1;
                        ctg.slice(c, target_st - cs - j.query.len(), target_st - cs, j.target.lo, j.target.hi);
                        jobs.push_back(j);
                    }
                }
            } else {
                const Anc t = one.back();
                if (t.s == 1) one.back() = Anc{t.x + t.l, t.y + t.l, 1, 0};
                else one.back() = Anc{t.x + t.l, t.y, -1, 0};
            }
        }
    }
}

static void extend_apply(AlnList &al, const std::vector<ExtJob> &jobs)
{
    for (const ExtJob &j : jobs) {
        Path &one = al[j.aln];
        if (j.side == 0) {
            if (j.strand == 1) one[0] = Anc{j.q_anchor - j.q_e, j.t_anchor - j.t_e, 1, 0};
            else one[0] = Anc{j.q_anchor - j.q_e, j.t_anchor + j.t_e, -1, 0};
        } else {
            if (j.strand == 1) one.back() = Anc{j.q_anchor + j.q_e, j.t_anchor + j.t_e, 1, 0};
            else one.back() = Anc{j.q_anchor + j.q_e, j.t_anchor - j.t_e, -1, 0};
        }
    }
}

static inline void gaps_of(const Anc &pre, const Anc &now, int64_t &readgap, int64_t &refgap)
{
    readgap = now.x - pre.x - pre.l;
    refgap = pre.s == 1 ? now.y - pre.y - pre.l : pre.y - now.y - now.l;
}

// drop_misplaced_alignment_test :726-786
static bool drop_misplaced(AlnList &al, size_t iloc)
{
    const Path &a = al[iloc], &b = al[iloc + 1], &c = al[iloc + 2];
    if (a[0].s == b[0].s && a[0].s == c[0].s) {
        const int64_t mid = b.back().x + b.back().l - b[0].x;
        if (mid > 1000) return false;
        int64_t readgap, refgap;
        gaps_of(a.back(), b[0], readgap, refgap);
        if (std::llabs(refgap) < 100000) {
            int DEL = 0, INS = 0;
            if (readgap - refgap < -30) ++DEL;
            else if (readgap - refgap > 30) ++INS;
            else return false;
            const int64_t gap_1 = std::llabs(readgap - refgap);
            gaps_of(b.back(), c[0], readgap, refgap);
            if (std::llabs(refgap) < 100000) {
                if (readgap - refgap < -30) ++DEL;
                else if (readgap - refgap > 30) ++INS;
                else return false;
                const int64_t gap_2 = std::llabs(readgap - refgap);
                if (DEL == 1 && INS == 1 && (mid < 500 || (double)std::max(gap_1, gap_2) / (double)mid > 0.5)) {
                    al.erase(al.begin() + iloc + 1);
                    return true;
                }
            }
        }
    }
    return false;
}

// getdupiloc_numba :16680-16734 (keeps the `[0][2]` strand-for-length quirk)
static void getdupiloc(const AlnList &al, std::vector<size_t> &dup)
{
    dup.clear();
    if (al.size() < 2) return;
    size_t iloc = 0;
    while (iloc + 1 < al.size()) {
        const Anc &la = al[iloc].back();
        const int64_t readpos_1 = la.x + la.l;
        const int64_t refpos_1 = la.s == 1 ? la.y + la.l : la.y;
        const int strand_1 = la.s == 1 ? 1 : -1;
        size_t jloc = iloc, new_iloc = 0;
        bool hit = false;
        int64_t dupsize = 0, readpos_2 = 0;
        while (jloc + 1 < al.size()) {
            ++jloc;
            int64_t refpos_2;
            int strand_2;
            if (al[jloc].back().s == 1) { refpos_2 = al[jloc][0].y; strand_2 = 1; }
            else { refpos_2 = al[jloc][0].y + al[jloc][0].s; strand_2 = -1; }
            if (strand_1 != strand_2) continue;
            const int64_t d = strand_1 == 1 ? refpos_2 - refpos_1 : refpos_1 - refpos_2;
            if (d < 50) { new_iloc = jloc; dupsize = d; readpos_2 = al[jloc][0].x; hit = true; }
        }
        if (hit) {
            const int64_t readgap = readpos_2 - readpos_1;
            if ((iloc + 1) < new_iloc || ((dupsize - readgap) < -30 && readgap < 30))
                for (size_t s = iloc; s < new_iloc; ++s) dup.push_back(s);
            iloc = new_iloc;
        } else ++iloc;
    }
}

// merge_conjacent_alignment :16736-16780
static void merge_conjacent(AlnList &al, const Contigs &ctg)
{
    if (al.size() < 2) return;
    std::vector<size_t> dup;
    getdupiloc(al, dup);
    size_t iloc = 0;
    while (iloc + 1 < al.size()) {
        if (std::find(dup.begin(), dup.end(), iloc) != dup.end()) { ++iloc; continue; }
        const Anc pre = al[iloc].back(), now = al[iloc + 1][0];
        if (pre.s != now.s || ctg.cid(pre.y) != ctg.cid(now.y)) { ++iloc; continue; }
        int64_t readgap, refgap;
        gaps_of(pre, now, readgap, refgap);
        if (refgap < 0) { ++iloc; continue; }
        if (std::min(readgap, refgap) < 50 && std::llabs(readgap - refgap) < 10000) {
            al[iloc].insert(al[iloc].end(), al[iloc + 1].begin(), al[iloc + 1].end());
            al.erase(al.begin() + iloc + 1);
        } else ++iloc;
    }
}

// fix_simple_inv :24226-24312.  `read` = oriented read (upper-case), ctg.seq = reference.
// Returns whether a breakpoint was re-cut.
static bool fix_simple_inv(AlnList &al, const Contigs &ctg, const char *read, int64_t L)
{
    if (al.size() <= 2) return false;
    bool changed = false;
    size_t iloc = 0;
    while (iloc + 2 < al.size()) {
        Path &A = al[iloc], &B = al[iloc + 1], &C = al[iloc + 2];
        if (A[0].s == C[0].s && A[0].s != B[0].s && A[0].s == 1) {
            const int c = ctg.cid(A[0].y);
            const int64_t bias0 = ctg.start[c];
            const int64_t refen_0 = A.back().y + A.back().l - bias0;
            const int64_t readen_0 = A.back().x + A.back().l;
            const int64_t refst_1 = B.back().y - bias0;
            const int64_t readst_1 = B[0].x;
            const int64_t refen_1 = B[0].y + B[0].l - bias0;
            const int64_t readen_1 = B.back().x + B.back().l;
            const int64_t refst_2 = C[0].y - bias0;
            const int64_t readst_2 = C[0].x;
            if (refst_2 - refen_0 == refen_1 - refst_1 && readst_1 - readen_0 + readst_2 - readen_1 == 0) {
                if (refst_1 - refen_0 != 0 && refst_1 - refen_0 + refst_2 - refen_1 == 0) {
                    if (refen_0 > refst_1) {
                        int64_t rlo, rhi, qlo, qhi;
                        ctg.slice(c, refen_1, refen_1 + refen_0 - refst_1, rlo, rhi);
                        pyslice(L, readen_0 - refen_0 + refst_1, readen_0, qlo, qhi);
                        bool same = (rhi - rlo) == (qhi - qlo);
                        for (int64_t t = 0; same && t < rhi - rlo; ++t) {
                            const char rc = ctg.seq[rhi - 1 - t];
                            const char cc = rc == 'A' ? 'T' : rc == 'T' ? 'A' : rc == 'G' ? 'C' : rc == 'C' ? 'G' : 'N';
                            if (cc != read[qlo + t]) same = false;
                        }
                        if (same) {
                            changed = true;
                            const int64_t bias = refen_0 - refst_1;
                            C[0] = Anc{readst_2 - bias, refst_2 - bias + bias0, 1, 0};
                            const Anc ins{readst_2 - bias, refen_0 + bias0, -1, 0};
                            for (;;) {
                                if (B.empty()) throw ReadDropped("fix_simple_inv emptied a sub-alignment");
                                if (ins.x <= B.back().x + B.back().l) B.pop_back();
                                else break;
                            }
                            B.push_back(ins);
                        }
                    } else {
                        int64_t rlo, rhi, qlo, qhi;
                        ctg.slice(c, refen_0, refst_1, rlo, rhi);
                        pyslice(L, readen_0, readen_0 - refen_0 + refst_1, qlo, qhi);
                        bool same = (rhi - rlo) == (qhi - qlo);
                        for (int64_t t = 0; same && t < rhi - rlo; ++t)
                            if (ctg.seq[rlo + t] != read[qlo + t]) same = false;
                        if (same) {
                            changed = true;
                            A.back() = Anc{readen_0 - refen_0 + refst_1, refst_1 + bias0, 1, 0};
                            const Anc ins{readen_0 - refen_0 + refst_1, refen_1 + refen_0 - refst_1 + bias0, -1, 0};
                            for (;;) {
                                if (B.empty()) throw ReadDropped("fix_simple_inv emptied a sub-alignment");
                                if (ins.x >= B[0].x) B.erase(B.begin());
                                else break;
                            }
                            B.insert(B.begin(), ins);
                        }
                    }
                }
            }
        }
        ++iloc;
    }
    return changed;
}

struct FillJob {     // one global dual-affine fill (k_cigar 2,-4,4,2,24,1 bw -1 zdrop -1)
    int32_t aln = 0;
    SeqRef target, query;
};

// split_alignment_test :21505-21617: which anchor pairs get a fill, in CIGAR order.
// `kept` receives the new (possibly reversed) anchor list of the sub-alignment.
static void split_alignment(Path &alignment, int aln_index, int64_t L, const Contigs &ctg, Path &kept,
                            std::vector<FillJob> &jobs)
{
    kept.clear();
    const size_t before = jobs.size();
    if (alignment[0].s == 1) {
        if (alignment.back().l != 0) {
            const Anc t = alignment.back();
            alignment.back() = Anc{t.x + t.l, t.y + t.l, 1, 0};
        }
        Anc pre = alignment[0];
        kept.push_back(pre);
        size_t iloc = 1;
        while (iloc < alignment.size()) {
            const Anc now = alignment[iloc];
            const int64_t readgap = now.x - pre.x - pre.l;
            const int64_t refgap = now.y - pre.y - pre.l;
            if (now.l < 19 || std::min(readgap, refgap) < 200) {
                if (iloc + 1 != alignment.size()) { ++iloc; continue; }
            }
            FillJob j;
            j.aln = aln_index;
            query_target(pre, now, L, ctg, j.target, j.query);
            if (j.target.len() > 0 && j.query.len() > 0) {
                jobs.push_back(j);
                kept.push_back(now);
            } else throw ReadDropped("Failed to compute CIGAR");
            pre = now;
            ++iloc;
        }
    } else {
        if (alignment[0].l != 0) {
            const Anc t = alignment[0];
            alignment[0] = Anc{t.x, t.y + t.l, -1, 0};
        }
        if (alignment.back().l != 0) {
            const Anc t = alignment.back();
            alignment.back() = Anc{t.x + t.l, t.y, -1, 0};
        }
        Path rev(alignment.rbegin(), alignment.rend());
        Anc pre = rev[0];
        kept.push_back(pre);
        size_t iloc = 1;
        while (iloc < rev.size()) {
            const Anc now = rev[iloc];
            const int64_t readgap = pre.x - now.x - now.l;
            const int64_t refgap = now.y - pre.y - pre.l;
            if (now.l < 19 || std::min(readgap, refgap) < 200) {
                if (iloc + 1 != rev.size()) { ++iloc; continue; }
            }
            FillJob j;
            j.aln = aln_index;
            query_target(now, pre, L, ctg, j.target, j.query);
            if (j.target.len() > 0 && j.query.len() > 0) {
                jobs.push_back(j);
                kept.push_back(now);
            } else throw ReadDropped("Failed to compute CIGAR");
            pre = now;
            ++iloc;
        }
    }
    if (jobs.size() == before) throw ReadDropped("cigarlist[-1] == []");
}

// ---------------------------------------------------------------------------
// record assembly (get_onemapinfolist :20731-20838); CIGAR kept as BAM-encoded ops
// ---------------------------------------------------------------------------
struct Record {
    int32_t contig = 0;
    int32_t strand = 1;         // +1 / -1 as emitted ('+' / '-')
    int64_t q_st = 0, q_en = 0, r_st = 0, r_en = 0;
    int32_t mapq = 0;
    std::vector<uint32_t> cigar;   // len<<4|op, op: 0 M, 1 I, 2 D, 4 S, 5 H, 7 =, 8 X; NOT run-merged (as the reference's string)
};

static int64_t cigar_query_len(const std::vector<uint32_t> &c)
{
    int64_t n = 0;
    for (uint32_t o : c) {
        const uint32_t op = o & 0xf;
        if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) n += o >> 4;
    }
    return n;
}

// a run of CIGAR ops inside a kernel's output array
struct OpSpan { const uint32_t *p; int32_t n; };

// kept[i]: anchors of sub-alignment i after split_alignment; cig[i]: the fill segments' ops of it, in order
static void make_records(const AlnList &kept, const std::vector<std::vector<OpSpan>> &cig, int mapq, int64_t L,
                         const Contigs &ctg, bool need_reverse, bool hardclip, std::vector<Record> &out)
{
    out.clear();
    const uint32_t clip = hardclip ? 5u : 4u;
    for (size_t i = 0; i < kept.size(); ++i) {
        const Path &al = kept[i];
        Record r;
        r.contig = ctg.cid(al[0].y);
        const int64_t bias = ctg.start[r.contig];
        r.mapq = mapq;
        int64_t tailM = 0;
        if (al[0].s == 1) {
            r.q_st = al[0].x;
            r.q_en = al.back().x + al.back().l;
            r.r_st = al[0].y - bias;
            r.r_en = al.back().y + al.back().l - bias;
            if (al.back().l > 0) tailM = al.back().l;
            r.strand = need_reverse ? -1 : 1;
        } else {
            r.q_st = L - al[0].x - al[0].l;
            r.q_en = L - al.back().x;
            r.r_st = al[0].y - bias;
            r.r_en = al.back().y + al.back().l - bias;
            r.strand = need_reverse ? 1 : -1;
        }
        size_t n_ops = 3;
        for (const OpSpan &sp : cig[i]) n_ops += (size_t)sp.n;
        r.cigar.reserve(n_ops);
        if (r.q_st > 0) r.cigar.push_back((uint32_t)r.q_st << 4 | clip);
        for (const OpSpan &sp : cig[i]) r.cigar.insert(r.cigar.end(), sp.p, sp.p + sp.n);
        if (tailM > 0) r.cigar.push_back((uint32_t)tailM << 4 | 0u);
        if (L - r.q_en > 0) r.cigar.push_back((uint32_t)(L - r.q_en) << 4 | clip);
        const int64_t want = hardclip ? (r.q_en - r.q_st) : L;
        if (want != cigar_query_len(r.cigar)) throw ReadDropped("cigar length check");
        out.push_back(std::move(r));
    }
    if (need_reverse) std::reverse(out.begin(), out.end());
}

// pairedindel :5604-5650 on BAM-encoded ops
static bool paired_indel(const std::vector<Record> &recs, double indelsize = 30)
{
    std::vector<double> indel;
    for (const Record &r : recs)
        for (uint32_t o : r.cigar) {
            const uint32_t op = o & 0xf;
            if ((op == 1 || op == 2) && (double)(o >> 4) > indelsize) indel.push_back((double)(o >> 4));
        }
    std::sort(indel.begin(), indel.end());
    double pre = 0;
    for (double now : indel) {
        const double mx = std::max(pre, now);
        if (mx > 0 && (std::min(pre, now) / mx) > 0.7) return true;
        pre = now;
    }
    return false;
}

} // namespace vmg
