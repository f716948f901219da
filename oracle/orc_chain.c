/*
 * orc_chain.c -- ORACLE (test infrastructure, not product code).
 *
 * Plain-C, single-threaded, literal restatement of the reference's anchor
 * sorting and non-linear chaining DPs.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may call this.
 *
 * Pinned against the reference's own numba functions imported in the build
 * container (tests/golden/make_golden.py -> tests/golden/chain_*.npz).
 *
 * Each function cites the reference lines it follows
 * (all in /root/reference/src/vacmap/mammap_clrnano.py unless noted).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define NOPRE (-9999999)

/* ------------------------------------------------------------------ */
/* numba np.argsort / list.sort(key=) : numba/misc/quicksort.py        */
/* (median-of-3, pivot stashed at high, Hoare scan, insertion < 15)    */
/* ------------------------------------------------------------------ */
#define QS_SMALL 15
#define QS_BODY(T)                                                            \
    int64_t stack_lo[100], stack_hi[100];                                     \
    int64_t sp = 0;                                                           \
    for (int64_t t = 0; t < n; ++t) R[t] = t;                                 \
    if (n < 2) return;                                                        \
    stack_lo[0] = 0; stack_hi[0] = n - 1; sp = 1;                             \
    while (sp > 0) {                                                          \
        --sp;                                                                 \
        int64_t low = stack_lo[sp], high = stack_hi[sp];                      \
        while (high - low >= QS_SMALL) {                                      \
            int64_t mid = (low + high) >> 1, tmp;                             \
            if (A[R[mid]] < A[R[low]]) { tmp = R[low]; R[low] = R[mid]; R[mid] = tmp; } \
            if (A[R[high]] < A[R[mid]]) { tmp = R[high]; R[high] = R[mid]; R[mid] = tmp; } \
            if (A[R[mid]] < A[R[low]]) { tmp = R[low]; R[low] = R[mid]; R[mid] = tmp; } \
            T pivot = A[R[mid]];                                              \
            tmp = R[high]; R[high] = R[mid]; R[mid] = tmp;                    \
            int64_t i = low, j = high - 1;                                    \
            for (;;) {                                                        \
                while (i < high && A[R[i]] < pivot) ++i;                      \
                while (j >= low && pivot < A[R[j]]) --j;                      \
                if (i >= j) break;                                            \
                tmp = R[i]; R[i] = R[j]; R[j] = tmp;                          \
                ++i; --j;                                                     \
            }                                                                 \
            tmp = R[i]; R[i] = R[high]; R[high] = tmp;                        \
            if (high - i > i - low) {                                         \
                if (high > i) { stack_lo[sp] = i + 1; stack_hi[sp] = high; ++sp; } \
                high = i - 1;                                                 \
            } else {                                                          \
                if (i > low) { stack_lo[sp] = low; stack_hi[sp] = i - 1; ++sp; } \
                low = i + 1;                                                  \
            }                                                                 \
        }                                                                     \
        if (high > low) {                                                     \
            for (int64_t i = low + 1; i <= high; ++i) {                       \
                int64_t k = R[i]; T v = A[k]; int64_t j = i;                  \
                while (j > low && v < A[R[j - 1]]) { R[j] = R[j - 1]; --j; }  \
                R[j] = k;                                                     \
            }                                                                 \
        }                                                                     \
    }

void orc_argsort_i64(const int64_t *A, int64_t n, int64_t *R) { QS_BODY(int64_t) }
void orc_argsort_f64(const double *A, int64_t n, int64_t *R) { QS_BODY(double) }

/* ------------------------------------------------------------------ */
/* Score tables                                                        */
/* ------------------------------------------------------------------ */
/* gapcost_list inside the njit DPs: global `24843-24846`, local `27317-27322`.
 * libm log2 is what numba's np.log2 lowers to. */
void orc_gapcost_table(int kmersize, int maxdiff, int local_variant, double *out)
{
    out[0] = 0.0;
    for (int g = 1; g <= maxdiff; ++g) {
        double lg = log2((double)g);
        if (!local_variant || g <= 10) out[g] = 0.01 * kmersize * g + 0.5 * lg;
        else out[g] = 0.01 * kmersize * g + 2 * lg;
    }
}

/* large_readgapcost_list of the multi-chain local DP, `28270-28275` (float32). */
void orc_large_readgap_table(int maxgap, int large_readgap, float *out)
{
    out[0] = 0.0f;
    for (int r = 1; r <= maxgap; ++r) {
        if (large_readgap <= r) out[r] = (float)(0.5 * r);
        else out[r] = (float)(0.1 * log2((double)(r + 1)));
    }
}

typedef struct {
    const float *extra;      /* module table `15371-15376`, float32 */
    int64_t extra_size;      /* len(extra) - 1 */
    const float *readgapcost;/* module table `26567-26569`, float32[100] */
    const double *log2cache; /* module table `27530`, float64[100000] */
    int64_t log2cache_size;  /* len - 1 */
} orc_tables;

/* pairwise gap geometry shared by every DP variant (`24953-24984`, `27418-27456`) */
static inline void pair_gaps(const int64_t *ai, const int64_t *aj,
                             int64_t *bonus, int64_t *readgap, int64_t *refgap)
{
    int64_t rg = ai[0] - aj[0] - aj[3];
    if (rg < 0) {
        int64_t b = ai[0] + ai[3] - aj[0] - aj[3];
        int64_t ov = aj[0] + aj[3] - ai[0];
        *bonus = b; *readgap = 0;
        if (ai[2] == aj[2]) {
            if (ai[2] == 1) *refgap = ai[1] + ov - (aj[1] + aj[3]);
            else *refgap = aj[1] - (ai[1] + b);
        } else {
            if (aj[2] == -1) *refgap = ai[1] + ov - aj[1] + 1;
            else *refgap = ai[1] + b - 1 - (aj[1] + aj[3]);
        }
    } else {
        *bonus = ai[3]; *readgap = rg;
        if (ai[2] == aj[2]) {
            if (ai[2] == 1) *refgap = ai[1] - aj[1] - aj[3];
            else *refgap = aj[1] - ai[1] - ai[3];
        } else {
            if (aj[2] == -1) *refgap = ai[1] - aj[1] + 1;
            else *refgap = ai[1] + ai[3] - 1 - aj[1] - aj[3];
        }
    }
}

/* insertpoint_score `19369-19387` */
static int64_t insertpoint_score(const double *S, double target, int64_t k, const int32_t *arg)
{
    int64_t i = 0, j = k;
    if (S[arg[0]] > target) return 0;
    if (S[arg[k - 1]] < target) return k;
    while (i < j) {
        int64_t mid = (i + j) / 2;
        double now = S[arg[mid]];
        if (now < target) i = mid + 1;
        else if (now > target) j = mid;
        else return mid + 1;
    }
    return j;
}

/*
 * Global DP, exact variant: get_optimal_chain_..._fine_list_d_all `24828-25031`.
 * a: int64[n][4] sorted by read position.  Returns g_max_index or -1 on the
 * opcount bail-out (`24914`).  opcount_out receives the evaluation counter.
 */
int64_t orc_chain_global_d_all(const int64_t *a, int64_t n, int kmersize,
                               double skipcost_in, int64_t maxdiff_in, int64_t maxgap,
                               const orc_tables *tb, int64_t max_factor,
                               double *S, int32_t *P, int32_t *S_arg, int64_t *opcount_out)
{
    const int64_t repeat_weight = 20;
    double *gapcost_list = (double *)malloc(sizeof(double) * (size_t)(maxdiff_in + 1));
    orc_gapcost_table(kmersize, (int)maxdiff_in, 0, gapcost_list);
    int64_t lastpos = a[(n - 1) * 4];
    int64_t *cov = (int64_t *)calloc((size_t)(lastpos + 1), sizeof(int64_t));
    for (int64_t i = 0; i < n; ++i) {
        int64_t x = a[i * 4];
        cov[x] = cov[x] + 1 < repeat_weight ? cov[x] + 1 : repeat_weight;
    }
    int64_t prereadloc = a[0];
    double skipcost = skipcost_in + (double)cov[a[0]];
    int64_t maxdiff = maxdiff_in - cov[a[0]] > 10 ? maxdiff_in - cov[a[0]] : 10;
    int64_t testspace_en = 1;
    S_arg[0] = 0;
    S[0] = (double)a[3];
    P[0] = NOPRE;
    double g_max_scores = (double)a[3];
    int64_t g_max_index = 0;
    int64_t opcount = 0;
    int64_t ret = 0;

    for (int64_t i = 1; i < n; ++i) {
        const int64_t *ai = a + i * 4;
        double max_scores = (double)ai[3];
        int64_t pre_index = NOPRE;
        if (prereadloc < ai[0]) {
            if (((double)opcount / (double)i) > (double)max_factor) { ret = -1; goto done; }
            for (int64_t k = testspace_en; k < i; ++k) {
                int64_t loc = insertpoint_score(S, S[k], k, S_arg);
                memmove(S_arg + loc + 1, S_arg + loc, sizeof(int32_t) * (size_t)(k - loc));
                S_arg[loc] = (int32_t)k;
            }
            testspace_en = i;
            skipcost = skipcost_in + (double)cov[ai[0]];
            maxdiff = maxdiff_in - cov[ai[0]] > 10 ? maxdiff_in - cov[ai[0]] : 10;
            prereadloc = ai[0];
        }
        for (int64_t q = testspace_en - 1; q >= 0; --q) {
            int64_t j = S_arg[q];
            if (S[j] > (max_scores - (double)ai[3])) {
                ++opcount;
                int64_t bonus, readgap, refgap;
                pair_gaps(ai, a + j * 4, &bonus, &readgap, &refgap);
                int64_t gapcost = llabs(readgap - refgap);
                double t;
                if (ai[2] == a[j * 4 + 2] && refgap >= 0 && readgap <= maxgap && gapcost <= maxdiff) {
                    t = S[j] + (double)bonus - gapcost_list[gapcost];
                } else {
                    if (gapcost > tb->extra_size) gapcost = tb->extra_size;
                    t = S[j] - skipcost + (double)bonus - (double)tb->extra[gapcost];
                }
                if (t > max_scores) { max_scores = t; pre_index = j; }
            } else break;
        }
        S[i] = max_scores;
        P[i] = (int32_t)pre_index;
        if (max_scores > g_max_scores) { g_max_scores = max_scores; g_max_index = i; }
    }
    for (int64_t k = testspace_en; k < n; ++k) {
        int64_t loc = insertpoint_score(S, S[k], k, S_arg);
        memmove(S_arg + loc + 1, S_arg + loc, sizeof(int32_t) * (size_t)(k - loc));
        S_arg[loc] = (int32_t)k;
    }
    ret = g_max_index;
done:
    if (opcount_out) *opcount_out = opcount;
    free(cov);
    free(gapcost_list);
    return ret;
}

/*
 * asm mode (mammap_asm.py), SURVEY 8f-1: pairwise gap geometry of the linked DPs (`21795-21824`).  Same shape as
 * pair_gaps above, but the older formulas: no +-1 on opposite strands, overlap handled through the non-overlapping
 * length of anchor i.
 */
static inline void pair_gaps_asm(const int64_t *ai, const int64_t *aj,
                                 int64_t *bonus, int64_t *readgap, int64_t *refgap)
{
    int64_t rg = ai[0] - aj[0] - aj[3];
    if (rg < 0) {
        int64_t nos = ai[0] - aj[0];
        *bonus = ai[0] + ai[3] - aj[0] - aj[3];
        *readgap = 0;
        if (ai[2] == aj[2]) {
            if (ai[2] == 1) *refgap = ai[1] - aj[1] - nos;
            else *refgap = aj[1] + aj[3] - nos - ai[1] - ai[3];
        } else {
            if (aj[2] == -1) *refgap = ai[1] + aj[3] - nos - aj[1];
            else *refgap = ai[1] + ai[3] - aj[1] - nos;
        }
    } else {
        *bonus = ai[3];
        *readgap = rg;
        if (ai[2] == aj[2]) {
            if (ai[2] == 1) *refgap = ai[1] - aj[1] - aj[3];
            else *refgap = aj[1] - ai[1] - ai[3];
        } else {
            if (aj[2] == -1) *refgap = ai[1] - aj[1];
            else *refgap = ai[1] + ai[3] - aj[1] - aj[3];
        }
    }
}

/*
 * asm mode, global DP with carry-in: linked_get_optimal_chain_..._fine_list_d_all (mammap_asm.py `21687-21871`).
 * a: int64[n][4] = the anchors carried over from the previous batch (pre_n of them, in ascending score order)
 * followed by this batch's anchors sorted by read position.  pre_n > 0: S / P of the first pre_n anchors are the
 * carried (rebased) scores and negated back-pointers, g_max_scores / g_max_index / prereadloc come from the caller
 * and only S_arg[0] = 0 is in the test space until the read position first advances (`21713-21718`).  pre_n == 0:
 * the plain start (`21720-21729`).  No coverage term; the scan stops at the first S_j <= best - l_i (`21777`).
 * Returns g_max_index, or -1 on the opcount bail-out (`21754`).
 * rgcost != NULL: the second-round twin linked_..._fine_list_all (`21505-21686`) instead -- the same loop, anchors
 * still ordered by read START, colinear pairs also pay readgapcost_list[readgap] (asm's table, `16536-16538`:
 * 0.1 log2(r), float32[100]) and there is no bail-out.
 */
int64_t orc_chain_linked_d_all(const int64_t *a, int64_t n, int64_t pre_n, const double *pre_S, const int32_t *pre_P,
                               double g_max_scores, int64_t g_max_index, int64_t prereadloc, int kmersize,
                               double skipcost, int64_t maxdiff, int64_t maxgap, const orc_tables *tb, int64_t max_factor,
                               const float *rgcost, double *S, int32_t *P, int32_t *S_arg, int64_t *opcount_out)
{
    double *gapcost_list = (double *)malloc(sizeof(double) * (size_t)(maxdiff + 1));
    orc_gapcost_table(kmersize, (int)maxdiff, 0, gapcost_list);
    int64_t testspace_en = 1, pre_size;
    S_arg[0] = 0;
    if (pre_n > 0) {
        memcpy(S, pre_S, sizeof(double) * (size_t)pre_n);
        memcpy(P, pre_P, sizeof(int32_t) * (size_t)pre_n);
        pre_size = pre_n;
    } else {
        S[0] = (double)a[3];
        P[0] = NOPRE;
        g_max_scores = (double)a[3];
        g_max_index = 0;
        prereadloc = a[0];
        pre_size = 1;
    }
    int64_t opcount = 0;
    int64_t ret = 0;
    for (int64_t i = pre_size; i < n; ++i) {
        const int64_t *ai = a + i * 4;
        double max_scores = (double)ai[3];
        int64_t pre_index = NOPRE;
        if (prereadloc < ai[0]) {
            if (!rgcost && ((double)opcount / (double)i) > (double)max_factor) { ret = -1; goto done; }
            for (int64_t k = testspace_en; k < i; ++k) {
                int64_t loc = insertpoint_score(S, S[k], k, S_arg);
                memmove(S_arg + loc + 1, S_arg + loc, sizeof(int32_t) * (size_t)(k - loc));
                S_arg[loc] = (int32_t)k;
            }
            testspace_en = i;
            prereadloc = ai[0];
        }
        for (int64_t q = testspace_en - 1; q >= 0; --q) {
            int64_t j = S_arg[q];
            if (S[j] > (max_scores - (double)ai[3])) {
                ++opcount;
                int64_t bonus, readgap, refgap;
                pair_gaps_asm(ai, a + j * 4, &bonus, &readgap, &refgap);
                int64_t gapcost = llabs(readgap - refgap);
                double t;
                if (ai[2] == a[j * 4 + 2] && refgap >= 0 && readgap <= maxgap && gapcost <= maxdiff) {
                    t = S[j] + (double)bonus - gapcost_list[gapcost];
                    if (rgcost) t = t - (double)rgcost[readgap];
                } else {
                    if (gapcost > tb->extra_size) gapcost = tb->extra_size;
                    t = S[j] - skipcost + (double)bonus - (double)tb->extra[gapcost];
                }
                if (t > max_scores) { max_scores = t; pre_index = j; }
            } else break;
        }
        S[i] = max_scores;
        P[i] = (int32_t)pre_index;
        if (max_scores > g_max_scores) { g_max_scores = max_scores; g_max_index = i; }
    }
    for (int64_t k = testspace_en; k < n; ++k) {
        int64_t loc = insertpoint_score(S, S[k], k, S_arg);
        memmove(S_arg + loc + 1, S_arg + loc, sizeof(int32_t) * (size_t)(k - loc));
        S_arg[loc] = (int32_t)k;
    }
    ret = g_max_index;
done:
    if (opcount_out) *opcount_out = opcount;
    free(gapcost_list);
    return ret;
}

/* insertpoint_score_distance `17200-17226` */
static int64_t insertpoint_score_distance(const int64_t *Si, int64_t target, int64_t k,
                                          const int32_t *arg, int64_t tdist, const int64_t *dist)
{
    int64_t i = 0, j = k;
    if (Si[arg[0]] > target) return 0;
    if (Si[arg[k - 1]] < target) return k;
    while (i < j) {
        int64_t mid = (i + j) / 2;
        int64_t now = Si[arg[mid]];
        if (now < target) i = mid + 1;
        else if (now > target) j = mid;
        else {
            int64_t nd = dist[arg[mid]];
            if (nd < tdist) i = mid + 1;
            else if (nd > tdist) j = mid;
            else return mid + 1;
        }
    }
    return j;
}

/* closest2targetdistance `17228-17251` */
static int64_t closest2targetdistance(int64_t tdist, const int64_t *dist, const int32_t *arg,
                                      int64_t st, int64_t en)
{
    int64_t i = st, j = en;
    if (dist[arg[i]] >= tdist) return i;
    if (dist[arg[j - 1]] <= tdist) return j - 1;
    while (i < j) {
        int64_t mid = (i + j) / 2;
        int64_t nd = dist[arg[mid]];
        if (nd < tdist) i = mid + 1;
        else if (nd > tdist) j = mid;
        else return mid;
    }
    if ((tdist - dist[arg[j - 1]]) < (dist[arg[j]] - tdist)) return j - 1;
    return j;
}

/*
 * Scoring variants.
 *   variant 0: global (`24987-24996`)
 *   variant 1: local single-chain fine_list (`27459-27480`)
 *   variant 2: local multi-chain fine_list_mismatch (`28416-28428`)
 * Returns 1 when the pair produced a candidate score in *t, 0 when skipped
 * (local `bonus <= 0` continue, `27423`).
 */
typedef struct {
    int variant;
    double skipcost;
    int64_t maxdiff, maxgap;
    const double *gapcost_list;
    const float *rgcost;  /* variant 1: readgapcost_list; variant 2: large_readgapcost_list */
    const orc_tables *tb;
} score_ctx;

static inline int pair_score(const score_ctx *c, const int64_t *ai, const int64_t *aj,
                             double Sj, double *t)
{
    int64_t bonus, readgap, refgap;
    pair_gaps(ai, aj, &bonus, &readgap, &refgap);
    if (c->variant != 0 && (ai[0] - aj[0] - aj[3]) < 0 && bonus <= 0) return 0;
    int64_t gapcost = llabs(readgap - refgap);
    if (ai[2] == aj[2] && refgap >= 0 && readgap <= c->maxgap && gapcost <= c->maxdiff) {
        if (c->variant == 0) *t = Sj + (double)bonus - c->gapcost_list[gapcost];
        else *t = Sj + (double)bonus - c->gapcost_list[gapcost] - (double)c->rgcost[readgap];
    } else {
        if (c->variant == 0) {
            if (gapcost > c->tb->extra_size) gapcost = c->tb->extra_size;
            *t = Sj - c->skipcost + (double)bonus - (double)c->tb->extra[gapcost];
        } else if (c->variant == 1) {
            if (gapcost > c->tb->extra_size) gapcost = c->tb->extra_size;
            double pen;
            if (ai[2] != aj[2]) pen = (50.0 < c->skipcost ? 50.0 : c->skipcost) + (double)c->tb->extra[gapcost];
            else pen = c->skipcost + (double)c->tb->extra[gapcost];
            *t = Sj + (double)bonus - pen;
        } else {
            int64_t g = gapcost < c->tb->log2cache_size ? gapcost : c->tb->log2cache_size;
            double pen = c->skipcost + c->tb->log2cache[g];
            *t = Sj + (double)bonus - pen;
        }
    }
    return 1;
}

/*
 * Heuristic DP shared by the global `_d_fast_all` (`25033-25339`, by_end = 0,
 * variant 0, coverage-adjusted skipcost/maxdiff) and the local `_fast`
 * fall-backs (`26938-27303`, `27891-28248`; by_end = 1, variants 1/2).
 * S_arg_i (int32) is the integer-score/diagonal ordered index list.
 */
int64_t orc_chain_fast(const int64_t *a, int64_t n, int kmersize, int variant,
                       double skipcost_in, int64_t maxdiff_in, int64_t maxgap, int64_t fast_t,
                       const orc_tables *tb, const float *rgcost,
                       double *S, int32_t *P, int32_t *S_arg_i)
{
    const int64_t repeat_weight = 20;
    const int by_end = variant != 0;
    double *gapcost_list = (double *)malloc(sizeof(double) * (size_t)(maxdiff_in + 1));
    orc_gapcost_table(kmersize, (int)maxdiff_in, variant != 0, gapcost_list);
    int64_t lastpos = a[(n - 1) * 4];
    int64_t *cov = (int64_t *)calloc((size_t)(lastpos + 5000), sizeof(int64_t));
    int64_t *target = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    int64_t *Si = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    int64_t readlength = lastpos + 1000;
    for (int64_t i = 0; i < n; ++i) {
        const int64_t *ai = a + i * 4;
        cov[ai[0]] = cov[ai[0]] + 1 < repeat_weight ? cov[ai[0]] + 1 : repeat_weight;
        if (ai[2] == 1) target[i] = ai[1] - ai[0] + readlength;
        else target[i] = -(ai[1] + ai[0] + readlength);
    }
    /* the reference sizes S_i_count as lastpos+50 and indexes it by integer score;
     * give it head-room so the restatement never reads out of bounds */
    int64_t cnt_size = lastpos + 50;
    {
        int64_t tot = 64;
        for (int64_t i = 0; i < n; ++i) tot += a[i * 4 + 3];
        if (tot > cnt_size) cnt_size = tot;
    }
    int64_t *Sicount = (int64_t *)calloc((size_t)cnt_size, sizeof(int64_t));

    score_ctx c;
    c.variant = variant; c.skipcost = skipcost_in; c.maxdiff = maxdiff_in; c.maxgap = maxgap;
    c.gapcost_list = gapcost_list; c.rgcost = rgcost; c.tb = tb;

    int64_t prereadloc = by_end ? a[0] + a[3] : a[0];
    int64_t testspace_en_i = 1;
    S_arg_i[0] = 0;
    S[0] = (double)a[3]; Si[0] = a[3]; P[0] = NOPRE;
    double g_max_scores = (double)a[3];
    int64_t g_max_index = 0;
    Sicount[a[3]] = 1;
    int64_t max_score_i = 0;

    for (int64_t i = 1; i < n; ++i) {
        const int64_t *ai = a + i * 4;
        P[i] = NOPRE;
        double max_scores = (double)ai[3];
        int64_t pre_index = NOPRE;
        int64_t key = by_end ? ai[0] + ai[3] : ai[0];
        if (prereadloc < key) {
            for (int64_t k = testspace_en_i; k < i; ++k) {
                Sicount[Si[k]] += 1;
                if (Si[k] > max_score_i) max_score_i = Si[k];
                int64_t loc = insertpoint_score_distance(Si, Si[k], k, S_arg_i, target[k], target);
                memmove(S_arg_i + loc + 1, S_arg_i + loc, sizeof(int32_t) * (size_t)(k - loc));
                S_arg_i[loc] = (int32_t)k;
            }
            testspace_en_i = i;
            if (variant == 0) {
                c.skipcost = skipcost_in + (double)cov[ai[0]];
                c.maxdiff = maxdiff_in - cov[ai[0]] > 10 ? maxdiff_in - cov[ai[0]] : 10;
            }
            prereadloc = key;
        }
        int64_t c_score_i = max_score_i;
        int64_t st_loc = testspace_en_i, en_loc = testspace_en_i;
        int64_t f_kmersize = ai[3] + 1;
        while ((double)c_score_i > (max_scores - (double)f_kmersize)) {
            int64_t now_count = Sicount[c_score_i];
            if (now_count == 0) { --c_score_i; continue; }
            st_loc = en_loc - now_count;
            if (now_count > fast_t) {
                int64_t j = S_arg_i[closest2targetdistance(target[i], target, S_arg_i, st_loc, en_loc)];
                double t;
                if (pair_score(&c, ai, a + j * 4, S[j], &t)) {
                    if (t > max_scores) { max_scores = t; pre_index = j; }
                }
            } else {
                for (int64_t q = en_loc - 1; q >= st_loc; --q) {
                    int64_t j = S_arg_i[q];
                    double t;
                    if (!pair_score(&c, ai, a + j * 4, S[j], &t)) continue;
                    if (t > max_scores) { max_scores = t; pre_index = j; }
                }
            }
            en_loc = st_loc;
            --c_score_i;
        }
        S[i] = max_scores;
        Si[i] = (int64_t)max_scores;
        P[i] = (int32_t)pre_index;
        if (max_scores > g_max_scores) { g_max_scores = max_scores; g_max_index = i; }
    }
    for (int64_t k = testspace_en_i; k < n; ++k) {
        Sicount[Si[k]] += 1;
        if (Si[k] > max_score_i) max_score_i = Si[k];
        int64_t loc = insertpoint_score_distance(Si, Si[k], k, S_arg_i, target[k], target);
        memmove(S_arg_i + loc + 1, S_arg_i + loc, sizeof(int32_t) * (size_t)(k - loc));
        S_arg_i[loc] = (int32_t)k;
    }
    free(gapcost_list); free(cov); free(target); free(Si); free(Sicount);
    return g_max_index;
}

/*
 * asm mode, heuristic twin of the linked DP: linked_..._fine_list_d_fast_all (mammap_asm.py `21872-22158`), taken
 * when the exact one bails out (`23246-23247`).  As orc_chain_fast(variant 0) above, with the asm gap geometry, no
 * coverage term, and the carried prefix: S_i[:pre_n] = int64(pre_S), only S_arg_i[0] = 0 / S_i_count[S_i[0]] = 1
 * in the test space and max_score_i = S_i[0] (`21906-21914`; the plain start keeps the reference's max_score_i = 0).
 */
int64_t orc_chain_linked_fast(const int64_t *a, int64_t n, int64_t pre_n, const double *pre_S, const int32_t *pre_P,
                              double g_max_scores, int64_t g_max_index, int64_t prereadloc, int kmersize,
                              double skipcost, int64_t maxdiff, int64_t maxgap, int64_t fast_t, const orc_tables *tb,
                              double *S, int32_t *P, int32_t *S_arg_i)
{
    double *gapcost_list = (double *)malloc(sizeof(double) * (size_t)(maxdiff + 1));
    orc_gapcost_table(kmersize, (int)maxdiff, 0, gapcost_list);
    int64_t lastpos = a[(n - 1) * 4];
    int64_t *target = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    int64_t *Si = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
    int64_t readlength = lastpos + 1000;
    for (int64_t i = 0; i < n; ++i) {
        const int64_t *ai = a + i * 4;
        if (ai[2] == 1) target[i] = ai[1] - ai[0] + readlength;
        else target[i] = -(ai[1] + ai[0] + readlength);
    }
    /* S_i_count is lastpos + 50 long in the reference and indexed by integer score: head-room as in orc_chain_fast */
    int64_t cnt_size = lastpos + 50;
    {
        int64_t tot = 64;
        for (int64_t i = 0; i < n; ++i) tot += a[i * 4 + 3];
        int64_t carried = 0;
        for (int64_t i = 0; i < pre_n; ++i)
            if ((int64_t)pre_S[i] > carried) carried = (int64_t)pre_S[i];
        tot += carried;
        if (tot > cnt_size) cnt_size = tot;
    }
    int64_t *Sicount = (int64_t *)calloc((size_t)cnt_size, sizeof(int64_t));
    int64_t testspace_en_i = 1, pre_size, max_score_i;
    S_arg_i[0] = 0;
    if (pre_n > 0) {
        for (int64_t i = 0; i < pre_n; ++i) { S[i] = pre_S[i]; Si[i] = (int64_t)pre_S[i]; P[i] = pre_P[i]; }
        pre_size = pre_n;
        Sicount[Si[0]] = 1;
        max_score_i = Si[0];
    } else {
        S[0] = (double)a[3]; Si[0] = a[3]; P[0] = NOPRE;
        pre_size = 1;
        g_max_scores = (double)a[3];
        g_max_index = 0;
        prereadloc = a[0];
        Sicount[a[3]] = 1;
        max_score_i = 0;
    }
    for (int64_t i = pre_size; i < n; ++i) {
        const int64_t *ai = a + i * 4;
        P[i] = NOPRE;
        double max_scores = (double)ai[3];
        int64_t pre_index = NOPRE;
        if (prereadloc < ai[0]) {
            for (int64_t k = testspace_en_i; k < i; ++k) {
                Sicount[Si[k]] += 1;
                if (Si[k] > max_score_i) max_score_i = Si[k];
                int64_t loc = insertpoint_score_distance(Si, Si[k], k, S_arg_i, target[k], target);
                memmove(S_arg_i + loc + 1, S_arg_i + loc, sizeof(int32_t) * (size_t)(k - loc));
                S_arg_i[loc] = (int32_t)k;
            }
            testspace_en_i = i;
            prereadloc = ai[0];
        }
        int64_t c_score_i = max_score_i;
        int64_t st_loc = testspace_en_i, en_loc = testspace_en_i;
        int64_t f_kmersize = ai[3] + 1;
        while ((double)c_score_i > (max_scores - (double)f_kmersize)) {
            int64_t now_count = Sicount[c_score_i];
            if (now_count == 0) { --c_score_i; continue; }
            st_loc = en_loc - now_count;
            int64_t q_hi = en_loc - 1, q_lo = st_loc;
            if (now_count > fast_t) q_hi = q_lo = closest2targetdistance(target[i], target, S_arg_i, st_loc, en_loc);
            for (int64_t q = q_hi; q >= q_lo; --q) {
                int64_t j = S_arg_i[q];
                const int64_t *aj = a + j * 4;
                int64_t bonus, readgap, refgap;
                pair_gaps_asm(ai, aj, &bonus, &readgap, &refgap);
                int64_t gapcost = llabs(readgap - refgap);
                double t;
                if (ai[2] == aj[2] && refgap >= 0 && readgap <= maxgap && gapcost <= maxdiff) {
                    t = S[j] + (double)bonus - gapcost_list[gapcost];
                } else {
                    if (gapcost > tb->extra_size) gapcost = tb->extra_size;
                    t = S[j] - skipcost + (double)bonus - (double)tb->extra[gapcost];
                }
                if (t > max_scores) { max_scores = t; pre_index = j; }
            }
            en_loc = st_loc;
            --c_score_i;
        }
        S[i] = max_scores;
        Si[i] = (int64_t)max_scores;
        P[i] = (int32_t)pre_index;
        if (max_scores > g_max_scores) { g_max_scores = max_scores; g_max_index = i; }
    }
    for (int64_t k = testspace_en_i; k < n; ++k) {
        Sicount[Si[k]] += 1;
        if (Si[k] > max_score_i) max_score_i = Si[k];
        int64_t loc = insertpoint_score_distance(Si, Si[k], k, S_arg_i, target[k], target);
        memmove(S_arg_i + loc + 1, S_arg_i + loc, sizeof(int32_t) * (size_t)(k - loc));
        S_arg_i[loc] = (int32_t)k;
    }
    free(gapcost_list); free(target); free(Si); free(Sicount);
    return g_max_index;
}

/* smallorequal2target_1d_point `13229-13264` (last index with S <= target, -1 if none) */
static int64_t smallorequal(const double *arr, double target, int64_t n, const int64_t *point)
{
    if (target < arr[point[0]]) return -1;
    if (target >= arr[point[n - 1]]) return n - 1;
    int64_t i = 0, j = n, mid = 0;
    while (i < j) {
        mid = (i + j) / 2;
        if (target == arr[point[mid]]) {
            if (mid < n - 1) {
                if (arr[point[mid + 1]] > target) return mid;
                else i = mid + 1;
            } else return mid;
        } else if (target < arr[point[mid]]) {
            if (mid > 0 && target >= arr[point[mid - 1]]) return mid - 1;
            j = mid;
        } else {
            if (mid < n - 1 && target < arr[point[mid + 1]]) return mid;
            i = mid + 1;
        }
    }
    return mid;
}

/*
 * Local DP, exact variants: `_fine_list` `27305-27528` (variant 1) and
 * `_fine_list_mismatch` `28250-28476` (variant 2).  Anchors sorted by read END.
 * Returns g_max_index, or -2 when the reference would switch to the `_fast`
 * variant (`27380-27384`); the caller then runs orc_chain_fast.
 */
int64_t orc_chain_local(const int64_t *a, int64_t n, int kmersize, int variant,
                        double skipcost, int64_t maxdiff, int64_t maxgap,
                        const orc_tables *tb, const float *rgcost,
                        double *S, int64_t *P, int64_t *S_arg, int64_t *opcount_out)
{
    double *gapcost_list = (double *)malloc(sizeof(double) * (size_t)(maxdiff + 1));
    orc_gapcost_table(kmersize, (int)maxdiff, 1, gapcost_list);
    score_ctx c;
    c.variant = variant; c.skipcost = skipcost; c.maxdiff = maxdiff; c.maxgap = maxgap;
    c.gapcost_list = gapcost_list; c.rgcost = rgcost; c.tb = tb;

    int64_t opcount = 0;
    int64_t prereadloc = a[0] + a[3];
    int64_t testspace_en = 1;
    S_arg[0] = 0;
    S[0] = (double)a[3]; P[0] = NOPRE;
    double g_max_scores = (double)a[3];
    int64_t g_max_index = 0;
    int64_t ret;

    for (int64_t i = 1; i < n; ++i) {
        const int64_t *ai = a + i * 4;
        double max_scores = (double)ai[3];
        int64_t pre_index = NOPRE;
        if (prereadloc < ai[0] + ai[3]) {
            if (opcount > 100000 && ((double)opcount / (double)prereadloc) > 1000.0) { ret = -2; goto done; }
            for (int64_t k = testspace_en; k < i; ++k) {
                int64_t loc = smallorequal(S, S[k], k, S_arg) + 1;
                memmove(S_arg + loc + 1, S_arg + loc, sizeof(int64_t) * (size_t)(k - loc));
                S_arg[loc] = k;
            }
            testspace_en = i;
            prereadloc = ai[0] + ai[3];
        }
        for (int64_t q = testspace_en - 1; q >= 0; --q) {
            int64_t j = S_arg[q];
            ++opcount;
            if (S[j] < (max_scores - (double)ai[3])) break;
            double t;
            if (!pair_score(&c, ai, a + j * 4, S[j], &t)) continue;
            if (t > max_scores) { max_scores = t; pre_index = j; }
        }
        S[i] = max_scores;
        P[i] = pre_index;
        if (max_scores > g_max_scores) { g_max_scores = max_scores; g_max_index = i; }
    }
    ret = g_max_index;
done:
    if (opcount_out) *opcount_out = opcount;
    free(gapcost_list);
    return ret;
}

/*
 * Local traceback with overlap trimming (`27508-27527`).  P may be int64 (exact
 * variants) or int32 (fast variants): pass it widened.  path: int64[cap][4] in
 * DESCENDING read order.  Returns the path length.
 */
int64_t orc_local_traceback(const int64_t *a, const int64_t *P, int64_t g_max_index, int64_t *path)
{
    int64_t m = 0;
    int64_t take = g_max_index;
    memcpy(path, a + take * 4, 32); m = 1;
    const int64_t *pre = a + take * 4;
    for (;;) {
        if (P[take] == NOPRE) break;
        take = P[take];
        const int64_t *now = a + take * 4;
        if (pre[0] < now[0] + now[3]) {
            int64_t ov = now[0] + now[3] - pre[0];
            int64_t *last = path + (m - 1) * 4;
            if (pre[2] == 1) { last[0] = pre[0] + ov; last[1] = pre[1] + ov; last[2] = pre[2]; last[3] = pre[3] - ov; }
            else { last[0] = pre[0] + ov; last[1] = pre[1]; last[2] = pre[2]; last[3] = pre[3] - ov; }
        }
        memcpy(path + m * 4, now, 32); ++m;
        pre = now;
    }
    return m;
}
