/*
 * orc_reseed.c -- ORACLE (test infrastructure, not product code).
 *
 * Literal C restatement of the local re-seeding scan
 * get_localmap_multi_all_forDP_inv_guide_1 (mammap_clrnano.py:23069-23345), from the
 * point where the reference windows are known: build the single/multi 9-mer tables
 * over the windows (:23073-23087, 23138-23140), scan every read position forward and
 * reverse-complement (:23207-23341) with the guide-proximity filter (:23216-23231)
 * and the same-diagonal merge (:23235-23252, 23294-23312), and flush the remaining
 * diagonals in first-seen order (:23343-23344).
 * Window construction (:23095-23114, 23142-23154) stays in oracle/pipeline.py.
 * k-mers are compared as strings (the reference uses numba's hash(str) purely as an
 * equality key); here they are packed 8 bits per base into a 72-bit key (k <= 9).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint64_t lo; uint8_t hi; } kkey;   /* 9 bytes of k-mer */

static inline kkey make_key(const char *s, int k)
{
    kkey key; key.lo = 0; key.hi = 0;
    for (int i = 0; i < k && i < 8; ++i) key.lo |= (uint64_t)(uint8_t)s[i] << (8 * i);
    if (k > 8) key.hi = (uint8_t)s[8];
    return key;
}
static inline int key_eq(kkey a, kkey b) { return a.lo == b.lo && a.hi == b.hi; }
static inline uint64_t key_hash(kkey a)
{
    uint64_t h = a.lo * 0x9E3779B97F4A7C15ULL ^ ((uint64_t)a.hi * 0xC2B2AE3D27D4EB4FULL);
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ULL; h ^= h >> 32;
    return h;
}

typedef struct { kkey key; int used; int64_t first; int64_t *more; int32_t n_more, cap_more; } kentry;
typedef struct { kentry *e; uint64_t mask; } ktable;

static kentry *kt_find(ktable *t, kkey key, int create)
{
    uint64_t i = key_hash(key) & t->mask;
    for (;;) {
        kentry *e = &t->e[i];
        if (!e->used) {
            if (!create) return NULL;
            e->used = 1; e->key = key; e->first = -1; e->more = NULL; e->n_more = e->cap_more = 0;
            return e;
        }
        if (key_eq(e->key, key)) return e;
        i = (i + 1) & t->mask;
    }
}

typedef struct { int64_t point; int used; int64_t c0, c1, c2, c3; } pentry;
typedef struct { pentry *e; uint64_t mask; } ptable;
static pentry *pt_find(ptable *t, int64_t point, int create)
{
    uint64_t h = (uint64_t)point * 0x9E3779B97F4A7C15ULL;
    h ^= h >> 31;
    uint64_t i = h & t->mask;
    for (;;) {
        pentry *e = &t->e[i];
        if (!e->used) {
            if (!create) return NULL;
            e->used = 2; e->point = point;   /* 2 = freshly created */
            return e;
        }
        if (e->point == point) return e;
        i = (i + 1) & t->mask;
    }
}

typedef struct { int64_t *rows; int64_t n, cap; } outvec;
static void out_push(outvec *o, int64_t a, int64_t b, int64_t c, int64_t d)
{
    if (o->n == o->cap) { o->cap = o->cap ? o->cap * 2 : 4096; o->rows = (int64_t *)realloc(o->rows, 32 * (size_t)o->cap); }
    int64_t *r = o->rows + o->n * 4; r[0] = a; r[1] = b; r[2] = c; r[3] = d; ++o->n;
}

/* findClosest_1 :17560-17581 on int32 read positions */
static void find_closest(const int32_t *arr, int64_t n, int64_t target, int64_t *b0, int64_t *b1, int64_t *i0, int64_t *i1)
{
    if (target <= arr[0]) { *b0 = *b1 = arr[0] - target; *i0 = *i1 = 0; return; }
    if (target >= arr[n - 1]) { *b0 = *b1 = target - arr[n - 1]; *i0 = *i1 = n - 1; return; }
    int64_t i = 0, j = n;
    while (i < j) {
        int64_t mid = (i + j) / 2;
        if (arr[mid] == target) { *b0 = *b1 = 0; *i0 = *i1 = mid; return; }
        if (target < arr[mid]) j = mid; else i = mid + 1;
    }
    *b0 = llabs((long long)arr[j - 1] - target); *b1 = llabs((long long)arr[j] - target); *i0 = j - 1; *i1 = j;
}

/*
 * ref: concatenated reference (global coordinates); windows [win_lo[w], win_hi[w]) in insertion order.
 * gx/gy: guide anchors sorted by read position (numba argsort order supplied by the caller).
 * Appends anchors to *rows_out (malloc'd, caller frees) and returns the count.
 */
int64_t orc_local_reseed(const char *ref, const int64_t *win_lo, const int64_t *win_hi, int32_t n_win,
                         const int32_t *gx, const int64_t *gy, int64_t n_guide,
                         const char *seq, const char *rc_seq, int64_t L, int32_t k,
                         int64_t readstart, int64_t readend, int64_t **rows_out)
{
    int64_t total = 0;
    for (int w = 0; w < n_win; ++w) total += win_hi[w] - win_lo[w];
    uint64_t cap = 1024;
    while (cap < (uint64_t)total * 2 + 16) cap <<= 1;
    ktable kt; kt.e = (kentry *)calloc(cap, sizeof(kentry)); kt.mask = cap - 1;
    char allN[16]; memset(allN, 'N', 16);
    const kkey skipkey = make_key(allN, k);
    for (int w = 0; w < n_win; ++w) {
        const int64_t lo = win_lo[w], hi = win_hi[w];
        for (int64_t p = lo; p + k <= hi; ++p) {
            kkey key = make_key(ref + p, k);
            if (key_eq(key, skipkey)) continue;
            kentry *e = kt_find(&kt, key, 1);
            if (e->first < 0) e->first = p;      /* onelookuptable_s */
            else {                               /* onelookuptable_m: [first, second, ...] */
                if (e->n_more == e->cap_more) { e->cap_more = e->cap_more ? e->cap_more * 2 : 4; e->more = (int64_t *)realloc(e->more, 8 * (size_t)e->cap_more); }
                e->more[e->n_more++] = p;
            }
        }
    }
    uint64_t pcap = 1024;
    while (pcap < (uint64_t)(readend > readstart ? readend - readstart : 1) * 8 + 16) pcap <<= 1;
    ptable pt; pt.e = (pentry *)calloc(pcap, sizeof(pentry)); pt.mask = pcap - 1;
    int64_t *pkeys = NULL, n_pkeys = 0, cap_pkeys = 0;
    outvec out = {NULL, 0, 0};
    uint64_t n_points = 0;

    for (int64_t iloc = readstart; iloc < readend; ++iloc) {
        /* forward k-mer and the reverse-complement k-mer rc[-(iloc+k):-iloc] ('' at iloc == 0, :23212) */
        const char *fwd = seq + iloc;
        const char *rev = rc_seq + (L - iloc - k);
        const int have_rev = iloc != 0;
        if (have_rev && memcmp(fwd, rev, (size_t)k) == 0) continue;
        int64_t b0, b1, ci0, ci1;
        find_closest(gx, n_guide, iloc, &b0, &b1, &ci0, &ci1);
        int64_t interval = b0 + b1 + 500; if (interval > 2000) interval = 2000;
        const int64_t r1 = gy[ci0], r2 = gy[ci1];
        const int64_t rgap = llabs((long long)iloc - gx[ci0]);
        for (int pass = 0; pass < 2; ++pass) {
            if (pass == 1 && !have_rev) break;
            const int64_t strand = pass == 0 ? 1 : -1;
            kentry *e = kt_find(&kt, make_key(pass == 0 ? fwd : rev, k), 0);
            if (!e) continue;
            const int64_t nloc = 1 + e->n_more;
            for (int64_t t = 0; t < nloc; ++t) {
                const int64_t refloc = t == 0 ? e->first : e->more[t - 1];
                const int64_t diff = llabs(rgap - llabs(refloc - r1));
                if (!(diff < 500 || (r1 + interval >= refloc && r1 - interval <= refloc) ||
                      (r2 + interval >= refloc && r2 - interval <= refloc))) continue;
                const int64_t point = strand == 1 ? refloc - iloc : -(refloc + iloc);
                if (n_points * 2 + 2 > pcap) {   /* grow */
                    ptable nt; uint64_t ncap = pcap * 2; nt.e = (pentry *)calloc(ncap, sizeof(pentry)); nt.mask = ncap - 1;
                    for (uint64_t q = 0; q < pcap; ++q) if (pt.e[q].used) { pentry *ne = pt_find(&nt, pt.e[q].point, 1); *ne = pt.e[q]; ne->used = 1; }
                    free(pt.e); pt = nt; pcap = ncap;
                }
                pentry *pe = pt_find(&pt, point, 1);
                if (pe->used == 2) {
                    pe->used = 1; ++n_points;
                    pe->c0 = iloc; pe->c1 = refloc; pe->c2 = strand; pe->c3 = k;
                    if (n_pkeys == cap_pkeys) { cap_pkeys = cap_pkeys ? cap_pkeys * 2 : 1024; pkeys = (int64_t *)realloc(pkeys, 8 * (size_t)cap_pkeys); }
                    pkeys[n_pkeys++] = point;
                } else if ((pe->c0 + pe->c3) >= iloc) {
                    const int64_t bonus = iloc - (pe->c0 + pe->c3) + k;
                    if (bonus > 0) {
                        if (pe->c3 + bonus < 20) {
                            if (strand == 1) { pe->c2 = 1; pe->c3 += bonus; }
                            else { pe->c1 = refloc; pe->c2 = -1; pe->c3 += bonus; }
                        } else {
                            out_push(&out, pe->c0, pe->c1, pe->c2, pe->c3);
                            if (strand == 1) { const int64_t c3 = pe->c3; pe->c0 += c3; pe->c1 += c3; pe->c2 = 1; pe->c3 = bonus; }
                            else { pe->c0 += pe->c3; pe->c1 = refloc; pe->c2 = -1; pe->c3 = bonus; }
                        }
                    }
                } else {
                    out_push(&out, pe->c0, pe->c1, pe->c2, pe->c3);
                    pe->c0 = iloc; pe->c1 = refloc; pe->c2 = strand; pe->c3 = k;
                }
            }
        }
    }
    for (int64_t q = 0; q < n_pkeys; ++q) {
        pentry *pe = pt_find(&pt, pkeys[q], 0);
        out_push(&out, pe->c0, pe->c1, pe->c2, pe->c3);
    }
    for (uint64_t q = 0; q < cap; ++q) if (kt.e[q].used && kt.e[q].more) free(kt.e[q].more);
    free(kt.e); free(pt.e); free(pkeys);
    *rows_out = out.rows;
    return out.n;
}

void orc_free(void *p) { free(p); }
