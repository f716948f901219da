/*
 * orc_align.c -- ORACLE (test infrastructure, not product code).
 *
 * CPU restatement of the base-level routines the reference calls in absent
 * third-party code:
 *   - `vacmap_index.k_cigar(target, query, match, mismatch, gap_open_1, gap_extend_1,
 *      gap_open_2, gap_extend_2, bw, zdropvalue[, eqx])`
 *        -> (cigar, zdropcode, q_e, t_e, ndel, nins)
 *     call sites mammap_clrnano.py:2381,2410,2477,2505 (edge extension: 2/-4,
 *     4,4,4,4, bw 100, zdrop 50) and :21554,21598 (global fill: 2/-4, 4,2,24,1,
 *     bw -1, zdrop -1, eqx).  Restated from the published ksw2 `ksw_extd2`
 *     dual-affine recurrences (minimap2 2.29 ksw2_extd2_sse.c, ksw2.h
 *     ksw_backtrack / ksw_apply_zdrop): i = target index, j = query index,
 *        H(i,j) = max{H(i-1,j-1)+s, E, F, E2, F2}   ties: diag > E > F > E2 > F2
 *        E(i+1,j) = max{H(i,j)-q, E(i,j)} - e        continue only if E > H-q (left-aligned)
 *        F(i,j+1) = max{H(i,j)-q, F(i,j)} - e        (same for q2/e2)
 *     anti-diagonal band t in [(r-w+1)>>1, (r+w)>>1]; cells outside the band are
 *     -inf; z-drop tested once per anti-diagonal on that diagonal's best cell.
 *     zdrop >= 0 selects extension mode (path ends at the best cell),
 *     zdrop < 0 global mode (path ends at the corner).
 *   - `edlib.align(query, target, task='distance')['editDistance']` (:19251):
 *     global unit-cost edit distance; the value is unique.
 *
 * PARITY UNPINNED against the real extension for CIGAR tie-breaking; pinned:
 * CUDA == this file, bit for bit.  Bases outside ACGT score 0 against anything
 * (ksw_gen_simple_mat's ambiguous row/column); '='/'X' is decided by equality of
 * the 5-letter codes.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NEG_INF (-0x40000000)

static inline int code5(unsigned char c)
{
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': case 'U': case 'u': return 3;
    default: return 4;
    }
}

typedef struct {
    int32_t score;      /* global: H(corner); extension: max */
    int32_t max_t, max_q; /* 0-based best cell (extension); -1,-1 when nothing scored above 0 */
    int32_t zdropped;
    int32_t n_cigar;
    int32_t q_e, t_e;   /* bases consumed by the reported path */
    int32_t ndel, nins;
} orc_kc_result;

/*
 * cigar out: uint32 ops, BAM encoding len<<4|op with op 0=M 1=I 2=D 7='=' 8=X.
 * Returns 0, or -1 when cigar_cap is too small (n_cigar holds the need).
 */
int orc_k_cigar(const char *target, int32_t tlen, const char *query, int32_t qlen,
                int32_t match, int32_t mismatch, int32_t q1, int32_t e1, int32_t q2, int32_t e2,
                int32_t bw, int32_t zdrop, int32_t eqx,
                uint32_t *cigar, int32_t cigar_cap, orc_kc_result *res)
{
    memset(res, 0, sizeof(*res));
    res->max_t = res->max_q = -1;
    if (tlen <= 0 || qlen <= 0) return 0;
    if (mismatch > 0) mismatch = -mismatch;
    const int ext_mode = zdrop >= 0;
    int w = bw;
    if (w < 0) w = tlen > qlen ? tlen : qlen;
    const size_t nrow = (size_t)tlen, ncol = (size_t)qlen;
    /* scores are rolled per anti-diagonal; one direction byte per cell is kept for the traceback */
    uint8_t *dir = (uint8_t *)malloc(nrow * ncol);
    uint8_t *tc = (uint8_t *)malloc(nrow), *qc = (uint8_t *)malloc(ncol);
    for (int i = 0; i < tlen; ++i) tc[i] = (uint8_t)code5((unsigned char)target[i]);
    for (int j = 0; j < qlen; ++j) qc[j] = (uint8_t)code5((unsigned char)query[j]);
    /* Anti-diagonal sweep (so the z-drop rule sees exactly ksw2's order).  Rolling arrays
     * indexed by target index t; for anti-diagonal r the cell is (t, q = r - t):
     *   Hm2[t] = H(t, r-2-t), Hm1[t] = H(t, r-1-t)
     *   E?p[t] = E value computed AT cell (t, r-1-t) for the cell below it, (t+1, r-1-t)
     *   F?p[t] = F value computed AT cell (t, r-1-t) for the cell right of it, (t, r-t)
     * entries of cells that were not computed (outside band / matrix) read as -inf. */
    int32_t *buf = (int32_t *)malloc(4 * nrow * 11);
    int32_t *Hm1 = buf, *Hm2 = buf + nrow, *Hc = buf + 2 * nrow;
    int32_t *E1p = buf + 3 * nrow, *E1c = buf + 4 * nrow, *F1p = buf + 5 * nrow, *F1c = buf + 6 * nrow;
    int32_t *E2p = buf + 7 * nrow, *E2c = buf + 8 * nrow, *F2p = buf + 9 * nrow, *F2c = buf + 10 * nrow;
    for (size_t t = 0; t < nrow * 11; ++t) buf[t] = NEG_INF;
#define BOUNDARY_H(L) ((-(q1 + e1 * (L))) > (-(q2 + e2 * (L))) ? (-(q1 + e1 * (L))) : (-(q2 + e2 * (L))))
    int32_t gmax = 0, gmax_t = -1, gmax_q = -1;
    int zdropped = 0;
    const int n_diag = tlen + qlen - 1;
    for (int r = 0; r < n_diag; ++r) {
        int st = r - qlen + 1 > 0 ? r - qlen + 1 : 0;
        int en = r < tlen - 1 ? r : tlen - 1;
        int bst = (r - w + 1) >> 1, ben = (r + w) >> 1;
        if (st < bst) st = bst;
        if (en > ben) en = ben;
        for (size_t t = 0; t < nrow; ++t) Hc[t] = E1c[t] = F1c[t] = E2c[t] = F2c[t] = NEG_INF;
        int32_t dmax = NEG_INF, dmax_t = -1;
        for (int t = st; t <= en; ++t) {
            const int q = r - t;
            int32_t hd, eu1, eu2, fl1, fl2;
            if (t == 0 && q == 0) hd = 0;
            else if (t == 0) hd = BOUNDARY_H(q);        /* H(-1, q-1) */
            else if (q == 0) hd = BOUNDARY_H(t);        /* H(t-1, -1) */
            else hd = Hm2[t - 1];
            if (t == 0) {
                const int32_t hb = BOUNDARY_H(q + 1);   /* H(-1, q) */
                eu1 = hb - q1 - e1; eu2 = hb - q2 - e2;
            } else { eu1 = E1p[t - 1]; eu2 = E2p[t - 1]; }
            if (q == 0) {
                const int32_t hb = BOUNDARY_H(t + 1);   /* H(t, -1) */
                fl1 = hb - q1 - e1; fl2 = hb - q2 - e2;
            } else { fl1 = F1p[t]; fl2 = F2p[t]; }
            int32_t sc;
            if (tc[t] > 3 || qc[q] > 3) sc = 0;
            else sc = tc[t] == qc[q] ? match : mismatch;
            int32_t z = hd > NEG_INF / 2 ? hd + sc : NEG_INF;
            uint8_t d = 0;
            if (eu1 > z) { d = 1; z = eu1; }
            if (fl1 > z) { d = 2; z = fl1; }
            if (eu2 > z) { d = 3; z = eu2; }
            if (fl2 > z) { d = 4; z = fl2; }
            const int32_t H = z;
            int32_t o;
            o = H - q1;
            if (eu1 > o) { d |= 0x08; E1c[t] = eu1 - e1; } else E1c[t] = o - e1;
            if (fl1 > o) { d |= 0x10; F1c[t] = fl1 - e1; } else F1c[t] = o - e1;
            o = H - q2;
            if (eu2 > o) { d |= 0x20; E2c[t] = eu2 - e2; } else E2c[t] = o - e2;
            if (fl2 > o) { d |= 0x40; F2c[t] = fl2 - e2; } else F2c[t] = o - e2;
            Hc[t] = H;
            dir[(size_t)t * ncol + (size_t)q] = d;
            if (H > dmax) { dmax = H; dmax_t = t; }
        }
        if (ext_mode && dmax_t >= 0) {
            /* ksw_apply_zdrop(ez, 1, max_H, r, max_t, zdrop, e2) */
            if (dmax > gmax) { gmax = dmax; gmax_t = dmax_t; gmax_q = r - dmax_t; }
            else if (dmax_t >= gmax_t && r - dmax_t >= gmax_q) {
                int tl = dmax_t - gmax_t, ql = (r - dmax_t) - gmax_q;
                int l = tl > ql ? tl - ql : ql - tl;
                if (gmax - dmax > zdrop + l * e2) zdropped = 1;
            }
        }
        { int32_t *tmp = Hm2; Hm2 = Hm1; Hm1 = Hc; Hc = tmp; }
        { int32_t *tmp;
          tmp = E1p; E1p = E1c; E1c = tmp; tmp = F1p; F1p = F1c; F1c = tmp;
          tmp = E2p; E2p = E2c; E2c = tmp; tmp = F2p; F2p = F2c; F2c = tmp; }
        if (zdropped) break;
    }
#undef BOUNDARY_H
    int ei, ej;
    if (ext_mode) {
        res->score = gmax; res->max_t = gmax_t; res->max_q = gmax_q; res->zdropped = zdropped;
        ei = gmax_t; ej = gmax_q;
    } else {
        res->score = Hm1[tlen - 1];   /* H(tlen-1, qlen-1): last diagonal rolled into Hm1 */
        ei = tlen - 1; ej = qlen - 1;
    }
    res->t_e = ei + 1; res->q_e = ej + 1;
    /* ksw_backtrack (is_rot, left-aligned) */
    int n = 0, rc = 0;
    uint32_t *rev = (uint32_t *)malloc(4 * (size_t)(tlen + qlen + 2));
#define PUSH(op, len_) do { if (n > 0 && (rev[n - 1] & 0xf) == (uint32_t)(op)) rev[n - 1] += (uint32_t)(len_) << 4; else rev[n++] = (uint32_t)(len_) << 4 | (uint32_t)(op); } while (0)
    int i = ei, j = ej, state = 0;
    while (i >= 0 && j >= 0) {
        uint8_t tmp = dir[(size_t)i * ncol + (size_t)j];
        if (state == 0) state = tmp & 7;
        else if (!(tmp >> (state + 2) & 1)) state = 0;
        if (state == 0) state = tmp & 7;
        if (state == 0) {
            int op = 0;
            if (eqx) op = (tc[i] == qc[j]) ? 7 : 8;
            PUSH(op, 1); --i; --j;
        } else if (state == 1 || state == 3) { PUSH(2, 1); --i; }
        else { PUSH(1, 1); --j; }
    }
    if (i >= 0) PUSH(2, i + 1);
    if (j >= 0) PUSH(1, j + 1);
#undef PUSH
    res->n_cigar = n;
    if (n > cigar_cap) rc = -1;
    else for (int t = 0; t < n; ++t) cigar[t] = rev[n - 1 - t];
    for (int t = 0; t < n; ++t) {
        uint32_t op = rev[t] & 0xf, ln = rev[t] >> 4;
        if (op == 2) res->ndel += (int32_t)ln;
        else if (op == 1) res->nins += (int32_t)ln;
    }
    free(rev); free(dir); free(tc); free(qc);
    free(buf);
    return rc;
}

/*
 * Global unit-cost edit distance, Myers / Hyyro bit-vector blocks (what edlib itself runs for
 * task='distance', mode NW).  Used by the timed CPU baseline so that the baseline is not
 * handicapped by a quadratic scalar DP; orc_edit_distance below (plain DP) pins it in the tests.
 */
int64_t orc_edit_distance_bv(const char *pat, int64_t m, const char *txt, int64_t n)
{
    if (m == 0) return n;
    if (n == 0) return m;
    const int64_t W = (m + 63) / 64;
    uint64_t *peq = (uint64_t *)calloc((size_t)(256 * W), 8);
    uint64_t *Pv = (uint64_t *)malloc(8 * (size_t)W), *Mv = (uint64_t *)calloc((size_t)W, 8);
    for (int64_t i = 0; i < m; ++i) peq[(size_t)(unsigned char)pat[i] * W + i / 64] |= 1ULL << (i % 64);
    for (int64_t w = 0; w < W; ++w) Pv[w] = ~0ULL;
    int64_t score = 64 * W;
    for (int64_t j = 0; j < n; ++j) {
        const uint64_t *eqrow = peq + (size_t)(unsigned char)txt[j] * W;
        int hin = 1;
        for (int64_t w = 0; w < W; ++w) {
            uint64_t Eq = eqrow[w];
            const uint64_t pv = Pv[w], mv = Mv[w];
            const uint64_t neg = hin < 0 ? 1ULL : 0ULL;
            const uint64_t Xv = Eq | mv;
            Eq |= neg;
            const uint64_t Xh = (((Eq & pv) + pv) ^ pv) | Eq;
            uint64_t Ph = mv | ~(Xh | pv);
            uint64_t Mh = pv & Xh;
            const int hout = (int)(Ph >> 63) - (int)(Mh >> 63);
            Ph <<= 1; Mh <<= 1;
            Mh |= neg;
            Ph |= hin > 0 ? 1ULL : 0ULL;
            Pv[w] = Mh | ~(Xv | Ph);
            Mv[w] = Ph & Xv;
            hin = hout;
        }
        score += hin;
    }
    for (int64_t b = m - 64 * (W - 1); b < 64; ++b) {
        if (Pv[W - 1] >> b & 1ULL) --score;
        if (Mv[W - 1] >> b & 1ULL) ++score;
    }
    free(peq); free(Pv); free(Mv);
    return score;
}

/* Global unit-cost edit distance (edlib NW, task='distance').  Plain two-row DP. */
int64_t orc_edit_distance(const char *a, int64_t n, const char *b, int64_t m)
{
    if (n == 0) return m;
    if (m == 0) return n;
    int32_t *prev = (int32_t *)malloc(4 * (size_t)(m + 1)), *cur = (int32_t *)malloc(4 * (size_t)(m + 1));
    for (int64_t j = 0; j <= m; ++j) prev[j] = (int32_t)j;
    for (int64_t i = 1; i <= n; ++i) {
        cur[0] = (int32_t)i;
        const char ca = a[i - 1];
        for (int64_t j = 1; j <= m; ++j) {
            int32_t v = prev[j - 1] + (ca != b[j - 1]);
            int32_t u = prev[j] + 1, l = cur[j - 1] + 1;
            if (u < v) v = u;
            if (l < v) v = l;
            cur[j] = v;
        }
        int32_t *t = prev; prev = cur; cur = t;
    }
    int64_t d = prev[m];
    free(prev); free(cur);
    return d;
}
