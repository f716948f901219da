// Banded global dual-affine fill with an optimality certificate: the same alignment (score, tie order, CIGAR) as
// the full-matrix kernel of vm_fill.cu -- mp.k_cigar(target, query, 2, -4, 4, 2, 24, 1, bw=-1, zdropvalue=-1, eqx)
// at mammap_clrnano.py:21554, 21598 -- for roughly half the cells.
//
// Why it is exact.  Only the diagonals k = j - i in [kmin, kmax] are computed; a cell whose upper (left) neighbour
// lies outside sees -infinity there.  The traced path has the banded optimum S.  A path that touches a diagonal
// above kmax needs nI >= kmax + 1 inserted bases and nI - D deleted ones to come back (D = qlen - tlen), so it
// scores at most 2 (qlen - nI) - g(nI) - g(nI - D), g(n) = min(q1 + e1 n, q2 + e2 n); symmetrically below kmin.
// If S is strictly above both bounds, no path leaving the band reaches S: every cell of the traced path has the
// same value and the same winning predecessor as in the full matrix (a better or tying prefix through the outside
// would, joined with the path's suffix, be an outside path scoring >= S), so the direction walk is identical.
// Jobs that fail the test are reported as such and the host re-runs them in the full-matrix kernel.
//
// Layout.  Anti-diagonal wavefront: at step r the band holds at most 32*C rows t (cells (t, r - t)); row t lives in
// lane (t / C) % 32, register slot t % C, and the slot moves on to row t + 32*C when t leaves the band.  u / y
// (differences towards the left neighbour) stay in the row's registers, v / x (towards the upper neighbour) are
// read from the slot above -- slot c - 1 of the same lane, or slot C - 1 of the previous lane through three
// shuffles per step.  A slot that is not computing publishes (v, x) = (0, -inf), which is exactly what a cell on
// the band's upper edge must see.  Two jobs share a warp, one per half of every half2 register (as in vm_fill.cu);
// they share the band geometry (the union of their bands), cells beyond a job's own matrix are never read back.
// Direction bytes go to the warp's scratch as [step][word][lane] (one 128-byte line per store instruction); the
// traceback stages 32-step tiles of the few lanes it can reach into shared memory, compares the bases itself
// (both sequences are staged in shared memory as fp16 codes) and sums the path's score for the certificate.
#include "vm_fill_cell.cuh"
#include "vm_hostpool.hpp"
#include <algorithm>

namespace {

#define VM_FB_CAP 768          // longest target / query the banded kernel stages in shared memory (whole-warp pairs)
#define VM_FB_CAP2 512         // ... of the half-warp pairs
#define VM_FB_CAP4 384         // ... of the quarter-warp pairs
#define VM_FB_NEG (-1000)      // "-infinity" of the difference recurrences (exact in fp16)

__device__ __forceinline__ int vm_gapcost(int n)
{
    constexpr VmGapPar2 g = vm_fill_par();
    const int a = g.q1 + g.e1 * n, b = g.q2 + g.e2 * n;
    return a < b ? a : b;
}

// upper bound on the score of any path of a (tlen x qlen) job that leaves the band [kmin, kmax]
__device__ __forceinline__ int vm_outside_bound(int tlen, int qlen, int kmin, int kmax)
{
    const int D = qlen - tlen;
    int best = -0x3fffffff;
    const int nI = kmax + 1;
    if (nI <= qlen && nI - D >= 1 && nI - D <= tlen) best = 2 * (qlen - nI) - vm_gapcost(nI) - vm_gapcost(nI - D);
    const int nD = -kmin + 1;
    if (nD <= tlen && nD + D >= 1 && nD + D <= qlen) {
        const int b = 2 * (tlen - nD) - vm_gapcost(nD) - vm_gapcost(nD + D);
        if (b > best) best = b;
    }
    return best;
}

// H(-1, n) - H(-1, n - 1) of the boundary row / column, n = index + 1: the gap pieces cross at 20 bases
// (4 + 2 n <= 24 + n), so the step is -(q1 + e1) for the first base, -e1 up to there and -e2 beyond
__device__ __forceinline__ __half2 vm_boundary_step(int index)
{
    constexpr VmGapPar2 g = vm_fill_par();
    static_assert(g.q1 == 4 && g.e1 == 2 && g.q2 == 24 && g.e2 == 1, "boundary steps are folded for these penalties");
    return index == 0 ? VM_H2C(-(g.q1 + g.e1)) : index < 20 ? VM_H2C(-g.e1) : VM_H2C(-g.e2);
}

__host__ __device__ constexpr int vm_fb_cap(int G) { return G == 1 ? VM_FB_CAP : G == 2 ? VM_FB_CAP2 : VM_FB_CAP4; }

// G pairs of jobs per warp (32 / G lanes per pair), C register slots per lane: the band of a pair holds at most
// (32 / G) * C rows per anti-diagonal.  The quarter-warp classes fit the common 200..350-base segments (their bands
// need 40..60 rows) into 48, 56 or 64 rows instead of 64 or 96, and share the per-step bookkeeping (bounds, three
// shuffles, the store, the loop) between eight jobs.
template <int C, int G>
__global__ void __launch_bounds__(128) vm_fillb_kernel(VmAlnJobDev *jobs, const VmFillBandPair *__restrict__ pairs, int pair_begin,
                                                       int pair_end, VmSeqSources S, int eqx, uint32_t *dir_all,
                                                       long long dir_words_per_warp, int *counter, uint32_t *cigar_out,
                                                       uint32_t *dense_out, unsigned long long *dense_count, uint2 *results)
{
    constexpr VmGapPar2 g = vm_fill_par();
    constexpr int LJ = 32 / G;                                    // lanes of one pair
    constexpr int CAP = vm_fb_cap(G);                             // longest sequence staged
    constexpr int CW = (C + 1) / 2;                               // direction words per lane and step
    constexpr int TS = LJ;                                        // traceback tile: TS steps x TS rows
    constexpr int NL = ((TS - 1) / C + 2) < LJ ? ((TS - 1) / C + 2) : LJ;     // lanes a TS-row tile can touch
    constexpr int TW = NL * CW;                                   // tile words per step
    constexpr int GROUP_WORDS = CAP + 2 * TS * TW;                // [T codes u16 | Q codes u16 | two tiles]
    extern __shared__ uint32_t vm_fb_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / LJ, gl = lane % LJ;
    uint32_t *gbase = vm_fb_smem + ((size_t)warp * G + grp) * GROUP_WORDS;
    uint16_t *sT = reinterpret_cast<uint16_t *>(gbase);           // high byte of the fp16 code: job A low byte, B high
    uint16_t *sQ = sT + CAP;
    uint32_t *tile = gbase + CAP;                                 // [2][TS][TW]
    const long long gw = (long long)blockIdx.x * 4 + warp;
    uint32_t *dir = dir_all + gw * dir_words_per_warp;
    const __half2 neg2 = VM_H2C(VM_FB_NEG), zero2 = VM_H2C(0);
    const __half2 open1 = VM_H2C(-(g.q1 + g.e1)), open2 = VM_H2C(-(g.q2 + g.e2)), ext2 = VM_H2C(-g.e2);
    for (;;) {
        int p = 0;
        if (lane == 0) p = pair_begin + atomicAdd(counter, G);
        p = __shfl_sync(VM_FULL, p, 0);
        if (p >= pair_end) break;
        p += grp;
        const bool live = p < pair_end;                           // the last warp's second half may have no pair
        VmFillBandPair pr;
        if (live) pr = pairs[p];
        else { pr.a = -1; pr.b = -1; pr.kmin = 0; pr.kmax = 0; }
        const bool hasB = pr.b >= 0;
        VmAlnJobDev &JA = jobs[live ? pr.a : 0];
        VmAlnJobDev &JB = jobs[hasB ? pr.b : (live ? pr.a : 0)];
        const VmSeqView TA = vm_view(S, JA.t, JA.read), QA = vm_view(S, JA.q, JA.read);
        const VmSeqView TB = vm_view(S, JB.t, JB.read), QB = vm_view(S, JB.q, JB.read);
        const int tA = live ? TA.len : 0, qA = live ? QA.len : 0, tB = hasB ? TB.len : 0, qB = hasB ? QB.len : 0;
        const int tlen = tA > tB ? tA : tB, qlen = qA > qB ? qA : qB;
        const int kmin = pr.kmin, kmax = pr.kmax;
        // ---------------- stage both sequences as fp16 code bytes (A low byte, B high byte) ----------------
        __syncwarp();
        for (int t = gl; t < tlen; t += LJ)
            sT[t] = (uint16_t)(vm_code_byte(t < tA ? vm_at(TA, t) : 4) | vm_code_byte(t < tB ? vm_at(TB, t) : 4) << 8);
        for (int q = gl; q < qlen; q += LJ)
            sQ[q] = (uint16_t)(vm_code_byte(q < qA ? vm_at(QA, q) : 4) | vm_code_byte(q < qB ? vm_at(QB, q) : 4) << 8);
        __syncwarp();
        // ---------------- forward pass over the anti-diagonals ----------------
        int row[C];
        __half2 tc[C], u[C], y1[C], y2[C], v[C], x1[C], x2[C];
        auto init_row = [&](int c, int t) {
            // the row's first cell is in column 0 (real boundary) or on the band's lower edge (nothing to its left)
            row[c] = t;
            tc[c] = vm_codes_half2(t < tlen ? sT[t] : 0x7f7fu);
            const bool edge = t + kmin > 0;
            u[c] = edge ? zero2 : vm_boundary_step(t);
            y1[c] = edge ? neg2 : open1;
            y2[c] = edge ? neg2 : open2;
        };
#pragma unroll
        for (int c = 0; c < C; ++c) {
            init_row(c, gl * C + c);
            v[c] = zero2; x1[c] = neg2; x2[c] = neg2;
        }
        int nsteps = tlen + qlen - 1;
#pragma unroll
        for (int o = LJ; o < 32; o <<= 1) {                       // every group runs to the longest pair's last step
            const int other = __shfl_xor_sync(VM_FULL, nsteps, o);
            nsteps = nsteps > other ? nsteps : other;
        }
        for (int r = 0; r < nsteps; ++r) {
            int tlo = r - (qlen - 1);
            const int e = r - kmax;                       // ceil((r - kmax) / 2)
            const int tk = e > 0 ? (e + 1) >> 1 : 0;
            if (tlo < tk) tlo = tk;
            if (tlo < 0) tlo = 0;
            int thi = r < tlen - 1 ? r : tlen - 1;
            const int f = r - kmin;                       // floor((r - kmin) / 2), r - kmin >= 0 always
            if ((f >> 1) < thi) thi = f >> 1;
            // what the slot above holds from the previous step: slot C-1 of the previous lane for slot 0
            __half2 inV = __shfl_sync(VM_FULL, v[C - 1], (gl + LJ - 1) & (LJ - 1), LJ);
            __half2 inX1 = __shfl_sync(VM_FULL, x1[C - 1], (gl + LJ - 1) & (LJ - 1), LJ);
            __half2 inX2 = __shfl_sync(VM_FULL, x2[C - 1], (gl + LJ - 1) & (LJ - 1), LJ);
            unsigned d[C];
#pragma unroll
            for (int c = C - 1; c >= 0; --c) {
                if (row[c] < tlo) {
                    // the slot moves on to the row one period further down; its first cell is on the band's lower edge,
                    // or (bands reaching far below the main diagonal) in column 0, more than 20 rows down the boundary
                    const int t = row[c] + LJ * C;
                    row[c] = t;
                    tc[c] = vm_codes_half2(t < tlen ? sT[t] : 0x7f7fu);
                    const bool edge = t + kmin > 0;
                    u[c] = edge ? zero2 : ext2;
                    y1[c] = edge ? neg2 : open1;
                    y2[c] = edge ? neg2 : open2;
                }
                const int t = row[c];
                d[c] = 0u;
                if (t <= thi) {
                    const int j = r - t;
                    __half2 cv, cx1, cx2;
                    if (c > 0) { cv = v[c - 1]; cx1 = x1[c - 1]; cx2 = x2[c - 1]; }
                    else { cv = inV; cx1 = inX1; cx2 = inX2; }
                    if (c == 0 && t == 0) {                // real boundary row (row 0 lives in lane 0, slot 0)
                        cv = vm_boundary_step(j);
                        cx1 = open1;
                        cx2 = open2;
                    }
                    const __half2 qc = vm_codes_half2(sQ[j]);
                    vm_cell2<false>(tc[c], qc, cv, cx1, cx2, u[c], y1[c], y2[c], d[c]);
                    v[c] = cv; x1[c] = cx1; x2[c] = cx2;
                }
            }
            // slots that did not compute this step publish (0, -inf); done after the reads above
#pragma unroll
            for (int c = 0; c < C; ++c)
                if (row[c] > thi) { v[c] = zero2; x1[c] = neg2; x2[c] = neg2; }
            uint32_t *dst = dir + (long long)r * (CW * 32) + lane;
#pragma unroll
            for (int m = 0; m < CW; ++m) {
                const unsigned lo = d[2 * m], hi = (2 * m + 1 < C) ? d[(2 * m + 1 < C) ? 2 * m + 1 : 0] : 0u;
                dst[m * 32] = __byte_perm(lo, hi, 0x6420);           // [A even, B even, A odd, B odd]
            }
        }
        __syncwarp();
        // ---------------- traceback (ksw_backtrack, left-aligned): lane 0 of the pair walks job A, lane 1 job B ----------------
        const int w = gl & 1;
        const int tw = w ? tB : tA, qw = w ? qB : qA;
        const bool walker = gl < 2 && tw > 0 && qw > 0;
        uint32_t *out = cigar_out + (w ? JB.out_off : JA.out_off);
        int i = tw - 1, j = qw - 1, state = 0, n = 0, score = 0;
        unsigned cur_op = 0, cur_len = 0;
        const int sh = w * 8;
        for (;;) {
            const bool need = walker && i >= 0 && j >= 0;
            const unsigned ball = __ballot_sync(VM_FULL, need);
            if (!ball) break;
            const unsigned needmask = (ball >> (grp * LJ)) & 3u;
#pragma unroll
            for (int ws = 0; ws < 2; ++ws) {
                const int ii = __shfl_sync(VM_FULL, i, ws, LJ), jj = __shfl_sync(VM_FULL, j, ws, LJ);
                if (needmask >> ws & 1u) {
                    const int rr = ii + jj - gl;                      // this lane stages one anti-diagonal of the tile
                    const int row_lo = ii > TS - 1 ? ii - (TS - 1) : 0;
                    const int lane_lo = (row_lo / C) & (LJ - 1);
                    if (rr >= 0) {
                        const uint32_t *src = dir + (long long)rr * (CW * 32) + grp * LJ;
                        uint32_t *tl = tile + ((size_t)ws * TS + gl) * TW;
#pragma unroll
                        for (int x = 0; x < NL; ++x)
#pragma unroll
                            for (int m = 0; m < CW; ++m) tl[x * CW + m] = src[m * 32 + ((lane_lo + x) & (LJ - 1))];
                    }
                }
            }
            __syncwarp();
            if (need) {
                const int r_hi = i + j, row_lo = i > TS - 1 ? i - (TS - 1) : 0, blk_lo = row_lo / C;
                const uint32_t *tl = tile + (size_t)w * TS * TW;
                while (i >= 0 && j >= 0) {
                    const int back = r_hi - (i + j);
                    if (back >= TS || i < row_lo) break;
                    const int c = i % C;
                    const unsigned tmp = (tl[back * TW + (i / C - blk_lo) * CW + (c >> 1)] >> (((c & 1) * 2 + w) * 8)) & 0xffu;
                    if (state == 0) state = tmp & 7;
                    else if (!((tmp >> (state + 2)) & 1)) state = 0;
                    if (state == 0) state = tmp & 7;
                    unsigned op;
                    if (state == 0) {
                        const unsigned a = (sT[i] >> sh) & 0xffu, b = (sQ[j] >> sh) & 0xffu;
                        const bool same = a == b;
                        if (a != 0x7fu && b != 0x7fu) score += same ? g.match : g.mismatch;
                        op = eqx ? (same ? 7u : 8u) : 0u;
                        --i; --j;
                    } else if (state == 1 || state == 3) { op = 2; --i; }
                    else { op = 1; --j; }
                    if (op == cur_op) ++cur_len;
                    else {
                        if (cur_len) {
                            out[n++] = cur_len << 4 | cur_op;
                            if (cur_op == 1u || cur_op == 2u) score -= vm_gapcost((int)cur_len);
                        }
                        cur_op = op;
                        cur_len = 1;
                    }
                }
            }
            __syncwarp();
        }
        bool certified = true;
        if (walker) {
            if (i >= 0) {
                if (cur_len && cur_op == 2u) cur_len += (unsigned)(i + 1);
                else {
                    if (cur_len) { out[n++] = cur_len << 4 | cur_op; if (cur_op == 1u) score -= vm_gapcost((int)cur_len); }
                    cur_op = 2u;
                    cur_len = (unsigned)(i + 1);
                }
            }
            if (j >= 0) {
                if (cur_len && cur_op == 1u) cur_len += (unsigned)(j + 1);
                else {
                    if (cur_len) { out[n++] = cur_len << 4 | cur_op; if (cur_op == 2u) score -= vm_gapcost((int)cur_len); }
                    cur_op = 1u;
                    cur_len = (unsigned)(j + 1);
                }
            }
            if (cur_len) {
                out[n++] = cur_len << 4 | cur_op;
                if (cur_op == 1u || cur_op == 2u) score -= vm_gapcost((int)cur_len);
            }
            certified = score > vm_outside_bound(tw, qw, kmin, kmax);
        }
        __syncwarp();
        // ops were pushed end to start: claim room in the dense CIGAR arena and copy them over flipped, the pair's lanes helping
#pragma unroll
        for (int ws = 0; ws < 2; ++ws) {
            const int ok = __shfl_sync(VM_FULL, certified ? 1 : 0, ws, LJ);
            const int nw = __shfl_sync(VM_FULL, walker ? n : 0, ws, LJ);
            const int nn = ok ? nw : 0;
            const int jid = ws ? pr.b : pr.a;                        // -1: no such job in this pair
            unsigned long long base = 0;
            if (gl == 0 && nn > 0) base = atomicAdd(dense_count, (unsigned long long)nn);
            base = __shfl_sync(VM_FULL, base, 0, LJ);
            if (jid >= 0) {
                const uint32_t *o = cigar_out + (ws ? JB.out_off : JA.out_off);
                for (int x = gl; x < nn; x += LJ) dense_out[base + x] = o[nn - 1 - x];
                if (gl == 0) results[jid] = ok ? make_uint2((unsigned)base, (unsigned)nn) : make_uint2(0xffffffffu, 0u);
            }
        }
        __syncwarp();
    }
}

template <int C, int G>
size_t vm_fillb_smem()
{
    constexpr int LJ = 32 / G, CAP = vm_fb_cap(G), CW = (C + 1) / 2;
    constexpr int NL = ((LJ - 1) / C + 2) < LJ ? ((LJ - 1) / C + 2) : LJ;
    return (size_t)4 * G * (CAP + 2 * LJ * NL * CW) * sizeof(uint32_t);
}

template <int C, int G>
int vm_fillb_occupancy()
{
    vm_smem_optin(vm_fillb_kernel<C, G>);
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, vm_fillb_kernel<C, G>, 128, vm_fillb_smem<C, G>()) != cudaSuccess || nb < 1) nb = 1;
    return nb;
}

// Slot classes by band rows per anti-diagonal: quarter-warp pairs up to 64 rows, half-warp pairs up to 96, whole-warp
// pairs beyond (and for sequences longer than the narrower groups stage)
struct VmFbClass { int G, C; };
constexpr VmFbClass VM_FB_CLASS[VM_FB_NCLASS + 1] = {{0, 0}, {4, 4}, {4, 6}, {4, 7}, {4, 8}, {2, 5}, {2, 6}, {1, 3},
                                                     {1, 4}, {1, 5}, {1, 6}, {1, 7}, {1, 8}};
inline int vm_fb_rows(int k) { return 32 / VM_FB_CLASS[k].G * VM_FB_CLASS[k].C; }
// smallest class that holds `rows` band rows and sequences of up to `mx` bases (0: none)
inline int vm_fb_class_of(int rows, int mx)
{
    for (int k = 1; k <= VM_FB_NCLASS; ++k)
        if (rows <= vm_fb_rows(k) && mx <= vm_fb_cap(VM_FB_CLASS[k].G)) return k;
    return 0;
}

#define VM_FB_DISPATCH(K, WHAT)                 \
    switch (K) {                                \
    case 1: WHAT(4, 4); break;                  \
    case 2: WHAT(6, 4); break;                  \
    case 3: WHAT(7, 4); break;                  \
    case 4: WHAT(8, 4); break;                  \
    case 5: WHAT(5, 2); break;                  \
    case 6: WHAT(6, 2); break;                  \
    case 7: WHAT(3, 1); break;                  \
    case 8: WHAT(4, 1); break;                  \
    case 9: WHAT(5, 1); break;                  \
    case 10: WHAT(6, 1); break;                 \
    case 11: WHAT(7, 1); break;                 \
    default: WHAT(8, 1); break;                 \
    }

int vm_fillb_blocks_per_sm(int k)
{
    int nb = 1;
#define VM_FB_OCC(CC, GG) nb = vm_fillb_occupancy<CC, GG>()
    VM_FB_DISPATCH(k, VM_FB_OCC)
#undef VM_FB_OCC
    return nb;
}

} // namespace

// The band a job gets on its own: half-width w around its corridor [min(0, D), max(0, D)], wide enough that a read
// with ~10 % errors certifies with margin; jobs too small to gain, or too long for the shared-memory staging, are
// left to the full-matrix kernel.
bool vm_fillb_own_band(int tlen, int qlen, int &kmin, int &kmax)
{
    const int mn = tlen < qlen ? tlen : qlen, mx = tlen > qlen ? tlen : qlen;
    if (mn < 96 || mx > VM_FB_CAP) return false;
    const int D = qlen - tlen;
    // a read at ~10 % error scores ~1.3 per base, a path leaving the band at most 2 (mn - w) - 2 (24 + w): this w
    // leaves ~0.15 per base of margin; the planner then widens the band to the capacity of its slot class
    int w = (int)(0.2125 * mn) - 12;
    if (w < 24) w = 24;
    kmin = (D < 0 ? D : 0) - w;
    kmax = (D > 0 ? D : 0) + w;
    return vm_fb_class_of((kmax - kmin) / 2 + 1, mx) > 0;
}

// Pairs of jobs with (nearly) the same band.  Jobs are bucketed by (slot class of their own band, D / 8, qlen / 8) with a
// parallel stable counting sort, neighbours in that order share a warp; a pair's band is the union of its jobs'
// bands, widened to the capacity of its slot class.  full_mask[j] is cleared for every job planned here.
void vm_fillb_plan(const VmAlnJobDev *J, int nj, int sm_count, int host_threads, VmFillBandPlan &plan, uint8_t *full_mask)
{
    plan.pairs.clear();
    plan.launches.clear();
    plan.dir_words = 0;
    plan.dir_bytes = 0;
    constexpr int ND = 256, NQ = VM_FB_CAP / 8 + 1, NKEY = VM_FB_NCLASS * ND * NQ;     // D / 8 in [-128, 128) covers |D| < 1024
    const int T = std::max(1, std::min(std::min(host_threads, 8), nj / 8192 + 1));
    std::vector<int32_t> keys((size_t)nj), bmin((size_t)nj), bmax((size_t)nj);      // key and own band of every job
    std::vector<std::vector<int32_t>> hist((size_t)T, std::vector<int32_t>((size_t)NKEY, 0));
    auto slice = [&](int t, int &lo, int &hi) { lo = (int)((long long)nj * t / T); hi = (int)((long long)nj * (t + 1) / T); };
    vmp::parallel_for(T, T, [&](int64_t t) {
        int lo, hi;
        slice((int)t, lo, hi);
        std::vector<int32_t> &h = hist[(size_t)t];
        for (int j = lo; j < hi; ++j) {
            int kmin, kmax;
            keys[j] = -1;
            if (J[j].t.len <= 0 || J[j].q.len <= 0 || !vm_fillb_own_band(J[j].t.len, J[j].q.len, kmin, kmax)) continue;
            bmin[j] = kmin;
            bmax[j] = kmax;
            const int c = vm_fb_class_of((kmax - kmin) / 2 + 1, std::max(J[j].t.len, J[j].q.len));
            int db = ((J[j].q.len - J[j].t.len) >> 3) + ND / 2;
            db = db < 0 ? 0 : db >= ND ? ND - 1 : db;
            keys[j] = ((c - 1) * ND + db) * NQ + (J[j].q.len >> 3);
            ++h[(size_t)keys[j]];
            full_mask[j] = 0;
        }
    }, 1);
    int32_t n_live = 0;
    int32_t class_lo[VM_FB_NCLASS + 2];
    for (int k = 0; k < NKEY; ++k) {
        if (k % (ND * NQ) == 0) class_lo[k / (ND * NQ) + 1] = n_live;     // first position of slot class c = k / (ND NQ) + 1
        for (int t = 0; t < T; ++t) {
            const int32_t cnt = hist[(size_t)t][(size_t)k];
            hist[(size_t)t][(size_t)k] = n_live;
            n_live += cnt;
        }
    }
    class_lo[VM_FB_NCLASS + 1] = n_live;
    std::vector<int32_t> order((size_t)n_live);
    vmp::parallel_for(T, T, [&](int64_t t) {
        int lo, hi;
        slice((int)t, lo, hi);
        std::vector<int32_t> &pos = hist[(size_t)t];
        for (int j = lo; j < hi; ++j)
            if (keys[j] >= 0) order[(size_t)pos[(size_t)keys[j]]++] = j;
    }, 1);
    // neighbours of the same slot class share a warp (the widest class runs its jobs alone: a union could need a
    // ninth slot); every pair is built independently, then bucketed by the slot class of its union
    struct Tmp { VmFillBandPair pr; int c, steps; };
    std::vector<int64_t> pair_lo(VM_FB_NCLASS + 2, 0);
    for (int c = 1; c <= VM_FB_NCLASS; ++c) {
        const int64_t n = class_lo[c + 1] - class_lo[c];
        pair_lo[c + 1] = pair_lo[c] + (c < VM_FB_NCLASS ? (n + 1) / 2 : n);
    }
    std::vector<Tmp> tmp((size_t)pair_lo[VM_FB_NCLASS + 1]);
    for (int c = 1; c <= VM_FB_NCLASS; ++c) {
        const int64_t np = pair_lo[c + 1] - pair_lo[c];
        const int lo = class_lo[c], hi = class_lo[c + 1];
        const bool alone = c == VM_FB_NCLASS;
        vmp::parallel_for(np, host_threads, [&](int64_t p) {
            Tmp &t = tmp[(size_t)(pair_lo[c] + p)];
            const int xa = alone ? lo + (int)p : lo + 2 * (int)p, xb = alone ? hi : xa + 1;
            VmFillBandPair &pr = t.pr;
            pr.a = order[(size_t)xa];
            pr.b = xb < hi ? order[(size_t)xb] : -1;
            pr.kmin = bmin[(size_t)pr.a];
            pr.kmax = bmax[(size_t)pr.a];
            t.steps = J[pr.a].t.len + J[pr.a].q.len;
            int mx = std::max(J[pr.a].t.len, J[pr.a].q.len);
            if (pr.b >= 0) {
                mx = std::max(mx, std::max(J[pr.b].t.len, J[pr.b].q.len));
                pr.kmin = std::min(pr.kmin, bmin[(size_t)pr.b]);
                pr.kmax = std::max(pr.kmax, bmax[(size_t)pr.b]);
                t.steps = std::max(J[pr.a].t.len, J[pr.b].t.len) + std::max(J[pr.a].q.len, J[pr.b].q.len);
            }
            const int rows = (pr.kmax - pr.kmin) / 2 + 1;
            t.c = vm_fb_class_of(rows, mx);
            if (t.c == 0) {          // cannot happen for neighbours of one bucket; if it does, the full-matrix kernel takes them
                full_mask[pr.a] = 1;
                if (pr.b >= 0) full_mask[pr.b] = 1;
                return;
            }
            // widen the band to the capacity of its slot class: free rows, more margin for the certificate
            const int spare = vm_fb_rows(t.c) - rows;
            pr.kmin -= spare;
            pr.kmax += spare;
        }, 4096);
    }
    std::vector<int64_t> cnt(VM_FB_NCLASS + 2, 0), at(VM_FB_NCLASS + 2, 0);
    std::vector<int> max_steps(VM_FB_NCLASS + 2, 0);
    for (const Tmp &t : tmp) {
        ++cnt[(size_t)t.c];
        max_steps[(size_t)t.c] = std::max(max_steps[(size_t)t.c], t.steps);
        if (t.c > 0) plan.dir_bytes += (double)t.steps * ((VM_FB_CLASS[t.c].C + 1) / 2) * 128.0 / VM_FB_CLASS[t.c].G;
    }
    cnt[0] = 0;
    for (int c = 1; c <= VM_FB_NCLASS; ++c) at[c + 1] = at[c] + cnt[c];
    plan.pairs.resize((size_t)at[VM_FB_NCLASS + 1]);
    {
        std::vector<int64_t> pos(at);
        for (const Tmp &t : tmp)
            if (t.c > 0) plan.pairs[(size_t)pos[(size_t)t.c]++] = t.pr;
    }
    for (int c = 1; c <= VM_FB_NCLASS; ++c) {
        if (cnt[c] == 0) continue;
        VmFillBandLaunch L;
        L.cls = c;
        L.pair_begin = (int)at[c];
        L.pair_end = (int)at[c + 1];
        L.dir_words_per_warp = (long long)max_steps[(size_t)c] * ((VM_FB_CLASS[c].C + 1) / 2) * 32;
        const int n_pairs = L.pair_end - L.pair_begin;
        const int per_block = 4 * VM_FB_CLASS[c].G;
        L.blocks = (int)std::max<long long>(1, std::min<long long>((n_pairs + per_block - 1) / per_block, (long long)sm_count * vm_fillb_blocks_per_sm(c)));
        plan.dir_words += (size_t)((long long)L.blocks * 4 * L.dir_words_per_warp);      // every launch has its own slice
        plan.launches.push_back(L);
    }
}

int vm_fillb_launch(const VmFillBandPlan &plan, VmAlnJobDev *jobs, const VmFillBandPair *pairs, VmSeqSources src, int eqx, uint32_t *dir,
                    int *counters, uint32_t *cigar_scratch, uint32_t *dense_out, unsigned long long *dense_count, void *results,
                    cudaStream_t main_stream, const cudaStream_t *side, int n_side, int *side_rr, size_t *dir_cursor)
{
    int n = 0;
    for (size_t li = 0; li < plan.launches.size(); ++li) {
        const VmFillBandLaunch &L = plan.launches[li];
        int *ctr = counters + li;
        // launches of a few blocks (rare job classes) are one warp's latency deep: they go to side streams, beside
        // the launches that fill the device; every launch gets its own slice of the direction scratch
        const bool small = n_side > 0 && L.blocks < VM_FILL_SMALL_BLOCKS;
        cudaStream_t stream = small ? side[(*side_rr)++ % n_side] : main_stream;
        uint32_t *dir_l = dir + *dir_cursor;
        *dir_cursor += (size_t)L.blocks * 4 * (size_t)L.dir_words_per_warp;
        uint32_t *dir_save = dir;
        dir = dir_l;
#define VM_FILLB_GO(CC, GG)                                                                                                        \
    vm_fillb_kernel<CC, GG><<<L.blocks, 128, vm_fillb_smem<CC, GG>(), stream>>>(jobs, pairs, L.pair_begin, L.pair_end, src, eqx, dir, \
                                                                                L.dir_words_per_warp, ctr, cigar_scratch,          \
                                                                                dense_out, dense_count, (uint2 *)results)
        VM_FB_DISPATCH(L.cls, VM_FILLB_GO)
#undef VM_FILLB_GO
        dir = dir_save;
        ++n;
    }
    return n;
}

// ---------------------------------------------------------------------------------------------------------------
// The same plan made on the device (vm_fillb_plan above is the host version the stage-level entry points use).
// ---------------------------------------------------------------------------------------------------------------
#include "vm_devsort.cuh"

namespace {

constexpr int FB_ND = 256, FB_NQ = VM_FB_CAP / 8 + 1, FB_NKEY = VM_FB_NCLASS * FB_ND * FB_NQ;
constexpr int FB_NB2 = 4096, FB_NKEY2 = (VM_FB_NCLASS + 1) * FB_NB2;

__host__ __device__ inline int fb_rows_d(int k)
{
    const int G = k <= 4 ? 4 : k <= 6 ? 2 : 1;
    const int C = k == 1 ? 4 : k == 2 ? 6 : k == 3 ? 7 : k == 4 ? 8 : k == 5 ? 5 : k == 6 ? 6 : k - 4;      // 7..12 -> 3..8
    return 32 / G * C;
}
__host__ __device__ inline int fb_cap_d(int k) { return k <= 4 ? VM_FB_CAP4 : k <= 6 ? VM_FB_CAP2 : VM_FB_CAP; }
__host__ __device__ inline int fb_class_of_d(int rows, int mx)
{
    for (int k = 1; k <= VM_FB_NCLASS; ++k)
        if (rows <= fb_rows_d(k) && mx <= fb_cap_d(k)) return k;
    return 0;
}
__device__ inline bool fb_own_band_d(int tlen, int qlen, int &kmin, int &kmax)
{
    const int mn = tlen < qlen ? tlen : qlen, mx = tlen > qlen ? tlen : qlen;
    if (mn < 96 || mx > VM_FB_CAP) return false;
    const int D = qlen - tlen;
    int w = (int)(0.2125 * mn) - 12;
    if (w < 24) w = 24;
    kmin = (D < 0 ? D : 0) - w;
    kmax = (D > 0 ? D : 0) + w;
    return fb_class_of_d((kmax - kmin) / 2 + 1, mx) > 0;
}

struct FbSmall { int32_t class_lo[16], pair_lo[16]; VmFbTable tab; };

__global__ void vm_fbp_key_kernel(const VmAlnJobDev *__restrict__ J, int nj, int32_t *keys, int32_t *bmin, int32_t *bmax, uint8_t *full_mask)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nj) return;
    const int tl = J[j].t.len, ql = J[j].q.len;
    int key = -1;
    uint8_t full = 0;
    if (tl > 0 && ql > 0) {
        int kmin, kmax;
        if (fb_own_band_d(tl, ql, kmin, kmax)) {
            bmin[j] = kmin;
            bmax[j] = kmax;
            const int c = fb_class_of_d((kmax - kmin) / 2 + 1, tl > ql ? tl : ql);
            int db = ((ql - tl) >> 3) + FB_ND / 2;
            db = db < 0 ? 0 : db >= FB_ND ? FB_ND - 1 : db;
            key = ((c - 1) * FB_ND + db) * FB_NQ + (ql >> 3);
        } else full = 1;
    }
    keys[j] = key;
    full_mask[j] = full;
}

__global__ void vm_fbp_bounds_kernel(const int32_t *__restrict__ start, FbSmall *sm)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int c = 1; c <= VM_FB_NCLASS; ++c) sm->class_lo[c] = start[(c - 1) * FB_ND * FB_NQ];
    sm->class_lo[VM_FB_NCLASS + 1] = start[FB_NKEY];
    sm->pair_lo[1] = 0;
    for (int c = 1; c <= VM_FB_NCLASS; ++c) {
        const int n = sm->class_lo[c + 1] - sm->class_lo[c];
        sm->pair_lo[c + 1] = sm->pair_lo[c] + (c < VM_FB_NCLASS ? (n + 1) / 2 : n);
    }
    for (int c = 0; c < 16; ++c) { sm->tab.cnt[c] = 0; sm->tab.at[c] = 0; sm->tab.max_steps[c] = 0; sm->tab.sum_steps[c] = 0ULL; }
    sm->tab.n_live = sm->class_lo[VM_FB_NCLASS + 1];
    sm->tab.n_pairs = 0;
}

// neighbours of the same slot class share a warp; a pair's band is the union of its jobs' bands, widened to the capacity
// of the union's slot class
__global__ void vm_fbp_pair_kernel(const VmAlnJobDev *__restrict__ J, int n_slots_max, const int32_t *__restrict__ order,
                                   const int32_t *__restrict__ bmin, const int32_t *__restrict__ bmax, FbSmall *sm, VmFillBandPair *tpairs,
                                   int32_t *keys2, uint8_t *full_mask)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_slots_max) return;
    const int n_slots = sm->pair_lo[VM_FB_NCLASS + 1];
    if (p >= n_slots) { keys2[p] = -1; return; }
    int c = 1;
    while (c < VM_FB_NCLASS && p >= sm->pair_lo[c + 1]) ++c;
    const int lo = sm->class_lo[c], hi = sm->class_lo[c + 1];
    const bool alone = c == VM_FB_NCLASS;
    const int rel = p - sm->pair_lo[c];
    const int xa = alone ? lo + rel : lo + 2 * rel, xb = alone ? hi : xa + 1;
    VmFillBandPair pr;
    pr.a = order[xa];
    pr.b = xb < hi ? order[xb] : -1;
    pr.kmin = bmin[pr.a];
    pr.kmax = bmax[pr.a];
    int steps = J[pr.a].t.len + J[pr.a].q.len;
    int mx = max(J[pr.a].t.len, J[pr.a].q.len);
    if (pr.b >= 0) {
        mx = max(mx, max(J[pr.b].t.len, J[pr.b].q.len));
        pr.kmin = min(pr.kmin, bmin[pr.b]);
        pr.kmax = max(pr.kmax, bmax[pr.b]);
        steps = max(J[pr.a].t.len, J[pr.b].t.len) + max(J[pr.a].q.len, J[pr.b].q.len);
    }
    const int rows = (pr.kmax - pr.kmin) / 2 + 1;
    const int tc = fb_class_of_d(rows, mx);
    if (tc == 0) {      // cannot happen for neighbours of one bucket; if it does, the full-matrix kernel takes them
        full_mask[pr.a] = 1;
        if (pr.b >= 0) full_mask[pr.b] = 1;
        keys2[p] = -1;
        return;
    }
    const int spare = fb_rows_d(tc) - rows;
    pr.kmin -= spare;
    pr.kmax += spare;
    tpairs[p] = pr;
    // second bucket sort: by the union's class, then coarsely by position so that neighbours stay neighbours
    keys2[p] = tc * FB_NB2 + (int)((long long)p * FB_NB2 / n_slots);
    atomicMax(&sm->tab.max_steps[tc], steps);
    atomicAdd(&sm->tab.sum_steps[tc], (unsigned long long)steps);
}

__global__ void vm_fbp_place_kernel(int n_slots_max, const int32_t *__restrict__ order2, const int32_t *__restrict__ start2,
                                    const VmFillBandPair *__restrict__ tpairs, VmFillBandPair *pairs_out, FbSmall *sm)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = start2[FB_NKEY2];
    if (i == 0) {
        for (int c = 1; c <= VM_FB_NCLASS; ++c) {
            sm->tab.at[c] = start2[c * FB_NB2];
            sm->tab.cnt[c] = start2[(c + 1) * FB_NB2] - start2[c * FB_NB2];
        }
        sm->tab.n_pairs = n;
    }
    if (i >= n_slots_max || i >= n) return;
    pairs_out[i] = tpairs[order2[i]];
}

} // namespace

int vm_fillb_plan_dev(const VmAlnJobDev *J, int nj, VmFillPlanBufs &B, VmFillBandPair *pairs_out, uint8_t *full_mask, cudaStream_t stream)
{
    if (nj <= 0) return 0;
    const size_t n = (size_t)nj;
    if (B.keys.ensure(n * 4) || B.bmin.ensure(n * 4) || B.bmax.ensure(n * 4) || B.order.ensure(n * 4) || B.tpairs.ensure(n * sizeof(VmFillBandPair)) ||
        B.keys2.ensure(n * 4) || B.order2.ensure(n * 4) || B.start.ensure(((size_t)FB_NKEY + 1) * 4) || B.cursor.ensure((size_t)FB_NKEY * 4) ||
        B.start2.ensure(((size_t)FB_NKEY2 + 1) * 4) || B.cursor2.ensure((size_t)FB_NKEY2 * 4) || B.small.ensure(sizeof(FbSmall) + 1024) ||
        B.table.ensure(sizeof(VmFbTable) + sizeof(VmFfTable) + 64))
        return -1;
    int launches = 0;
    const int nb = (nj + 127) / 128;
    FbSmall *sm = B.small.as<FbSmall>();
    vm_fbp_key_kernel<<<nb, 128, 0, stream>>>(J, nj, B.keys.as<int32_t>(), B.bmin.as<int32_t>(), B.bmax.as<int32_t>(), full_mask);
    launches += 1 + vm_bucket_sort(B.keys.as<int32_t>(), nj, FB_NKEY, B.start.as<int32_t>(), B.cursor.as<int32_t>(), B.order.as<int32_t>(), stream);
    vm_fbp_bounds_kernel<<<1, 32, 0, stream>>>(B.start.as<int32_t>(), sm);
    vm_fbp_pair_kernel<<<nb, 128, 0, stream>>>(J, nj, B.order.as<int32_t>(), B.bmin.as<int32_t>(), B.bmax.as<int32_t>(), sm,
                                               B.tpairs.as<VmFillBandPair>(), B.keys2.as<int32_t>(), full_mask);
    launches += 2 + vm_bucket_sort(B.keys2.as<int32_t>(), nj, FB_NKEY2, B.start2.as<int32_t>(), B.cursor2.as<int32_t>(), B.order2.as<int32_t>(), stream);
    vm_fbp_place_kernel<<<nb, 128, 0, stream>>>(nj, B.order2.as<int32_t>(), B.start2.as<int32_t>(), B.tpairs.as<VmFillBandPair>(), pairs_out, sm);
    ++launches;
    if (cudaMemcpyAsync(B.table.p, &sm->tab, sizeof(VmFbTable), cudaMemcpyDeviceToHost, stream) != cudaSuccess) return -1;
    return launches;
}

void vm_fillb_plan_finish(const VmFillPlanBufs &B, int sm_count, VmFillBandPlan &plan)
{
    plan.pairs.clear();
    plan.launches.clear();
    plan.dir_words = 0;
    plan.dir_bytes = 0;
    const VmFbTable &T = *B.table.as<VmFbTable>();
    for (int c = 1; c <= VM_FB_NCLASS; ++c) {
        if (T.cnt[c] <= 0) continue;
        plan.dir_bytes += (double)T.sum_steps[c] * ((VM_FB_CLASS[c].C + 1) / 2) * 128.0 / VM_FB_CLASS[c].G;
        VmFillBandLaunch L;
        L.cls = c;
        L.pair_begin = T.at[c];
        L.pair_end = T.at[c] + T.cnt[c];
        L.dir_words_per_warp = (long long)T.max_steps[c] * ((VM_FB_CLASS[c].C + 1) / 2) * 32;
        const int n_pairs = T.cnt[c];
        const int per_block = 4 * VM_FB_CLASS[c].G;
        static const double frac = getenv("VM_FILL_SM_FRAC") ? atof(getenv("VM_FILL_SM_FRAC")) : 1.0;   // experiment knob
        L.blocks = (int)std::max<long long>(1, std::min<long long>((n_pairs + per_block - 1) / per_block,
                                                                   (long long)(frac * sm_count * vm_fillb_blocks_per_sm(c))));
        plan.dir_words += (size_t)((long long)L.blocks * 4 * L.dir_words_per_warp);
        plan.launches.push_back(L);
    }
}
