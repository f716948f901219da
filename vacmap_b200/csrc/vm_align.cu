// Base-level kernels (integer DP; no tensor cores -- nothing here is a dense contraction).
//
//  * vm_edit_distance_kernel: global unit-cost edit distance, replaces
//    edlib.align(query, target, task='distance') (mammap_clrnano.py:19251).  Myers/Hyyro
//    bit-vector blocks; one warp per job, the pattern's 64-row blocks are spread over the
//    lanes and the columns flow through the lanes as a systolic pipeline (lane l works on
//    column tau - l), the horizontal carry moving lane to lane with one shuffle per step.
//  * vm_extend_kernel: z-drop banded extension, replaces mp.k_cigar(2,-4,4,4,4,4, bw=100,
//    zdrop=50) at :2381,2410,2477,2505 -- only (q_e, t_e) of the best cell are consumed
//    there, so no traceback is kept.  One warp per job, anti-diagonal wavefront, the band
//    (<= 101 cells) lives in a 128-entry shared-memory ring.
//  * vm_fill_kernel: unbanded global dual-affine alignment with traceback, replaces
//    mp.k_cigar(2,-4,4,2,24,1, bw=-1, zdrop=-1, eqx) at :21554,21598.  One warp per job,
//    anti-diagonal wavefront; scores in shared memory, one direction byte per cell written
//    diagonal-major (coalesced) to global memory, traceback by lane 0.
// Recurrences, tie order (diag > E1 > F1 > E2 > F2), left-aligned gaps, z-drop rule and
// boundary conditions are those of oracle/orc_align.c (ksw2 extd2 semantics).
#include "vm_align.cuh"
#include "vm_index.cuh"

#define VM_NEG (-0x40000000)
#define VM_ED_MAXG 64

struct VmSeqView {
    const uint8_t *p;   // address of element 0
    int step;           // +1 / -1
    int comp;
    int len;
};

__device__ __forceinline__ VmSeqView vm_view(const VmSeqSources &S, const VmSeqSpec &s, int read)
{
    const uint8_t *base = s.src == 0 ? S.ref : ((s.src == 1 ? S.reads_fwd : S.reads_rc) + S.read_off[read]);
    VmSeqView v;
    v.len = s.len;
    v.comp = s.comp;
    if (s.reverse) { v.p = base + s.lo + s.len - 1; v.step = -1; }
    else { v.p = base + s.lo; v.step = 1; }
    return v;
}

__device__ __forceinline__ int vm_at(const VmSeqView &v, int i)
{
    int c = vm_code5(__ldg(v.p + (long long)i * v.step));
    if (v.comp && c < 4) c = 3 - c;
    return c;
}

// ---------------------------------------------------------------------------
// edit distance
// ---------------------------------------------------------------------------
// GT > 0: every lane owns GT 64-row blocks whose Pv/Mv words live in REGISTERS (loops fully
// unrolled); GT == 0: generic version for very long patterns, words per lane decided at run time
// (state in local memory).
template <int GT>
__global__ void __launch_bounds__(32) vm_edit_distance_kernel(VmAlnJobDev *jobs, const int *__restrict__ job_ids, VmSeqSources S,
                                                              int max_words)
{
    extern __shared__ unsigned long long vm_peq[];   // [5][max_words]
    VmAlnJobDev &J = jobs[job_ids[blockIdx.x]];
    const int lane = threadIdx.x;
    const VmSeqView pat = vm_view(S, J.q, J.read), txt = vm_view(S, J.t, J.read);
    const int m = pat.len, n = txt.len;
    if (m == 0 || n == 0) { if (lane == 0) J.result0 = m + n; return; }
    const int W = (m + 63) >> 6;
    const int G = GT > 0 ? GT : (W + 31) >> 5;
    const int used = (W + G - 1) / G;
    for (int w = lane; w < W; w += 32) {
        unsigned long long e0 = 0, e1 = 0, e2 = 0, e3 = 0, e4 = 0;
        for (int b = 0; b < 64; ++b) {
            const int i = w * 64 + b;
            if (i < m) {
                const int c = vm_at(pat, i);
                const unsigned long long bit = 1ULL << b;
                if (c == 0) e0 |= bit; else if (c == 1) e1 |= bit; else if (c == 2) e2 |= bit; else if (c == 3) e3 |= bit; else e4 |= bit;
            }
        }
        vm_peq[w] = e0; vm_peq[max_words + w] = e1; vm_peq[2 * max_words + w] = e2; vm_peq[3 * max_words + w] = e3;
        vm_peq[4 * max_words + w] = e4;
    }
    __syncwarp();
    constexpr int NW = GT > 0 ? GT : VM_ED_MAXG;
    unsigned long long Pv[NW], Mv[NW];
#pragma unroll
    for (int g = 0; g < NW; ++g) { Pv[g] = ~0ULL; Mv[g] = 0ULL; }
    int score = 64 * W;
    const int w0 = lane * G;
    const bool owns_last = lane < used && (W - 1) >= w0 && (W - 1) < w0 + G;
    int hout_prev = 0;
    for (int tau = 0; tau < n + used - 1; ++tau) {
        const int hin_from = __shfl_up_sync(VM_FULL, hout_prev, 1);
        const int j = tau - lane;
        int hout = 0;
        if (lane < used && j >= 0 && j < n) {
            int hin = lane == 0 ? 1 : hin_from;   // D[0][j] - D[0][j-1] = 1 (global alignment)
            const unsigned long long *eqrow = vm_peq + vm_at(txt, j) * max_words + w0;
#pragma unroll
            for (int g = 0; g < NW; ++g) {
                if (g < G && w0 + g < W) {
                    unsigned long long Eq = eqrow[g];
                    const unsigned long long pv = Pv[g], mv = Mv[g];
                    const unsigned long long neg = hin < 0 ? 1ULL : 0ULL;
                    const unsigned long long Xv = Eq | mv;
                    Eq |= neg;
                    const unsigned long long Xh = (((Eq & pv) + pv) ^ pv) | Eq;
                    unsigned long long Ph = mv | ~(Xh | pv);
                    unsigned long long Mh = pv & Xh;
                    const int ho = (int)(Ph >> 63) - (int)(Mh >> 63);
                    Ph <<= 1;
                    Mh <<= 1;
                    Mh |= neg;
                    Ph |= hin > 0 ? 1ULL : 0ULL;
                    Pv[g] = Mh | ~(Xv | Ph);
                    Mv[g] = Ph & Xv;
                    hin = ho;
                    if (w0 + g == W - 1) score += ho;
                }
            }
            hout = hin;
        }
        hout_prev = hout;
    }
    if (owns_last) {
        const int gl = (W - 1) - w0;
        unsigned long long pvl = 0, mvl = 0;
#pragma unroll
        for (int g = 0; g < NW; ++g)
            if (g == gl) { pvl = Pv[g]; mvl = Mv[g]; }
        const int first_pad = m - 64 * (W - 1);     // bits >= first_pad of the last word are padding rows
        for (int b = first_pad; b < 64; ++b) {
            if (pvl >> b & 1ULL) --score;
            if (mvl >> b & 1ULL) ++score;
        }
        J.result0 = score;
    }
}

template <int GT>
static void vm_ed_launch_class(VmAlnJobDev *jobs, const int *ids, int n_ids, VmSeqSources src, int max_words, cudaStream_t stream)
{
    if (n_ids <= 0) return;
    const size_t smem = (size_t)max_words * 5 * 8;
    cudaFuncSetAttribute(vm_edit_distance_kernel<GT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    vm_edit_distance_kernel<GT><<<n_ids, 32, smem, stream>>>(jobs, ids, src, max_words);
}

// ids_dev: job indices grouped by class; class_start[0..6]: classes GT = 1, 2, 4, 8, 16, generic;
// class_words[c]: largest pattern word count in class c (sizes the Peq table in shared memory)
int vm_launch_edit_distance(VmAlnJobDev *jobs, const int *ids_dev, const int *class_start, const int *class_words, VmSeqSources src,
                            cudaStream_t stream)
{
    int launches = 0;
    for (int c = 0; c < 6; ++c) {
        const int n_ids = class_start[c + 1] - class_start[c];
        if (n_ids <= 0) continue;
        const int *ids = ids_dev + class_start[c];
        const int mw = class_words[c] > 0 ? class_words[c] : 1;
        switch (c) {
        case 0: vm_ed_launch_class<1>(jobs, ids, n_ids, src, mw, stream); break;
        case 1: vm_ed_launch_class<2>(jobs, ids, n_ids, src, mw, stream); break;
        case 2: vm_ed_launch_class<4>(jobs, ids, n_ids, src, mw, stream); break;
        case 3: vm_ed_launch_class<8>(jobs, ids, n_ids, src, mw, stream); break;
        case 4: vm_ed_launch_class<16>(jobs, ids, n_ids, src, mw, stream); break;
        default: vm_ed_launch_class<0>(jobs, ids, n_ids, src, mw, stream); break;
        }
        ++launches;
    }
    return launches;
}

// ---------------------------------------------------------------------------
// shared cell update (ksw2 extd2 recurrences, see oracle/orc_align.c)
// ---------------------------------------------------------------------------
struct VmGapPar {
    int match, mismatch, q1, e1, q2, e2;
};

__device__ __forceinline__ int vm_boundary_h(const VmGapPar &g, int len)
{
    const int a = -(g.q1 + g.e1 * len), b = -(g.q2 + g.e2 * len);
    return a > b ? a : b;
}

// inputs: hd (diag H or VM_NEG), eu1/eu2 (E from the cell above), fl1/fl2 (F from the cell left)
__device__ __forceinline__ void vm_cell(const VmGapPar &g, int tc, int qc, int hd, int eu1, int fl1, int eu2, int fl2, int &H,
                                        int &E1n, int &F1n, int &E2n, int &F2n, unsigned &dir)
{
    int sc;
    if (tc > 3 || qc > 3) sc = 0;
    else sc = tc == qc ? g.match : g.mismatch;
    int z = hd > VM_NEG / 2 ? hd + sc : VM_NEG;
    unsigned d = 0;
    if (eu1 > z) { d = 1; z = eu1; }
    if (fl1 > z) { d = 2; z = fl1; }
    if (eu2 > z) { d = 3; z = eu2; }
    if (fl2 > z) { d = 4; z = fl2; }
    H = z;
    int o = z - g.q1;
    if (eu1 > o) { d |= 0x08; E1n = eu1 - g.e1; } else E1n = o - g.e1;
    if (fl1 > o) { d |= 0x10; F1n = fl1 - g.e1; } else F1n = o - g.e1;
    o = z - g.q2;
    if (eu2 > o) { d |= 0x20; E2n = eu2 - g.e2; } else E2n = o - g.e2;
    if (fl2 > o) { d |= 0x40; F2n = fl2 - g.e2; } else F2n = o - g.e2;
    dir = d;
}

// ---------------------------------------------------------------------------
// z-drop banded extension (score only)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(32) vm_extend_kernel(VmAlnJobDev *jobs, VmSeqSources S)
{
    __shared__ int sH[3][128], sE1[2][128], sF1[2][128], sE2[2][128], sF2[2][128];
    VmAlnJobDev &J = jobs[blockIdx.x];
    const int lane = threadIdx.x;
    const VmSeqView T = vm_view(S, J.t, J.read), Q = vm_view(S, J.q, J.read);
    const int tlen = T.len, qlen = Q.len;
    if (tlen <= 0 || qlen <= 0) { if (lane == 0) { J.result0 = 0; J.result1 = 0; } return; }
    const VmGapPar g{2, -4, 4, 4, 4, 4};
    const int w = 100, zdrop = 50;
    int gmax = 0, gmax_t = -1, gmax_q = -1;
    int st1 = 1, en1 = 0, st2 = 1, en2 = 0;
    const int n_diag = tlen + qlen - 1;
    for (int r = 0; r < n_diag; ++r) {
        int st = r - qlen + 1 > 0 ? r - qlen + 1 : 0;
        int en = r < tlen - 1 ? r : tlen - 1;
        const int bst = (r - w + 1) >> 1, ben = (r + w) >> 1;
        if (st < bst) st = bst;
        if (en > ben) en = ben;
        const int hc = r % 3, h2 = (r + 1) % 3;   // H buffers: current, r-2  (r-1 is (r+2)%3, not read)
        const int ec = r & 1, ep = ec ^ 1;
        int best = VM_NEG, best_t = 0x7fffffff;
        for (int t0 = st; t0 <= en; t0 += 32) {
            const int t = t0 + lane;
            if (t <= en) {
                const int q = r - t;
                int hd, eu1, eu2, fl1, fl2;
                if (t == 0 && q == 0) hd = 0;
                else if (t == 0) hd = vm_boundary_h(g, q);
                else if (q == 0) hd = vm_boundary_h(g, t);
                else hd = (t - 1 >= st2 && t - 1 <= en2) ? sH[h2][(t - 1) & 127] : VM_NEG;
                if (t == 0) {
                    const int hb = vm_boundary_h(g, q + 1);
                    eu1 = hb - g.q1 - g.e1; eu2 = hb - g.q2 - g.e2;
                } else if (t - 1 >= st1 && t - 1 <= en1) { eu1 = sE1[ep][(t - 1) & 127]; eu2 = sE2[ep][(t - 1) & 127]; }
                else { eu1 = VM_NEG; eu2 = VM_NEG; }
                if (q == 0) {
                    const int hb = vm_boundary_h(g, t + 1);
                    fl1 = hb - g.q1 - g.e1; fl2 = hb - g.q2 - g.e2;
                } else if (t >= st1 && t <= en1) { fl1 = sF1[ep][t & 127]; fl2 = sF2[ep][t & 127]; }
                else { fl1 = VM_NEG; fl2 = VM_NEG; }
                int H, E1n, F1n, E2n, F2n;
                unsigned d;
                vm_cell(g, vm_at(T, t), vm_at(Q, q), hd, eu1, fl1, eu2, fl2, H, E1n, F1n, E2n, F2n, d);
                sH[hc][t & 127] = H;
                sE1[ec][t & 127] = E1n; sF1[ec][t & 127] = F1n; sE2[ec][t & 127] = E2n; sF2[ec][t & 127] = F2n;
                if (H > best) { best = H; best_t = t; }   // ascending t within a lane: first maximum kept
            }
        }
        // diagonal maximum, smallest t on ties
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            const int ob = __shfl_xor_sync(VM_FULL, best, d);
            const int ot = __shfl_xor_sync(VM_FULL, best_t, d);
            if (ob > best || (ob == best && ot < best_t)) { best = ob; best_t = ot; }
        }
        __syncwarp();
        bool stop = false;
        if (st <= en) {
            // ksw_apply_zdrop
            if (best > gmax) { gmax = best; gmax_t = best_t; gmax_q = r - best_t; }
            else if (best_t >= gmax_t && r - best_t >= gmax_q) {
                const int tl = best_t - gmax_t, ql = (r - best_t) - gmax_q;
                const int l = tl > ql ? tl - ql : ql - tl;
                if (gmax - best > zdrop + l * g.e2) stop = true;
            }
        }
        st2 = st1; en2 = en1; st1 = st; en1 = en;
        if (stop) break;
    }
    if (lane == 0) { J.result0 = gmax_q + 1; J.result1 = gmax_t + 1; }
}

int vm_launch_extend(VmAlnJobDev *jobs, int n_jobs, VmSeqSources src, cudaStream_t stream)
{
    if (n_jobs <= 0) return 0;
    vm_extend_kernel<<<n_jobs, 32, 0, stream>>>(jobs, src);
    return 1;
}

// ---------------------------------------------------------------------------
// global fill with traceback: strip-per-lane pipelined wavefront, registers only
//
// Lane l of the warp owns VM_FILL_R consecutive target rows; at pipeline step s it computes the
// column q = s - l of its strip top to bottom in registers (H of the previous column, F1, F2 per
// row; E1/E2 carried down the column), takes H/E1/E2 of the row above from lane l-1 with three
// shuffles, and stores its VM_FILL_R direction bytes as ONE 8-byte word at [step][lane] -- a
// coalesced 256-byte line per warp step.  No shared memory, so 32 warps/SM stay resident and the
// latency of the (global-memory) traceback of one warp hides behind the DP of the others.
// Targets longer than 32*VM_FILL_R rows are processed in bands; the bottom row of a band is
// handed to the next through 3*qlen ints of global scratch.
// ---------------------------------------------------------------------------
#define VM_FILL_R 8
#define VM_FILL_ROWS (32 * VM_FILL_R)

__device__ __forceinline__ long long vm_fill_dir_index(int t, int q, int qlen)
{
    const int band = t / VM_FILL_ROWS, tb = t % VM_FILL_ROWS;
    const int lane = tb / VM_FILL_R, r = tb % VM_FILL_R;
    return (long long)band * (qlen + 32) * VM_FILL_ROWS + (long long)(q + lane) * VM_FILL_ROWS + lane * VM_FILL_R + r;
}

__global__ void __launch_bounds__(128) vm_fill_kernel(VmAlnJobDev *jobs, int n_jobs, VmSeqSources S, int eqx,
                                                      uint8_t *__restrict__ dir_all, int32_t *__restrict__ band_scratch,
                                                      uint32_t *__restrict__ cigar_out)
{
    const int job = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (job >= n_jobs) return;
    VmAlnJobDev &J = jobs[job];
    const int lane = threadIdx.x & 31;
    const VmSeqView T = vm_view(S, J.t, J.read), Q = vm_view(S, J.q, J.read);
    const int tlen = T.len, qlen = Q.len;
    if (tlen <= 0 || qlen <= 0) { if (lane == 0) J.n_out = 0; return; }
    const VmGapPar g{2, -4, 4, 2, 24, 1};
    uint8_t *dir = dir_all + J.dir_off;
    int32_t *bs = J.sc_off >= 0 ? band_scratch + J.sc_off : nullptr;
    const int nbands = (tlen + VM_FILL_ROWS - 1) / VM_FILL_ROWS;
    for (int band = 0; band < nbands; ++band) {
        const int row0 = band * VM_FILL_ROWS + lane * VM_FILL_R;
        int tc[VM_FILL_R], Hleft[VM_FILL_R], F1[VM_FILL_R], F2[VM_FILL_R];
#pragma unroll
        for (int r = 0; r < VM_FILL_R; ++r) {
            const int t = row0 + r;
            tc[r] = t < tlen ? vm_at(T, t) : 4;
            const int hb = vm_boundary_h(g, t + 1);     // H(t, -1)
            Hleft[r] = hb;
            F1[r] = hb - g.q1 - g.e1;
            F2[r] = hb - g.q2 - g.e2;
        }
        int Hdiag_top = row0 == 0 ? 0 : vm_boundary_h(g, row0);   // H(row0-1, -1)
        const int rows_in_band = (tlen - band * VM_FILL_ROWS) < VM_FILL_ROWS ? (tlen - band * VM_FILL_ROWS) : VM_FILL_ROWS;
        const int last_lane = (rows_in_band - 1) / VM_FILL_R;
        uint8_t *dband = dir + (long long)band * (qlen + 32) * VM_FILL_ROWS;
        int outH = 0, outE1 = 0, outE2 = 0, qc_pipe = 4;
        const int nsteps = qlen + last_lane;
        for (int s = 0; s < nsteps; ++s) {
            int inH = __shfl_up_sync(VM_FULL, outH, 1);
            int inE1 = __shfl_up_sync(VM_FULL, outE1, 1);
            int inE2 = __shfl_up_sync(VM_FULL, outE2, 1);
            int qc = __shfl_up_sync(VM_FULL, qc_pipe, 1);
            const int q = s - lane;
            if (lane == 0) {
                qc = s < qlen ? vm_at(Q, s) : 4;
                if (band == 0) {
                    inH = vm_boundary_h(g, q + 1);      // H(-1, q)
                    inE1 = inH - g.q1 - g.e1;
                    inE2 = inH - g.q2 - g.e2;
                } else if (q < qlen) {
                    inH = bs[q];
                    inE1 = bs[qlen + q];
                    inE2 = bs[2 * qlen + q];
                }
            }
            qc_pipe = qc;
            if (q >= 0 && q < qlen && row0 < tlen) {
                int hd = Hdiag_top, e1 = inE1, e2 = inE2, lastH = 0;
                unsigned long long packed = 0;
#pragma unroll
                for (int r = 0; r < VM_FILL_R; ++r) {
                    if (row0 + r < tlen) {
                        const int hold = Hleft[r];
                        int H, E1n, F1n, E2n, F2n;
                        unsigned d;
                        vm_cell(g, tc[r], qc, hd, e1, F1[r], e2, F2[r], H, E1n, F1n, E2n, F2n, d);
                        Hleft[r] = H; F1[r] = F1n; F2[r] = F2n;
                        e1 = E1n; e2 = E2n;
                        hd = hold;
                        lastH = H;
                        packed |= (unsigned long long)d << (8 * r);
                    }
                }
                outH = lastH; outE1 = e1; outE2 = e2;
                Hdiag_top = inH;
                *(unsigned long long *)(dband + (long long)s * VM_FILL_ROWS + lane * VM_FILL_R) = packed;
                if (band + 1 < nbands && lane == 31) { bs[q] = outH; bs[qlen + q] = e1; bs[2 * qlen + q] = e2; }
            }
        }
        __syncwarp();
    }
    if (lane != 0) return;
    // ksw_backtrack (left-aligned), ops pushed in reverse then flipped
    uint32_t *out = cigar_out + J.out_off;
    int n = 0;
    int i = tlen - 1, j = qlen - 1, state = 0;
    while (i >= 0 && j >= 0) {
        const unsigned tmp = dir[vm_fill_dir_index(i, j, qlen)];
        if (state == 0) state = tmp & 7;
        else if (!((tmp >> (state + 2)) & 1)) state = 0;
        if (state == 0) state = tmp & 7;
        unsigned op;
        if (state == 0) {
            op = 0;
            if (eqx) op = vm_at(T, i) == vm_at(Q, j) ? 7 : 8;
            --i; --j;
        } else if (state == 1 || state == 3) { op = 2; --i; }
        else { op = 1; --j; }
        if (n > 0 && (out[n - 1] & 0xf) == op) out[n - 1] += 1u << 4;
        else out[n++] = 1u << 4 | op;
    }
    if (i >= 0) {
        if (n > 0 && (out[n - 1] & 0xf) == 2u) out[n - 1] += (unsigned)(i + 1) << 4;
        else out[n++] = (unsigned)(i + 1) << 4 | 2u;
    }
    if (j >= 0) {
        if (n > 0 && (out[n - 1] & 0xf) == 1u) out[n - 1] += (unsigned)(j + 1) << 4;
        else out[n++] = (unsigned)(j + 1) << 4 | 1u;
    }
    for (int x = 0, y = n - 1; x < y; ++x, --y) { const uint32_t t = out[x]; out[x] = out[y]; out[y] = t; }
    J.n_out = n;
}

size_t vm_fill_dir_bytes(int tlen, int qlen)
{
    const size_t nbands = ((size_t)tlen + VM_FILL_ROWS - 1) / VM_FILL_ROWS;
    return nbands * ((size_t)qlen + 32) * VM_FILL_ROWS;
}
int vm_fill_band_rows() { return VM_FILL_ROWS; }

int vm_launch_fill(VmAlnJobDev *jobs, int n_jobs, VmSeqSources src, int eqx, uint8_t *dir, int32_t *band_scratch,
                   uint32_t *cigar_out, cudaStream_t stream)
{
    if (n_jobs <= 0) return 0;
    vm_fill_kernel<<<(n_jobs + 3) / 4, 128, 0, stream>>>(jobs, n_jobs, src, eqx, dir, band_scratch, cigar_out);
    return 1;
}
