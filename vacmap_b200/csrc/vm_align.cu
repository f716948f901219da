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

