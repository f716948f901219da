// z-drop banded edge extension as a warp-collective device function (used by vm_extend_kernel, one job per warp, and by
// the per-read extension kernel of vm_dglue.cuh, which runs a read's extensions one after the other as the reference does).
// Replaces mp.k_cigar(target, query, 2, -4, 4, 4, 4, 4, bw=100, zdropvalue=50) at mammap_clrnano.py:2381, 2410, 2478, 2504.
#pragma once
#include "vm_align.cuh"
#include "vm_index.cuh"

#ifndef VM_NEG
#define VM_NEG (-0x40000000)
#endif

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
// Shared-memory scratch of one warp's extension: 11 rings of 128 cells (the 101-cell band)
struct VmExtSmem { int sH[3][128], sE1[2][128], sF1[2][128], sE2[2][128], sF2[2][128]; };

// Warp-collective: all 32 lanes call with the same arguments; q_e / t_e (query / target bases consumed at the best
// cell) are returned in every lane.
__device__ __forceinline__ void vm_extend_warp(const VmSeqView &T, const VmSeqView &Q, VmExtSmem &M, int &q_e_out, int &t_e_out)
{
    int (&sH)[3][128] = M.sH;
    int (&sE1)[2][128] = M.sE1, (&sF1)[2][128] = M.sF1, (&sE2)[2][128] = M.sE2, (&sF2)[2][128] = M.sF2;
    const int lane = threadIdx.x & 31;
    const int tlen = T.len, qlen = Q.len;
    q_e_out = 0; t_e_out = 0;
    if (tlen <= 0 || qlen <= 0) return;
    __syncwarp();
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
    q_e_out = gmax_q + 1;
    t_e_out = gmax_t + 1;
    __syncwarp();
}

