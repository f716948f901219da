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

#ifndef VM_NEG
#define VM_NEG (-0x40000000)
#endif

// ---------------------------------------------------------------------------
// edit distance
// ---------------------------------------------------------------------------
// Banded Myers/Hyyro bit-vector NW distance, one warp per job.
//  * The pattern's 64-row blocks are dealt to the lanes CYCLICALLY (block b -> lane b % 32); block b works on
//    text column j at step tau = j + b, so its horizontal carry comes from the previous lane's previous step
//    (lane 31 hands over to lane 0), one shuffle per slot and step.
//  * Only the blocks that intersect the Ukkonen band are computed.  An edit path of cost <= k between strings
//    whose lengths differ by D = n - m stays on the diagonals [min(0,D) - x, max(0,D) + x], x = (k - |D|) / 2,
//    so the band is only k + 1 diagonals wide.  A block entering the band starts from vertical deltas of +1
//    below the score of the block above it, a block whose upper neighbour has left the band gets a horizontal
//    carry of +1: both are costs of real edit paths, so the result is never below the true distance and equals
//    it whenever the true distance is <= k.  Result > k therefore means "distance > k" (the value is then only
//    an upper bound); k < 0 or k >= max(m, n) is the plain full computation.
//  * A lane time-multiplexes G register slots: blocks b and b + 32*G never overlap in time because
//    65 * 32 * G > band width + 63 (the host picks G that way), so the Pv/Mv words of a lane live in registers
//    and their updates within a step are independent (instruction-level parallelism G).
//  * Two-level band: a job that needs G > 1 slots first runs with the widest band ONE slot per lane can hold
//    (2016 diagonals); a result inside that band is already exact (typical reads diverge far less than the
//    filter's threshold), otherwise the full band is run.
//  * The text flows through a shared-memory ring refilled with one coalesced load every 32 steps; the Peq
//    table is built with warp ballots from coalesced pattern loads.
#define VM_ED_SPAN 2080   // 65 * 32: time between the starts of blocks b and b + 32

// one banded pass: rows i with j - ku <= i <= j + kd of every column j; GE = slots per lane in use (1 or G)
template <int G>
__device__ __forceinline__ int vm_ed_pass(const unsigned long long *peq, uint8_t *ring, int ring_mask, const VmSeqView &txt,
                                          int m, int n, int W, int kd, int ku, int GE, int lane)
{
    unsigned long long Pv[G], Mv[G];
    int sc[G], cb[G], lo[G], hi[G];
    unsigned out[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
        Pv[g] = ~0ULL; Mv[g] = 0ULL; sc[g] = 0; out[g] = 0u;
        cb[g] = g < GE ? lane + 32 * g : W;
        lo[g] = 64 * cb[g] - kd > 0 ? 64 * cb[g] - kd : 0;
        hi[g] = 64 * cb[g] + 63 + ku < n - 1 ? 64 * cb[g] + 63 + ku : n - 1;
    }
    int result = -1;
    const int nsteps = n + W - 1;
    for (int tau = 0; tau < nsteps; ++tau) {
        if ((tau & 31) == 0) {
            __syncwarp();
            const int col = tau + lane;
            if (col < n) ring[col & ring_mask] = (uint8_t)vm_at(txt, col);
            __syncwarp();
        }
        unsigned rot[G];
#pragma unroll
        for (int g = 0; g < G; ++g) rot[g] = (g == 0 || GE > 1) ? __shfl_sync(VM_FULL, out[g], (lane + 31) & 31) : 0u;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const unsigned inp = lane != 0 ? rot[g] : (GE == 1 ? rot[0] : rot[(g + G - 1) % G]);
            int b = cb[g];
            if (b < W && tau - b > hi[g]) {          // this block has left the band: the slot moves on
                b += 32 * GE;
                cb[g] = b;
                lo[g] = 64 * b - kd > 0 ? 64 * b - kd : 0;
                hi[g] = 64 * b + 63 + ku < n - 1 ? 64 * b + 63 + ku : n - 1;
            }
            const int j = tau - b;
            unsigned o = 0u;
            if (b < W && j >= lo[g] && j <= hi[g]) {
                const int in_ho = (int)((inp >> 1) & 3u) - 1;
                unsigned long long pv = Pv[g], mv = Mv[g];
                int s = sc[g];
                if (j == lo[g]) {                    // entering the band (or column 0)
                    pv = ~0ULL;
                    mv = 0ULL;
                    s = lo[g] == 0 ? 64 * (b + 1) : (int)(inp >> 3) - in_ho + 64;
                }
                const int hin = (b >= 1 && j <= 64 * b - 1 + ku) ? in_ho : 1;
                unsigned long long Eq = peq[(size_t)ring[j & ring_mask] * W + b];
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
                pv = Mh | ~(Xv | Ph);
                mv = Ph & Xv;
                s += ho;
                Pv[g] = pv; Mv[g] = mv; sc[g] = s;
                o = 1u | (unsigned)(ho + 1) << 1 | (unsigned)s << 3;
                if (b == W - 1 && j == n - 1) {
                    // rows >= m of the last block are padding: take their vertical deltas back out
                    const int first_pad = m - 64 * (W - 1);
                    const unsigned long long padmask = first_pad >= 64 ? 0ULL : ~0ULL << first_pad;
                    result = s - __popcll(pv & padmask) + __popcll(mv & padmask);
                }
            }
            out[g] = o;
        }
    }
    __syncwarp();
    return __reduce_max_sync(VM_FULL, result);
}

template <int G>
__global__ void __launch_bounds__(32) vm_edit_distance_kernel(VmAlnJobDev *jobs, const int *__restrict__ job_ids, VmSeqSources S,
                                                              int max_words, int ring_mask)
{
    extern __shared__ unsigned long long vm_peq[];   // [5][W] of this job, then the text ring
    VmAlnJobDev &J = jobs[job_ids[blockIdx.x]];
    const int lane = threadIdx.x;
    const VmSeqView pat = vm_view(S, J.q, J.read), txt = vm_view(S, J.t, J.read);
    const int m = pat.len, n = txt.len;
    if (m == 0 || n == 0) { if (lane == 0) J.result0 = m + n; return; }
    const int longest = m > n ? m : n;
    const int k = (J.out_off < 0 || J.out_off >= longest) ? longest : (int)J.out_off;
    const int D = n - m, aD = D < 0 ? -D : D;
    if (aD > k) { if (lane == 0) J.result0 = (long long)k + 1; return; }
    const int W = (m + 63) >> 6;
    unsigned *peq32 = reinterpret_cast<unsigned *>(vm_peq);
    uint8_t *ring = reinterpret_cast<uint8_t *>(vm_peq + 5 * (size_t)max_words);
    for (int i0 = 0; i0 < 64 * W; i0 += 32) {
        const int i = i0 + lane;
        const int c = i < m ? vm_at(pat, i) : 5;
        const unsigned e0 = __ballot_sync(VM_FULL, c == 0), e1 = __ballot_sync(VM_FULL, c == 1),
                       e2 = __ballot_sync(VM_FULL, c == 2), e3 = __ballot_sync(VM_FULL, c == 3),
                       e4 = __ballot_sync(VM_FULL, c == 4);
        if (lane < 5) peq32[(size_t)lane * 2 * W + (i0 >> 5)] = lane == 0 ? e0 : lane == 1 ? e1 : lane == 2 ? e2 : lane == 3 ? e3 : e4;
    }
    const int up = D > 0 ? D : 0, dn = D < 0 ? -D : 0;
    if (G > 1 && W > 32 && aD <= VM_ED_SPAN - 65) {
        const int x1 = (VM_ED_SPAN - 65 - aD) / 2, k1 = aD + 2 * x1;
        if (k1 < k) {
            const int r = vm_ed_pass<G>(vm_peq, ring, ring_mask, txt, m, n, W, x1 + dn + (k1 == 0), x1 + up, 1, lane);
            if (r <= k1) { if (lane == 0) J.result0 = r; return; }
        }
    }
    const int x = (k - aD) / 2;
    const int r = vm_ed_pass<G>(vm_peq, ring, ring_mask, txt, m, n, W, x + dn + (aD + 2 * x == 0), x + up, G, lane);
    if (lane == 0) J.result0 = r;
}

template <int G>
static void vm_ed_launch_class(VmAlnJobDev *jobs, const int *ids, int n_ids, VmSeqSources src, int max_words, cudaStream_t stream)
{
    if (n_ids <= 0) return;
    int ring = 64;
    while (ring < max_words + 64) ring <<= 1;
    const size_t smem = (size_t)max_words * 5 * 8 + (size_t)ring;
    vm_smem_optin(vm_edit_distance_kernel<G>);
    vm_edit_distance_kernel<G><<<n_ids, 32, smem, stream>>>(jobs, ids, src, max_words, ring - 1);
}

// register slots per lane a job needs: its band must fit (65*32*G > band width + 63) unless every block has its own slot
int vm_ed_slots(int m, int n, long long band)
{
    const int longest = m > n ? m : n;
    const long long k = (band < 0 || band >= longest) ? longest : band;
    const int W = (m + 63) / 64;
    const int g_all = (W + 31) / 32;
    const long long d = n > m ? n - m : m - n, x = k > d ? (k - d) / 2 : 0;
    const int g_band = (int)((d + 2 * x + 1 + 63) / VM_ED_SPAN + 1);
    const int g = g_all < g_band ? g_all : g_band;
    return g < 1 ? 1 : g;
}

const int VM_ED_CLASS_G[VM_ED_NCLASS] = {1, 2, 3, 4, 6, 8, 12, 16, 32, 64};

// ids_dev: job indices grouped by class; class_start[0..VM_ED_NCLASS]; class_words[c]: largest pattern word
// count in class c (sizes the Peq table and the text ring in shared memory)
int vm_launch_edit_distance(VmAlnJobDev *jobs, const int *ids_dev, const int *class_start, const int *class_words, VmSeqSources src,
                            cudaStream_t stream)
{
    int launches = 0;
    for (int c = 0; c < VM_ED_NCLASS; ++c) {
        const int n_ids = class_start[c + 1] - class_start[c];
        if (n_ids <= 0) continue;
        const int *ids = ids_dev + class_start[c];
        const int mw = class_words[c] > 0 ? class_words[c] : 1;
        switch (VM_ED_CLASS_G[c]) {
        case 1: vm_ed_launch_class<1>(jobs, ids, n_ids, src, mw, stream); break;
        case 2: vm_ed_launch_class<2>(jobs, ids, n_ids, src, mw, stream); break;
        case 3: vm_ed_launch_class<3>(jobs, ids, n_ids, src, mw, stream); break;
        case 4: vm_ed_launch_class<4>(jobs, ids, n_ids, src, mw, stream); break;
        case 6: vm_ed_launch_class<6>(jobs, ids, n_ids, src, mw, stream); break;
        case 8: vm_ed_launch_class<8>(jobs, ids, n_ids, src, mw, stream); break;
        case 12: vm_ed_launch_class<12>(jobs, ids, n_ids, src, mw, stream); break;
        case 16: vm_ed_launch_class<16>(jobs, ids, n_ids, src, mw, stream); break;
        case 32: vm_ed_launch_class<32>(jobs, ids, n_ids, src, mw, stream); break;
        default: vm_ed_launch_class<64>(jobs, ids, n_ids, src, mw, stream); break;
        }
        ++launches;
    }
    return launches;
}

// ---------------------------------------------------------------------------
// edit distance, upper bound through the chain's exact-match segments
// ---------------------------------------------------------------------------
// The divergence filter (mammap_clrnano.py:19251-19254) only asks whether distance / min(len) exceeds a
// threshold.  Any alignment's cost bounds the distance from above, and the sub-alignment's own anchors spell
// one out: walk the match segments (mismatches on them are counted, so nothing is assumed about the
// anchors) and align each gap between consecutive segments optimally (unit-cost NW, Myers bit-vector with
// the <= 128-base query gap in two registers).  If that cost is already within the threshold the exact
// distance is not needed; the few jobs it does not settle go to vm_edit_distance_kernel.
// One warp per job, one lane per segment (+ the gap in front of it); J.dir_off / J.n_out = offset / number
// of the job's segments, J.result0 = the bound.
struct VmMatchSeg { int32_t q, t, l; };

__device__ __forceinline__ int vm_gap_distance(const VmSeqView &Q, int q0, int m, const VmSeqView &T, int t0, int n)
{
    if (m <= 0) return n;
    if (n <= 0) return m;
    unsigned long long P[5][2];
#pragma unroll
    for (int c = 0; c < 5; ++c) { P[c][0] = 0ULL; P[c][1] = 0ULL; }
    for (int i = 0; i < m; ++i) {
        const int c = vm_at(Q, q0 + i);
        const unsigned long long lo = i < 64 ? 1ULL << i : 0ULL, hi = i >= 64 ? 1ULL << (i - 64) : 0ULL;
#pragma unroll
        for (int cc = 0; cc < 5; ++cc)
            if (c == cc) { P[cc][0] |= lo; P[cc][1] |= hi; }
    }
    const bool two = m > 64;
    unsigned long long Pv0 = ~0ULL, Mv0 = 0ULL, Pv1 = ~0ULL, Mv1 = 0ULL;
    const unsigned long long last = 1ULL << ((m - 1) & 63);
    int score = m;
    for (int j = 0; j < n; ++j) {
        const int c = vm_at(T, t0 + j);
        unsigned long long E0 = 0ULL, E1 = 0ULL;
#pragma unroll
        for (int cc = 0; cc < 5; ++cc)
            if (c == cc) { E0 = P[cc][0]; E1 = P[cc][1]; }
        // block 0, horizontal input +1 (global alignment: D[0][j] = j)
        unsigned long long Xv = E0 | Mv0;
        unsigned long long Xh = (((E0 & Pv0) + Pv0) ^ Pv0) | E0;
        unsigned long long Ph = Mv0 | ~(Xh | Pv0);
        unsigned long long Mh = Pv0 & Xh;
        if (!two) score += (Ph & last) ? 1 : ((Mh & last) ? -1 : 0);
        const int ho = (int)(Ph >> 63) - (int)(Mh >> 63);
        Ph = Ph << 1 | 1ULL;
        Mh <<= 1;
        Pv0 = Mh | ~(Xv | Ph);
        Mv0 = Ph & Xv;
        if (two) {
            const unsigned long long neg = ho < 0 ? 1ULL : 0ULL;
            Xv = E1 | Mv1;
            E1 |= neg;
            Xh = (((E1 & Pv1) + Pv1) ^ Pv1) | E1;
            Ph = Mv1 | ~(Xh | Pv1);
            Mh = Pv1 & Xh;
            score += (Ph & last) ? 1 : ((Mh & last) ? -1 : 0);
            Ph <<= 1;
            Mh <<= 1;
            Mh |= neg;
            Ph |= ho > 0 ? 1ULL : 0ULL;
            Pv1 = Mh | ~(Xv | Ph);
            Mv1 = Ph & Xv;
        }
    }
    return score;
}

// Match segments of a job from the anchors of its sub-alignment (device copy of vmg::match_segments, vm_glue.hpp):
// the anchors of job j are anc[J.dir_off .. + J.n_out) in ascending read order; the segments replace them in
// segs[J.dir_off ..) and J.n_out becomes their number (0: no bound possible -- a query gap beyond the 128 bases the
// gap DP holds in registers, or mixed strands -- the exact kernel will take the job).
__global__ void __launch_bounds__(128) vm_match_segments_kernel(VmAlnJobDev *jobs, int n_jobs, const VmAnchor *__restrict__ anc,
                                                                VmMatchSeg *segs, int max_qgap)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_jobs) return;
    VmAlnJobDev &J = jobs[t];
    const int n = J.n_out;
    const VmAnchor *A = anc + J.dir_off;
    VmMatchSeg *out = segs + J.dir_off;
    const long long qlen = J.q.len, tlen = J.t.len;
    int m = 0;
    bool ok = n >= 2;
    if (ok) {
        const VmAnchor pre = A[0], now = A[n - 1];
        const bool fwd = pre.s == 1;
        long long cq = 0, ct = 0;
        for (int k = 0; k < n && ok; ++k) {
            const VmAnchor a = fwd ? A[k] : A[n - 1 - k];
            if (a.s != pre.s) { ok = false; break; }
            long long q = fwd ? (long long)a.x - pre.x : (long long)now.x - a.x - a.l;
            long long tt = fwd ? (long long)a.y - (long long)pre.y : (long long)a.y - ((long long)now.y + now.l);
            long long l = a.l;
            long long d = cq - q > ct - tt ? cq - q : ct - tt;
            if (d < 0) d = 0;
            q += d; tt += d; l -= d;
            if (l > qlen - q) l = qlen - q;
            if (l > tlen - tt) l = tlen - tt;
            if (l <= 0) continue;
            if (q - cq > max_qgap) { ok = false; break; }
            VmMatchSeg sg;
            sg.q = (int32_t)q; sg.t = (int32_t)tt; sg.l = (int32_t)l;
            out[m++] = sg;
            cq = q + l;
            ct = tt + l;
        }
        if (ok && qlen - cq > max_qgap) ok = false;
    }
    J.n_out = ok ? m : -1;
}

int vm_launch_match_segments(VmAlnJobDev *jobs, int n_jobs, const VmAnchor *anc_dev, void *segs_dev, cudaStream_t stream)
{
    if (n_jobs <= 0) return 0;
    vm_match_segments_kernel<<<(n_jobs + 127) / 128, 128, 0, stream>>>(jobs, n_jobs, anc_dev, (VmMatchSeg *)segs_dev, 128);
    return 1;
}

__global__ void __launch_bounds__(128) vm_ed_upper_kernel(VmAlnJobDev *jobs, const int *__restrict__ job_ids, int n_jobs,
                                                          const VmMatchSeg *__restrict__ segs, VmSeqSources S)
{
    const int w = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (w >= n_jobs) return;
    VmAlnJobDev &J = jobs[job_ids ? job_ids[w] : w];
    const VmSeqView Q = vm_view(S, J.q, J.read), T = vm_view(S, J.t, J.read);
    const VmMatchSeg *sg = segs + J.dir_off;
    const int n = J.n_out;
    if (n < 0) { if (lane == 0) J.result0 = 1LL << 40; return; }      // no segments: no bound
    long long cost = 0;
    for (int i = lane; i <= n; i += 32) {
        int cq = 0, ct = 0;
        if (i > 0) { const VmMatchSeg p = sg[i - 1]; cq = p.q + p.l; ct = p.t + p.l; }
        int nq = Q.len, nt = T.len, l = 0;
        if (i < n) { const VmMatchSeg a = sg[i]; nq = a.q; nt = a.t; l = a.l; }
        cost += vm_gap_distance(Q, cq, nq - cq, T, ct, nt - ct);
        for (int x = 0; x < l; ++x) cost += vm_at(Q, nq + x) != vm_at(T, nt + x);
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) cost += __shfl_xor_sync(VM_FULL, cost, d);
    if (lane == 0) J.result0 = cost;
}

int vm_launch_ed_upper(VmAlnJobDev *jobs, const int *ids_dev, int n_jobs, const void *segs_dev, VmSeqSources src, cudaStream_t stream)
{
    if (n_jobs <= 0) return 0;
    vm_ed_upper_kernel<<<(n_jobs + 3) / 4, 128, 0, stream>>>(jobs, ids_dev, n_jobs, (const VmMatchSeg *)segs_dev, src);
    return 1;
}

#include "vm_extend.cuh"

__global__ void __launch_bounds__(32) vm_extend_kernel(VmAlnJobDev *jobs, VmSeqSources S)
{
    __shared__ VmExtSmem M;
    VmAlnJobDev &J = jobs[blockIdx.x];
    const VmSeqView T = vm_view(S, J.t, J.read), Q = vm_view(S, J.q, J.read);
    int q_e, t_e;
    vm_extend_warp(T, Q, M, q_e, t_e);
    if (threadIdx.x == 0) { J.result0 = q_e; J.result1 = t_e; }
}

int vm_launch_extend(VmAlnJobDev *jobs, int n_jobs, VmSeqSources src, cudaStream_t stream)
{
    if (n_jobs <= 0) return 0;
    vm_extend_kernel<<<n_jobs, 32, 0, stream>>>(jobs, src);
    return 1;
}

