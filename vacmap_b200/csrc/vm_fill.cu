// Global dual-affine fill with traceback: replaces mp.k_cigar(target, query, 2, -4, 4, 2, 24, 1, bw=-1,
// zdropvalue=-1, eqx) at mammap_clrnano.py:21554, 21598.  Recurrences, tie order (diag > E1 > F1 > E2 > F2),
// left-aligned gaps and boundary conditions are those of oracle/orc_align.c (ksw2 extd2 semantics).
//
// Design (sm_100a, no tensor cores -- integer-valued DP, nothing here is a dense contraction):
//  * Difference recurrences.  The kernel carries ksw2's u/v/x/y quantities (differences between neighbouring
//    cells, all within +-64 for these penalties) instead of absolute scores.  Every comparison of the absolute
//    recurrence is a comparison of the same two numbers shifted by H(i-1,j-1), so the direction bits are
//    identical, and the small range makes the arithmetic EXACT in fp16.  Two jobs share a warp: job A lives in
//    the low half and job B in the high half of every half2 register, one HADD2 / HMNMX2 / HSET2 serves both
//    (the packed-integer video instructions are emulated on this architecture, half2 is native and splits
//    between the FMA pipe (HADD2, HFMA2.RELU) and the ALU pipe (HSET2, HMNMX2, LOP3)).
//  * Strip-per-lane pipelined wavefront.  Lane l owns R consecutive target rows (R = 2..16 by capacity
//    class, so a 270-row job runs with R = 10 and no idle rows); at step s it sweeps column s - l of its strip
//    in registers and hands v/x1/x2 of its bottom row to lane l+1 with three shuffles.
//  * Direction bytes (3 source bits, 4 gap-continuation bits, 1 match bit for eqx) go to a per-warp scratch
//    region laid out [step][word][lane] -- every store instruction writes one full 128-byte line -- that is
//    reused by the next pair of the persistent warp, so the traceback reads mostly hit L2.
//  * Traceback: two walker lanes (one per job) run on 32-step x R-row tiles the whole warp stages into shared
//    memory with one round trip, instead of one dependent global load per path cell.
#include "vm_fill_cell.cuh"
#include "vm_hostpool.hpp"
#include <algorithm>
#include <cstdlib>
#include <cuda_fp16.h>

namespace {

template <int R, bool MB>
__global__ void __launch_bounds__(128) vm_fill_kernel(VmAlnJobDev *jobs, const VmFillPair *__restrict__ pairs, int pair_begin,
                                                      int pair_end, VmSeqSources S, int eqx, uint32_t *dir_all,
                                                      long long dir_words_per_warp, uint32_t *band_all,
                                                      long long band_words_per_warp, int *counter, uint32_t *cigar_out,
                                                      uint32_t *dense_out, unsigned long long *dense_count, uint2 *results)
{
    static_assert(R % 2 == 0 && R >= 2 && R <= 16, "rows per lane");
    constexpr VmGapPar2 g = vm_fill_par();
    static_assert(g.match == 2, "score constants are folded into vm_cell2");
    constexpr int RW = R / 2, ROWS = 32 * R;
    __shared__ uint32_t tile[4][2][32][RW];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long gw = (long long)blockIdx.x * 4 + warp;
    uint32_t *dir = dir_all + gw * dir_words_per_warp;
    uint32_t *bs = MB ? band_all + gw * band_words_per_warp : nullptr;
    const __half2 nan2 = vm_h2(0x7fff7fffu);
    for (;;) {
        int p = 0;
        if (lane == 0) p = pair_begin + atomicAdd(counter, 1);
        p = __shfl_sync(VM_FULL, p, 0);
        if (p >= pair_end) break;
        const VmFillPair pr = pairs[p];
        const bool hasB = pr.b >= 0;
        VmAlnJobDev &JA = jobs[pr.a];
        VmAlnJobDev &JB = jobs[hasB ? pr.b : pr.a];
        const VmSeqView TA = vm_view(S, JA.t, JA.read), QA = vm_view(S, JA.q, JA.read);
        const VmSeqView TB = vm_view(S, JB.t, JB.read), QB = vm_view(S, JB.q, JB.read);
        const int tA = TA.len, qA = QA.len, tB = hasB ? TB.len : 0, qB = hasB ? QB.len : 0;
        const int tlen = tA > tB ? tA : tB, qlen = qA > qB ? qA : qB;
        const int nbands = MB ? (tlen + ROWS - 1) / ROWS : 1;
        unsigned tN = 0;     // bit 0 / 1: the target of job A / B holds an N (eqx needs the sequences there)
        // ---------------- forward pass ----------------
        for (int band = 0; band < nbands; ++band) {
            const int row0 = band * ROWS + lane * R;
            __half2 tc[R], u[R], y1[R], y2[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int t = row0 + r;
                const int ca = t < tA ? vm_at(TA, t) : 4, cb = t < tB ? vm_at(TB, t) : 4;
                if (t < tA && ca == 4) tN |= 1u;
                if (t < tB && cb == 4) tN |= 2u;
                tc[r] = vm_h2(vm_code_half(ca) | vm_code_half(cb) << 16);
                u[r] = vm_h2i(vm_hb(t + 1) - vm_hb(t));           // H(t,-1) - H(t-1,-1)
                y1[r] = VM_H2C(-(g.q1 + g.e1));
                y2[r] = VM_H2C(-(g.q2 + g.e2));
            }
            const int rows_in_band = (tlen - band * ROWS) < ROWS ? (tlen - band * ROWS) : ROWS;
            const int last_lane = (rows_in_band - 1) / R;
            const int nsteps = qlen + last_lane;
            uint32_t *dband = dir + (long long)band * (qlen + 32) * (RW * 32) + lane;
            __half2 outV = nan2, outX1 = nan2, outX2 = nan2, qc_pipe = nan2, qbuf = nan2;
            __half2 hv = nan2, hx1 = nan2, hx2 = nan2;
            for (int s = 0; s < nsteps; ++s) {
                if ((s & 31) == 0) {          // stage the next 32 query columns (and band hand-over values)
                    const int col = s + lane;
                    const int ca = col < qA ? vm_at(QA, col) : 4, cb = col < qB ? vm_at(QB, col) : 4;
                    qbuf = vm_h2(vm_code_half(ca) | vm_code_half(cb) << 16);
                    if (MB && band > 0 && col < qlen) {
                        hv = vm_h2(bs[col]);
                        hx1 = vm_h2(bs[qlen + col]);
                        hx2 = vm_h2(bs[2 * qlen + col]);
                    }
                }
                __half2 inV = __shfl_up_sync(VM_FULL, outV, 1);
                __half2 inX1 = __shfl_up_sync(VM_FULL, outX1, 1);
                __half2 inX2 = __shfl_up_sync(VM_FULL, outX2, 1);
                __half2 qc = __shfl_up_sync(VM_FULL, qc_pipe, 1);
                const __half2 q0 = __shfl_sync(VM_FULL, qbuf, s & 31);
                __half2 v0 = nan2, x10 = nan2, x20 = nan2;
                if (MB) {
                    v0 = __shfl_sync(VM_FULL, hv, s & 31);
                    x10 = __shfl_sync(VM_FULL, hx1, s & 31);
                    x20 = __shfl_sync(VM_FULL, hx2, s & 31);
                }
                if (lane == 0) {
                    qc = q0;
                    if (!MB || band == 0) {
                        inV = vm_h2i(vm_hb(s + 1) - vm_hb(s));      // H(-1,q) - H(-1,q-1)
                        inX1 = VM_H2C(-(g.q1 + g.e1));
                        inX2 = VM_H2C(-(g.q2 + g.e2));
                    } else { inV = v0; inX1 = x10; inX2 = x20; }
                }
                qc_pipe = qc;
                const int q = s - lane;
                if (q >= 0 && q < qlen && lane <= last_lane) {
                    __half2 v = inV, x1 = inX1, x2 = inX2;
                    uint32_t *dst = dband + (long long)s * (RW * 32);
                    unsigned dprev = 0;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        unsigned d;
                        vm_cell2(tc[r], qc, v, x1, x2, u[r], y1[r], y2[r], d);
                        if (r & 1) dst[(r >> 1) * 32] = __byte_perm(dprev, d, 0x6420);   // [A even, B even, A odd, B odd]
                        else dprev = d;
                    }
                    outV = v; outX1 = x1; outX2 = x2;
                    if (MB && band + 1 < nbands && lane == 31) {
                        bs[q] = vm_u32(v);
                        bs[qlen + q] = vm_u32(x1);
                        bs[2 * qlen + q] = vm_u32(x2);
                    }
                }
            }
            __syncwarp();
        }
        // ---------------- traceback (ksw_backtrack, left-aligned): lane 0 walks job A, lane 1 job B ----------------
        tN = __reduce_or_sync(VM_FULL, tN);
        const int w = lane & 1;
        const VmSeqView &Tw = w ? TB : TA, &Qw = w ? QB : QA;
        const int tw = w ? tB : tA, qw = w ? qB : qA;
        const bool walker = lane < 2 && tw > 0 && qw > 0;
        uint32_t *out = cigar_out + (w ? JB.out_off : JA.out_off);
        int i = tw - 1, j = qw - 1, state = 0, n = 0;
        unsigned cur_op = 0, cur_len = 0;
        const bool checkN = eqx && ((tN >> w) & 1u);
        for (;;) {
            const bool need = walker && i >= 0 && j >= 0;
            const unsigned needmask = __ballot_sync(VM_FULL, need) & 3u;
            if (!needmask) break;
#pragma unroll
            for (int ws = 0; ws < 2; ++ws) {
                const int ii = __shfl_sync(VM_FULL, i, ws), jj = __shfl_sync(VM_FULL, j, ws);
                if (needmask >> ws & 1u) {
                    const int band_t = ii / ROWS, lane_t = (ii % ROWS) / R, step = jj + lane_t - lane;
                    if (step >= 0) {
                        const uint32_t *src = dir + ((long long)band_t * (qlen + 32) + step) * (RW * 32) + lane_t;
#pragma unroll
                        for (int m = 0; m < RW; ++m) tile[warp][ws][lane][m] = src[m * 32];
                    }
                }
            }
            __syncwarp();
            if (need) {
                const int band_t = i / ROWS, lane_t = (i % ROWS) / R, s_hi = j + lane_t;
                while (i >= 0 && j >= 0) {
                    const int tb = i % ROWS, lc = tb / R, r = tb % R, s = j + lc;
                    if (i / ROWS != band_t || lc != lane_t || s <= s_hi - 32) break;
                    const unsigned tmp = (tile[warp][w][s_hi - s][r >> 1] >> (((r & 1) * 2 + w) * 8)) & 0xffu;
                    if (state == 0) state = tmp & 7;
                    else if (!((tmp >> (state + 2)) & 1)) state = 0;
                    if (state == 0) state = tmp & 7;
                    unsigned op;
                    if (state == 0) {
                        op = 0;
                        if (eqx) {
                            op = (tmp & 0x80u) ? 7u : 8u;
                            if (checkN && op == 8u && vm_at(Tw, i) == vm_at(Qw, j)) op = 7u;
                        }
                        --i; --j;
                    } else if (state == 1 || state == 3) { op = 2; --i; }
                    else { op = 1; --j; }
                    if (op == cur_op) ++cur_len;
                    else {
                        if (cur_len) out[n++] = cur_len << 4 | cur_op;
                        cur_op = op;
                        cur_len = 1;
                    }
                }
            }
            __syncwarp();
        }
        if (walker) {
            if (i >= 0) {
                if (cur_len && cur_op == 2u) cur_len += (unsigned)(i + 1);
                else {
                    if (cur_len) out[n++] = cur_len << 4 | cur_op;
                    cur_op = 2u;
                    cur_len = (unsigned)(i + 1);
                }
            }
            if (j >= 0) {
                if (cur_len && cur_op == 1u) cur_len += (unsigned)(j + 1);
                else {
                    if (cur_len) out[n++] = cur_len << 4 | cur_op;
                    cur_op = 1u;
                    cur_len = (unsigned)(j + 1);
                }
            }
            if (cur_len) out[n++] = cur_len << 4 | cur_op;
        }
        __syncwarp();
        // ops were pushed end to start: claim room in the dense CIGAR arena and copy them over flipped, all lanes helping
#pragma unroll
        for (int ws = 0; ws < 2; ++ws) {
            if (ws == 1 && !hasB) break;
            const int nn = __shfl_sync(VM_FULL, walker ? n : 0, ws);
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(dense_count, (unsigned long long)nn);
            base = __shfl_sync(VM_FULL, base, 0);
            const uint32_t *o = cigar_out + (ws ? JB.out_off : JA.out_off);
            for (int x = lane; x < nn; x += 32) dense_out[base + x] = o[nn - 1 - x];
            if (lane == 0) results[ws ? pr.b : pr.a] = make_uint2((unsigned)base, (unsigned)nn);
        }
        __syncwarp();
    }
}

template <int R, bool MB>
int vm_fill_occupancy()
{
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, vm_fill_kernel<R, MB>, 128, 0) != cudaSuccess || nb < 1) nb = 1;
    return nb;
}

int vm_fill_blocks_per_sm(int R, bool mb)
{
    if (mb) return vm_fill_occupancy<16, true>();
    switch (R) {
    case 2: return vm_fill_occupancy<2, false>();
    case 4: return vm_fill_occupancy<4, false>();
    case 6: return vm_fill_occupancy<6, false>();
    case 8: return vm_fill_occupancy<8, false>();
    case 10: return vm_fill_occupancy<10, false>();
    case 12: return vm_fill_occupancy<12, false>();
    case 14: return vm_fill_occupancy<14, false>();
    default: return vm_fill_occupancy<16, false>();
    }
}

} // namespace

// Capacity classes: rows class rc = ceil(tlen / 64) for tlen <= 512 (R = 2 * rc rows per lane, one band),
// rc = 9 beyond (R = 16, several bands); column class by qlen (<= 512, <= 4096, longer) so that one very long
// query does not size the scratch of every resident warp.  Inside a class jobs are ordered by qlen and paired
// with their neighbour: both halves of the half2 registers do useful work for (nearly) the whole sweep.
void vm_fill_plan(const VmAlnJobDev *J, int nj, int sm_count, VmFillPlan &plan, int host_threads, const uint8_t *only_mask)
{
    plan.pairs.clear();
    plan.launches.clear();
    plan.dir_words = plan.band_words = 0;
    plan.dir_bytes = 0;
    constexpr int NCLS = 27, NB = 1024, NKEY = NCLS * NB;
    auto key_of = [&](int j) {
        const int tl = J[j].t.len, ql = J[j].q.len;
        if (tl <= 0 || ql <= 0 || (only_mask && !only_mask[j])) return -1;
        const int rc = tl <= 512 ? (tl + 63) / 64 : 9;
        const int qc = ql <= 512 ? 0 : ql <= 4096 ? 1 : 2;
        const int cls = (rc - 1) * 3 + qc;
        const int b = qc == 0 ? ql >> 1 : qc == 1 ? ql >> 3 : std::min(ql >> 8, NB - 1);
        return cls * NB + b;
    };
    // stable counting sort of the jobs by key: per-slice histograms, then one pass of offsets
    const int T = std::max(1, std::min(host_threads, nj / 8192 + 1));
    std::vector<int32_t> keys((size_t)nj);
    std::vector<std::vector<int32_t>> hist((size_t)T, std::vector<int32_t>((size_t)NKEY, 0));
    auto slice = [&](int t, int &lo, int &hi) { lo = (int)((long long)nj * t / T); hi = (int)((long long)nj * (t + 1) / T); };
    vmp::parallel_for(T, T, [&](int64_t t) {
        int lo, hi;
        slice((int)t, lo, hi);
        std::vector<int32_t> &h = hist[(size_t)t];
        for (int j = lo; j < hi; ++j) {
            keys[j] = key_of(j);
            if (keys[j] >= 0) ++h[(size_t)keys[j]];
        }
    }, 1);
    std::vector<int32_t> count((size_t)NKEY + 1, 0);
    {
        int32_t run = 0;
        for (int k = 0; k < NKEY; ++k) {
            count[(size_t)k] = run;
            for (int t = 0; t < T; ++t) {
                const int32_t c = hist[(size_t)t][(size_t)k];
                hist[(size_t)t][(size_t)k] = run;     // where slice t starts writing key k
                run += c;
            }
        }
        count[(size_t)NKEY] = run;
    }
    const int n_live = count[(size_t)NKEY];
    std::vector<int32_t> order((size_t)n_live);
    vmp::parallel_for(T, T, [&](int64_t t) {
        int lo, hi;
        slice((int)t, lo, hi);
        std::vector<int32_t> &pos = hist[(size_t)t];
        for (int j = lo; j < hi; ++j)
            if (keys[j] >= 0) order[(size_t)pos[(size_t)keys[j]]++] = j;
    }, 1);
    const size_t mem_cap_words = (size_t)6 << 28;     // 6 GiB of direction scratch at most
    for (int cls = 0; cls < NCLS; ++cls) {
        const int lo = count[(size_t)cls * NB], hi = count[(size_t)(cls + 1) * NB];
        if (hi <= lo) continue;
        const int rc = cls / 3 + 1;
        VmFillLaunch L;
        L.multiband = rc == 9;
        L.R = L.multiband ? 16 : 2 * rc;
        L.pair_begin = (int)plan.pairs.size();
        int max_q = 0, max_t = 0;
        for (int x = lo; x < hi; x += 2) {
            VmFillPair pr;
            pr.a = order[x];
            pr.b = x + 1 < hi ? order[x + 1] : -1;
            plan.pairs.push_back(pr);
        }
        for (int x = lo; x < hi; ++x) {
            max_q = std::max(max_q, J[order[x]].q.len);
            max_t = std::max(max_t, J[order[x]].t.len);
        }
        for (int x = lo; x < hi; x += 2) {
            const int tl = std::max(J[order[x]].t.len, x + 1 < hi ? J[order[x + 1]].t.len : 0);
            const int ql = std::max(J[order[x]].q.len, x + 1 < hi ? J[order[x + 1]].q.len : 0);
            const int rows_pb = 32 * (rc == 9 ? 16 : 2 * rc);
            plan.dir_bytes += (double)((tl + rows_pb - 1) / rows_pb) * (ql + 31) * (rows_pb / 64) * 128.0;
        }
        L.pair_end = (int)plan.pairs.size();
        const int rows = 32 * L.R;
        const long long nbands = L.multiband ? (max_t + rows - 1) / rows : 1;
        L.dir_words_per_warp = nbands * ((long long)max_q + 32) * (L.R / 2) * 32;
        L.band_words_per_warp = L.multiband ? 3LL * max_q + 32 : 0;
        const int n_pairs = L.pair_end - L.pair_begin;
        static const double frac = getenv("VM_FILL_SM_FRAC") ? atof(getenv("VM_FILL_SM_FRAC")) : 1.0;   // experiment knob
        long long blocks = std::min<long long>((n_pairs + 3) / 4,
                                               (long long)(frac * sm_count * vm_fill_blocks_per_sm(L.R, L.multiband != 0)));
        const long long fit = (long long)(mem_cap_words / (size_t)(4 * L.dir_words_per_warp));
        blocks = std::max<long long>(1, std::min(blocks, std::max<long long>(fit, 1)));
        L.blocks = (int)blocks;
        plan.dir_words += (size_t)(blocks * 4 * L.dir_words_per_warp);          // every launch has its own slice
        plan.band_words += (size_t)(blocks * 4 * L.band_words_per_warp);
        plan.launches.push_back(L);
    }
}

int vm_fill_launch(const VmFillPlan &plan, VmAlnJobDev *jobs, const VmFillPair *pairs, VmSeqSources src, int eqx, uint32_t *dir,
                   uint32_t *band, int *counters, uint32_t *cigar_out, uint32_t *dense_out, unsigned long long *dense_count,
                   void *results, cudaStream_t main_stream, const cudaStream_t *side, int n_side, int *side_rr, size_t *dir_cursor,
                   size_t *band_cursor)
{
    int n = 0;
    uint32_t *const dir0 = dir, *const band0 = band;
    for (size_t li = 0; li < plan.launches.size(); ++li) {
        const VmFillLaunch &L = plan.launches[li];
        int *ctr = counters + li;
        // see vm_fillb_launch: small launches on side streams, one scratch slice per launch
        const bool small = n_side > 0 && L.blocks < VM_FILL_SMALL_BLOCKS;
        cudaStream_t stream = small ? side[(*side_rr)++ % n_side] : main_stream;
        dir = dir0 + *dir_cursor;
        band = band0 + *band_cursor;
        *dir_cursor += (size_t)L.blocks * 4 * (size_t)L.dir_words_per_warp;
        *band_cursor += (size_t)L.blocks * 4 * (size_t)L.band_words_per_warp;
#define VM_FILL_GO(RR, MBB)                                                                                           \
    vm_fill_kernel<RR, MBB><<<L.blocks, 128, 0, stream>>>(jobs, pairs, L.pair_begin, L.pair_end, src, eqx, dir,       \
                                                           L.dir_words_per_warp, band, L.band_words_per_warp, ctr, cigar_out,      \
                                                           dense_out, dense_count, (uint2 *)results)
        if (L.multiband) VM_FILL_GO(16, true);
        else switch (L.R) {
            case 2: VM_FILL_GO(2, false); break;
            case 4: VM_FILL_GO(4, false); break;
            case 6: VM_FILL_GO(6, false); break;
            case 8: VM_FILL_GO(8, false); break;
            case 10: VM_FILL_GO(10, false); break;
            case 12: VM_FILL_GO(12, false); break;
            case 14: VM_FILL_GO(14, false); break;
            default: VM_FILL_GO(16, false); break;
            }
#undef VM_FILL_GO
        ++n;
    }
    return n;
}

// ---------------------------------------------------------------------------------------------------------------
// The same plan made on the device (vm_fill_plan above is the host version the stage-level entry points use).
// ---------------------------------------------------------------------------------------------------------------
#include "vm_devsort.cuh"

namespace {

constexpr int FF_NCLS = 27, FF_NB = 1024, FF_NKEY = FF_NCLS * FF_NB;
struct FfSmall { VmFfTable tab; };

__global__ void vm_ffp_key_kernel(const VmAlnJobDev *__restrict__ J, int nj, const uint8_t *__restrict__ only_mask, int32_t *keys, FfSmall *sm)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nj) return;
    const int tl = J[j].t.len, ql = J[j].q.len;
    int key = -1;
    if (tl > 0 && ql > 0 && (!only_mask || only_mask[j])) {
        const int rc = tl <= 512 ? (tl + 63) / 64 : 9;
        const int qc = ql <= 512 ? 0 : ql <= 4096 ? 1 : 2;
        const int cls = (rc - 1) * 3 + qc;
        const int b = qc == 0 ? ql >> 1 : qc == 1 ? ql >> 3 : min(ql >> 8, FF_NB - 1);
        key = cls * FF_NB + b;
        atomicMax(&sm->tab.max_q[cls], ql);
        atomicMax(&sm->tab.max_t[cls], tl);
    }
    keys[j] = key;
}

__global__ void vm_ffp_bounds_kernel(const int32_t *__restrict__ start, FfSmall *sm)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int pb = 0;
    for (int c = 0; c < FF_NCLS; ++c) {
        const int n = start[(c + 1) * FF_NB] - start[c * FF_NB];
        sm->tab.n[c] = n;
        sm->tab.pair_begin[c] = pb;
        pb += (n + 1) / 2;
    }
    sm->tab.n_live = start[FF_NKEY];
    sm->tab.n_pairs = pb;
}

__global__ void vm_ffp_pair_kernel(const VmAlnJobDev *__restrict__ J, int nj, const int32_t *__restrict__ keys, const int32_t *__restrict__ order,
                                   const int32_t *__restrict__ start, FfSmall *sm, VmFillPair *pairs_out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= nj || x >= start[FF_NKEY]) return;
    const int a = order[x];
    const int cls = keys[a] / FF_NB;
    const int lo = start[cls * FF_NB], hi = start[(cls + 1) * FF_NB];
    const int rel = x - lo;
    if (rel & 1) return;
    VmFillPair pr;
    pr.a = a;
    pr.b = x + 1 < hi ? order[x + 1] : -1;
    pairs_out[sm->tab.pair_begin[cls] + rel / 2] = pr;
    const int rc = cls / 3 + 1;
    const int tl = max(J[a].t.len, pr.b >= 0 ? J[pr.b].t.len : 0), ql = max(J[a].q.len, pr.b >= 0 ? J[pr.b].q.len : 0);
    const int rows_pb = 32 * (rc == 9 ? 16 : 2 * rc);
    atomicAdd(&sm->tab.dir_bytes, (double)((tl + rows_pb - 1) / rows_pb) * (ql + 31) * (rows_pb / 64) * 128.0);
}

} // namespace

int vm_fill_plan_dev(const VmAlnJobDev *J, int nj, const uint8_t *only_mask, VmFillPlanBufs &B, VmFillPair *pairs_out, cudaStream_t stream)
{
    if (nj <= 0) return 0;
    const size_t n = (size_t)nj;
    if (B.keys.ensure(n * 4) || B.order.ensure(n * 4) || B.start.ensure(((size_t)FF_NKEY + 1) * 4) || B.cursor.ensure((size_t)FF_NKEY * 4) ||
        B.small.ensure(sizeof(FfSmall) + 1024) || B.table.ensure(sizeof(VmFbTable) + sizeof(VmFfTable) + 64))
        return -1;
    FfSmall *sm = B.small.as<FfSmall>();
    const int nb = (nj + 127) / 128;
    cudaMemsetAsync(sm, 0, sizeof(FfSmall), stream);
    vm_ffp_key_kernel<<<nb, 128, 0, stream>>>(J, nj, only_mask, B.keys.as<int32_t>(), sm);
    int launches = 1 + vm_bucket_sort(B.keys.as<int32_t>(), nj, FF_NKEY, B.start.as<int32_t>(), B.cursor.as<int32_t>(), B.order.as<int32_t>(), stream);
    vm_ffp_bounds_kernel<<<1, 32, 0, stream>>>(B.start.as<int32_t>(), sm);
    vm_ffp_pair_kernel<<<nb, 128, 0, stream>>>(J, nj, B.keys.as<int32_t>(), B.order.as<int32_t>(), B.start.as<int32_t>(), sm, pairs_out);
    launches += 2;
    if (cudaMemcpyAsync(B.table.p, &sm->tab, sizeof(VmFfTable), cudaMemcpyDeviceToHost, stream) != cudaSuccess) return -1;
    return launches;
}

void vm_fill_plan_finish(const VmFillPlanBufs &B, int sm_count, VmFillPlan &plan)
{
    plan.pairs.clear();
    plan.launches.clear();
    plan.dir_words = plan.band_words = 0;
    const VmFfTable &T = *B.table.as<VmFfTable>();
    plan.dir_bytes = T.dir_bytes;
    const size_t mem_cap_words = (size_t)6 << 28;
    for (int cls = 0; cls < FF_NCLS; ++cls) {
        if (T.n[cls] <= 0) continue;
        const int rc = cls / 3 + 1;
        VmFillLaunch L;
        L.multiband = rc == 9;
        L.R = L.multiband ? 16 : 2 * rc;
        L.pair_begin = T.pair_begin[cls];
        L.pair_end = T.pair_begin[cls] + (T.n[cls] + 1) / 2;
        const int rows = 32 * L.R;
        const long long nbands = L.multiband ? (T.max_t[cls] + rows - 1) / rows : 1;
        L.dir_words_per_warp = nbands * ((long long)T.max_q[cls] + 32) * (L.R / 2) * 32;
        L.band_words_per_warp = L.multiband ? 3LL * T.max_q[cls] + 32 : 0;
        const int n_pairs = L.pair_end - L.pair_begin;
        static const double frac = getenv("VM_FILL_SM_FRAC") ? atof(getenv("VM_FILL_SM_FRAC")) : 1.0;   // experiment knob
        long long blocks = std::min<long long>((n_pairs + 3) / 4, (long long)(frac * sm_count * vm_fill_blocks_per_sm(L.R, L.multiband != 0)));
        const long long fit = (long long)(mem_cap_words / (size_t)(4 * L.dir_words_per_warp));
        blocks = std::max<long long>(1, std::min(blocks, std::max<long long>(fit, 1)));
        L.blocks = (int)blocks;
        plan.dir_words += (size_t)(blocks * 4 * L.dir_words_per_warp);
        plan.band_words += (size_t)(blocks * 4 * L.band_words_per_warp);
        plan.launches.push_back(L);
    }
}

namespace {

__global__ void vm_fill_stats_kernel(const VmAlnJobDev *__restrict__ J, int nj, double *cells, double *bases)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    double c = 0, b = 0;
    if (j < nj && J[j].t.len > 0 && J[j].q.len > 0) { c = (double)J[j].t.len * (double)J[j].q.len; b = (double)J[j].t.len + (double)J[j].q.len; }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) { c += __shfl_xor_sync(VM_FULL, c, d); b += __shfl_xor_sync(VM_FULL, b, d); }
    if ((threadIdx.x & 31) == 0 && c > 0) { atomicAdd(cells, c); atomicAdd(bases, b); }
}

__global__ void vm_fill_redo_mask_kernel(const uint2 *__restrict__ results, int nj, uint8_t *mask, unsigned long long *count)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nj) return;
    const bool redo = results[j].x == 0xffffffffu;
    mask[j] = redo ? 1 : 0;
    if (redo) atomicAdd(count, 1ULL);
}

} // namespace

int vm_launch_fill_stats(const VmAlnJobDev *jobs, int nj, double *cells, double *bases, cudaStream_t stream)
{
    if (nj <= 0) return 0;
    vm_fill_stats_kernel<<<(nj + 255) / 256, 256, 0, stream>>>(jobs, nj, cells, bases);
    return 1;
}

int vm_launch_fill_redo_mask(const void *results, int nj, uint8_t *mask, unsigned long long *count, cudaStream_t stream)
{
    if (nj <= 0) return 0;
    vm_fill_redo_mask_kernel<<<(nj + 255) / 256, 256, 0, stream>>>((const uint2 *)results, nj, mask, count);
    return 1;
}
