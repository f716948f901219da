// Local 9-mer re-seeding on the GPU.
//
// Replaces get_localmap_multi_all_forDP_inv_guide_1 (mammap_clrnano.py:23069-23345) from the
// point where the reference windows are known (window construction is host glue).  The
// reference builds a per-read hash table of every 9-mer of the windows and probes it with every
// read position; here the 9-mer positions of the WHOLE reference sit in HBM sorted by
// (code, position) (vm_index.cuh), so a probe is a range query: lower_bound(window start) in
// the code's position run, then walk while inside the window -- the same occurrences, in the
// same (window, ascending position) order the reference inserted them, duplicates from
// overlapping windows included.
//
//   kernel 1 (vm_reseed_hits_kernel): one block per guide job; each thread owns one read
//     position per chunk, applies the guide-proximity filter (:23216-23231) and writes its hits
//     at an offset obtained from a block scan, so hits are ordered exactly like the reference's
//     scan: (read position, forward before reverse, window, reference position).  HBM-bound
//     gather: 2 strands x (2 offsets + ~log2(run) probes) per read position.
//   kernel 2 (vm_reseed_merge_kernel): one warp per job; the same-diagonal merge (:23235-23252,
//     23294-23312) with the diagonals split between the lanes (each replays its diagonals' hits in
//     scan order), the diagonal table (`pointdict`) in global memory, and the reference's emission
//     order restored by slotting every anchor at the scan index that emitted it.
#include "vm_reseed.cuh"

struct VmHit {
    int32_t iloc_s;    // iloc << 1 | (strand == -1)
    uint32_t refloc;
};

__device__ __forceinline__ int vm_kmer_code(const uint8_t *s)
{
    int code = 0;
#pragma unroll
    for (int i = 0; i < VM_K9; ++i) code = code * 5 + vm_code5(s[i]);
    return code;
}

// findClosest_1 :17560-17581
__device__ __forceinline__ void vm_find_closest(const int32_t *arr, int n, int target, int &b0, int &b1, int &i0, int &i1)
{
    if (target <= arr[0]) { b0 = b1 = arr[0] - target; i0 = i1 = 0; return; }
    if (target >= arr[n - 1]) { b0 = b1 = target - arr[n - 1]; i0 = i1 = n - 1; return; }
    int i = 0, j = n;
    while (i < j) {
        const int mid = (i + j) >> 1;
        const int v = arr[mid];
        if (v == target) { b0 = b1 = 0; i0 = i1 = mid; return; }
        if (target < v) j = mid; else i = mid + 1;
    }
    b0 = abs(arr[j - 1] - target);
    b1 = abs(arr[j] - target);
    i0 = j - 1;
    i1 = j;
}

// visit every accepted (refloc) of one k-mer code for one read position, in reference order
template <typename F>
__device__ __forceinline__ void vm_probe(const VmIndexDev &ix, int code, const int64_t *win_lo, const int64_t *win_hi, int n_win,
                                         long long rgap, long long r1, long long r2, long long interval, F visit)
{
    const int64_t b = ix.koff[code], e = ix.koff[code + 1];
    if (b == e) return;
    for (int w = 0; w < n_win; ++w) {
        const long long lo = win_lo[w], hi = win_hi[w] - VM_K9;   // k-mer start must be <= hi
        int64_t l = vm_kpos_lower_bound(ix, code, b, e, lo);
        for (; l < e; ++l) {
            const long long refloc = ix.kpos[l];
            if (refloc > hi) break;
            long long d = refloc - r1;
            if (d < 0) d = -d;
            long long diff = rgap - d;
            if (diff < 0) diff = -diff;
            if (diff < 500 || (r1 + interval >= refloc && r1 - interval <= refloc) ||
                (r2 + interval >= refloc && r2 - interval <= refloc))
                visit((uint32_t)refloc);
        }
    }
}

// 9-mer code of the staged 5-letter codes c[0..8]
__device__ __forceinline__ int vm_kmer_code_staged(const uint8_t *c)
{
    int code = 0;
#pragma unroll
    for (int i = 0; i < VM_K9; ++i) code = code * 5 + c[i];
    return code;
}

#define VM_RS_THREADS 256
#define VM_RS_KEEP 3      // hits a thread keeps in registers between counting and writing

// Single pass: every thread probes its read position once, keeps its first VM_RS_KEEP hits in registers, the
// block scans the counts (warp shuffles + one shared-memory exchange) and the hits are written in the
// reference's scan order.  A job whose hits exceed its capacity is only counted (n_hits > hit_cap tells the
// host to re-run it with the exact capacity).
__global__ void __launch_bounds__(VM_RS_THREADS) vm_reseed_hits_kernel(VmIndexDev ix, const VmReseedJobDev *__restrict__ jobs,
                                                                       const uint8_t *__restrict__ reads_fwd,
                                                                       const uint8_t *__restrict__ reads_rc,
                                                                       const int64_t *__restrict__ read_off,
                                                                       const int64_t *__restrict__ win_lo_all,
                                                                       const int64_t *__restrict__ win_hi_all,
                                                                       const int32_t *__restrict__ gx_all,
                                                                       const int64_t *__restrict__ gy_all,
                                                                       VmHit *__restrict__ hits_all, int32_t *__restrict__ n_hits)
{
    __shared__ int s_warp[VM_RS_THREADS / 32];
    __shared__ uint8_t s_f[VM_RS_THREADS + VM_K9], s_r[VM_RS_THREADS + VM_K9];
    const VmReseedJobDev J = jobs[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t rbase = read_off[J.read];
    const int L = (int)(read_off[J.read + 1] - rbase);
    const uint8_t *seq = (J.need_reverse ? reads_rc : reads_fwd) + rbase;     // oriented testseq
    const uint8_t *rcs = (J.need_reverse ? reads_fwd : reads_rc) + rbase;     // oriented rc_testseq
    const int64_t *win_lo = win_lo_all + J.win_off, *win_hi = win_hi_all + J.win_off;
    const int32_t *gx = gx_all + J.g_off;
    const int64_t *gy = gy_all + J.g_off;
    VmHit *hits = hits_all + J.hit_off;
    int running = 0;
    for (int i0 = J.readstart; i0 < J.readend; i0 += VM_RS_THREADS) {
        // stage the 5-letter codes of this chunk: forward bases [i0, i0+264), and the reverse-complement strand
        // bases the chunk's k-mers read, s_r[u] = rc[L - i0 - 9 - (THREADS - 1) + u]
        __syncthreads();
        for (int u = tid; u < VM_RS_THREADS + VM_K9; u += VM_RS_THREADS) {
            const int pf = i0 + u;
            s_f[u] = pf < L ? (uint8_t)vm_code5(seq[pf]) : (uint8_t)4;
            const int pr = L - i0 - VM_K9 - (VM_RS_THREADS - 1) + u;
            s_r[u] = (pr >= 0 && pr < L) ? (uint8_t)vm_code5(rcs[pr]) : (uint8_t)4;
        }
        __syncthreads();
        const int iloc = i0 + tid;
        int cnt = 0;
        int fcode = -1, rcode = -1;
        long long rgap = 0, r1 = 0, r2 = 0, interval = 0;
        uint32_t keep[VM_RS_KEEP];
        unsigned keep_rev = 0;
        if (iloc < J.readend) {
            fcode = vm_kmer_code_staged(s_f + tid);
            const bool have_rev = iloc != 0;                      // rc[-(0+k):-0] == '' (:23212)
            if (have_rev) rcode = vm_kmer_code_staged(s_r + (VM_RS_THREADS - 1 - tid));
            if (have_rev && fcode == rcode) { fcode = -1; rcode = -1; }   // palindrome: skip the position
            else {
                int b0, b1, c0, c1;
                vm_find_closest(gx, J.n_guide, iloc, b0, b1, c0, c1);
                interval = b0 + b1 + 500;
                if (interval > 2000) interval = 2000;
                r1 = gy[c0];
                r2 = gy[c1];
                rgap = iloc - gx[c0];
                if (rgap < 0) rgap = -rgap;
                if (fcode >= 0)
                    vm_probe(ix, fcode, win_lo, win_hi, J.n_win, rgap, r1, r2, interval, [&](uint32_t refloc) {
#pragma unroll
                        for (int t = 0; t < VM_RS_KEEP; ++t)
                            if (cnt == t) keep[t] = refloc;
                        ++cnt;
                    });
                if (rcode >= 0)
                    vm_probe(ix, rcode, win_lo, win_hi, J.n_win, rgap, r1, r2, interval, [&](uint32_t refloc) {
#pragma unroll
                        for (int t = 0; t < VM_RS_KEEP; ++t)
                            if (cnt == t) { keep[t] = refloc; keep_rev |= 1u << t; }
                        ++cnt;
                    });
            }
        }
        // block-wide exclusive scan of cnt
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(VM_FULL, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < VM_RS_THREADS / 32; ++w) {
            const int t = s_warp[w];
            if (w < warp) before += t;
            total += t;
        }
        int o = running + before + incl - cnt;
        if (cnt > 0 && (long long)running + total <= (long long)J.hit_cap) {
            if (cnt <= VM_RS_KEEP) {
#pragma unroll
                for (int t = 0; t < VM_RS_KEEP; ++t)
                    if (t < cnt) { hits[o + t].iloc_s = iloc << 1 | (int)((keep_rev >> t) & 1u); hits[o + t].refloc = keep[t]; }
            } else {
                if (fcode >= 0)
                    vm_probe(ix, fcode, win_lo, win_hi, J.n_win, rgap, r1, r2, interval,
                             [&](uint32_t refloc) { hits[o].iloc_s = iloc << 1; hits[o].refloc = refloc; ++o; });
                if (rcode >= 0)
                    vm_probe(ix, rcode, win_lo, win_hi, J.n_win, rgap, r1, r2, interval,
                             [&](uint32_t refloc) { hits[o].iloc_s = iloc << 1 | 1; hits[o].refloc = refloc; ++o; });
            }
        }
        running += total;
    }
    if (tid == 0) n_hits[blockIdx.x] = running;
}

// The diagonal table of a job: tab_size keys (64-bit, contiguous: cleared with coalesced stores) followed by tab_size values
struct VmPointVal {
    int c0;
    unsigned c1;
    int c23;       // segment length | (strand == -1) << 8
    int q0;        // scan index of the hit that opened the diagonal (the reference's dict insertion order)
};
struct VmPoint { long long key; VmPointVal v; };      // 24 bytes per slot (sizing only)
#define VM_PT_EMPTY 0x7fffffffffffffffLL
__device__ __forceinline__ long long *vm_tab_keys(VmPoint *table_all, const VmReseedJobDev &J) { return (long long *)((char *)table_all + J.tab_off * 24); }
__device__ __forceinline__ VmPointVal *vm_tab_vals(VmPoint *table_all, const VmReseedJobDev &J)
{
    return (VmPointVal *)((char *)table_all + J.tab_off * 24 + (long long)J.tab_size * 8);
}

// Same-diagonal merge (:23235-23252, 23294-23312), one warp per job.
//
// The reference walks the hits in scan order; every hit touches only the running segment of its own diagonal
// (`pointdict[point]`), emits at most one anchor (when the segment is cut by a gap or reaches 20 bases), and the
// segments still open at the end are emitted in the order their diagonals first appeared.  Diagonals are
// independent, so the lanes split them: lane `hash(point) & 31` owns a diagonal and replays its hits in scan order,
// 32 hits (one per lane) being staged at a time; the table latency of up to 32 diagonals overlaps instead of
// being paid one hit after the other.  Every lane has its own interleaved slice of the job's table (slots
// lane, lane + 32, ...): no atomics.  The reference's output ORDER is kept without sorting: an anchor emitted while
// processing hit q goes to slot q of a sparse array, a left-over segment to slot n + q0, and a final in-place
// compaction of the 2n slots (in slot order) yields exactly the sequential emission order.
// flags: one byte per slot (the job's slice of the `order` scratch, 4 bytes per hit).
// A lane whose slice fills up gives up: n_out = -1, and the job is redone by the sequential kernel below.
__global__ void __launch_bounds__(32) vm_reseed_merge_kernel(const VmReseedJobDev *__restrict__ jobs,
                                                             const VmHit *__restrict__ hits_all,
                                                             const int32_t *__restrict__ n_hits,
                                                             VmPoint *__restrict__ table_all, int32_t *__restrict__ order_all,
                                                             VmAnchor *__restrict__ out_all, int32_t *__restrict__ n_out)
{
    __shared__ VmHit s_hit[32];
    __shared__ unsigned s_mask[32];
    const VmReseedJobDev J = jobs[blockIdx.x];
    const int lane = threadIdx.x;
    const int n = n_hits[blockIdx.x];
    if (n > J.hit_cap || n <= 0) { if (lane == 0) n_out[blockIdx.x] = 0; return; }   // over capacity: never after the host's re-run
    const VmHit *hits = hits_all + J.hit_off;
    long long *keys = vm_tab_keys(table_all, J);
    VmPointVal *vals = vm_tab_vals(table_all, J);
    const int sub = J.tab_size >> 5;                               // slots of one lane's slice (a power of two)
    uint8_t *flag = (uint8_t *)(order_all + J.dense_off);          // [2n]
    VmAnchor *out = out_all + 2 * J.dense_off;                     // [2n] sparse slots, compacted in place at the end
    for (int t = lane; t < J.tab_size; t += 32) keys[t] = VM_PT_EMPTY;
    for (int t = lane; t < (2 * n + 3) / 4; t += 32) ((uint32_t *)flag)[t] = 0u;
    __syncwarp();
    const int k = VM_K9;
    int used = 0;
    bool full = false;
    // The running segment of the diagonal this lane touched last stays in registers: consecutive hits of one diagonal
    // (long exact stretches -- most of a HiFi read) cost no table round trip; it goes back to the table when the lane
    // moves on to another diagonal and before the left-over pass.
    long long cpoint = VM_PT_EMPTY;
    int cslot = 0, c0 = 0, c2 = 1, c3 = 0;
    unsigned c1 = 0;
    bool dirty = false;
    auto flush = [&]() {
        if (dirty) {
            VmPointVal *e = vals + cslot;
            e->c0 = c0; e->c1 = c1; e->c23 = c3 | (c2 == 1 ? 0 : 256);
            dirty = false;
        }
    };
    for (int q0 = 0; q0 < n; q0 += 32) {
        const int q = q0 + lane;
        int own = -1;
        if (q < n) {
            const VmHit h = hits[q];
            s_hit[lane] = h;
            const int iloc = h.iloc_s >> 1;
            const long long point = (h.iloc_s & 1) ? -((long long)h.refloc + iloc) : (long long)h.refloc - iloc;
            const unsigned long long hh = (unsigned long long)point * 0x9E3779B97F4A7C15ULL;
            own = (int)((hh >> 40) & 31u);
        }
        s_mask[lane] = 0u;
        __syncwarp();
        // lane L learns which of the 32 staged hits it owns: every group of equal owners reports its member mask
        const unsigned peers = __match_any_sync(VM_FULL, own);
        if (own >= 0 && (int)(__ffs(peers) - 1) == lane) s_mask[own] = peers;
        __syncwarp();
        unsigned mine = s_mask[lane];
        while (mine && !full) {
            const int t = __ffs(mine) - 1;
            mine &= mine - 1;
            const VmHit hit = s_hit[t];
            const int qq = q0 + t;
            const int iloc = hit.iloc_s >> 1;
            const int strand = (hit.iloc_s & 1) ? -1 : 1;
            const long long refloc = hit.refloc;
            const long long point = strand == 1 ? refloc - iloc : -(refloc + iloc);
            if (point != cpoint) {
                // another diagonal of this lane: put the running segment back, fetch (or open) the new one
                flush();
                const unsigned long long h = (unsigned long long)point * 0x9E3779B97F4A7C15ULL;
                int s = (int)((h ^ (h >> 31)) & (unsigned long long)(sub - 1));
                long long key;
                while ((key = keys[(s << 5) + lane]) != VM_PT_EMPTY && key != point) s = (s + 1) & (sub - 1);
                cslot = (s << 5) + lane;
                cpoint = point;
                if (key == VM_PT_EMPTY) {
                    if (++used >= sub) { full = true; break; }         // keep one slot free: the probe loop must terminate
                    keys[cslot] = point;
                    c0 = iloc; c1 = (unsigned)refloc; c2 = strand; c3 = k;
                    VmPointVal nv; nv.c0 = c0; nv.c1 = c1; nv.c23 = k | (strand == 1 ? 0 : 256); nv.q0 = qq;
                    vals[cslot] = nv;
                    continue;
                }
                const VmPointVal pv = vals[cslot];
                c0 = pv.c0; c1 = pv.c1; c3 = pv.c23 & 255; c2 = (pv.c23 & 256) ? -1 : 1;
            }
            if (c0 + c3 >= iloc) {
                const int bonus = iloc - (c0 + c3) + k;
                if (bonus > 0) {
                    if (c3 + bonus < 20) {
                        if (strand == 1) { c2 = 1; c3 += bonus; }
                        else { c1 = (unsigned)refloc; c2 = -1; c3 += bonus; }
                    } else {
                        VmAnchor a; a.x = c0; a.y = c1; a.s = c2; a.l = c3;
                        out[qq] = a;
                        flag[qq] = 1;
                        if (strand == 1) { const int l3 = c3; c0 += l3; c1 += (unsigned)l3; c2 = 1; c3 = bonus; }
                        else { c0 += c3; c1 = (unsigned)refloc; c2 = -1; c3 = bonus; }
                    }
                    dirty = true;
                }
            } else {
                VmAnchor a; a.x = c0; a.y = c1; a.s = c2; a.l = c3;
                out[qq] = a;
                flag[qq] = 1;
                c0 = iloc; c1 = (unsigned)refloc; c2 = strand; c3 = k;
                dirty = true;
            }
        }
        if (__any_sync(VM_FULL, full)) { if (lane == 0) n_out[blockIdx.x] = -1; return; }
    }
    flush();
    __syncwarp();
    // the segments still open, at slot n + (scan index of their diagonal's first hit)
    for (int t = lane; t < J.tab_size; t += 32) {
        if (keys[t] == VM_PT_EMPTY) continue;
        const VmPointVal p = vals[t];
        VmAnchor a; a.x = p.c0; a.y = p.c1; a.s = (p.c23 & 256) ? -1 : 1; a.l = p.c23 & 255;
        out[n + p.q0] = a;
        flag[n + p.q0] = 1;
    }
    __syncwarp();
    // in-place compaction in slot order, 128 slots per step (the write position never overtakes the read position)
    int m = 0;
    const int n_words = (2 * n + 3) / 4;
    for (int b = 0; b < n_words; b += 32) {
        const int w = b + lane;
        const uint32_t f = w < n_words ? ((const uint32_t *)flag)[w] : 0u;
        const int base = w * 4;
        VmAnchor a[4];
        int cnt = 0;
#pragma unroll
        for (int t = 0; t < 4; ++t)
            if ((f >> (8 * t)) & 0xffu) { a[cnt] = out[base + t]; ++cnt; }
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(VM_FULL, incl, d);
            if (lane >= d) incl += o;
        }
        const int total = __shfl_sync(VM_FULL, incl, 31);
        __syncwarp();
        int o = m + incl - cnt;
        for (int t = 0; t < cnt; ++t) out[o + t] = a[t];
        m += total;
        __syncwarp();
    }
    if (lane == 0) n_out[blockIdx.x] = m;
}

// The same merge replayed by one lane, hit after hit (the formulation above falls back to it for a job whose hits
// crowd one lane's table slice; `only_failed`: run only the jobs marked n_out = -1).
__global__ void __launch_bounds__(32) vm_reseed_merge_seq_kernel(const VmReseedJobDev *__restrict__ jobs,
                                                                 const VmHit *__restrict__ hits_all,
                                                                 const int32_t *__restrict__ n_hits,
                                                                 VmPoint *__restrict__ table_all, int32_t *__restrict__ order_all,
                                                                 VmAnchor *__restrict__ out_all, int32_t *__restrict__ n_out)
{
    if (n_out[blockIdx.x] != -1) return;
    const VmReseedJobDev J = jobs[blockIdx.x];
    const int lane = threadIdx.x;
    const int n = n_hits[blockIdx.x];
    const VmHit *hits = hits_all + J.hit_off;
    long long *keys = vm_tab_keys(table_all, J);
    VmPointVal *vals = vm_tab_vals(table_all, J);
    const int tmask = J.tab_size - 1;
    int32_t *order = order_all + J.dense_off;
    VmAnchor *out = out_all + 2 * J.dense_off;
    for (int t = lane; t < J.tab_size; t += 32) keys[t] = VM_PT_EMPTY;
    __syncwarp();
    if (lane != 0) return;
    int n_pts = 0, m = 0;
    const int k = VM_K9;
    for (int q = 0; q < n; ++q) {
        const VmHit hit = hits[q];
        const int iloc = hit.iloc_s >> 1;
        const int strand = (hit.iloc_s & 1) ? -1 : 1;
        const long long refloc = hit.refloc;
        const long long point = strand == 1 ? refloc - iloc : -(refloc + iloc);
        unsigned long long h = (unsigned long long)point * 0x9E3779B97F4A7C15ULL;
        int s = (int)((h ^ (h >> 31)) & (unsigned long long)tmask);
        while (keys[s] != VM_PT_EMPTY && keys[s] != point) s = (s + 1) & tmask;
        VmPointVal p = vals[s];
        int c3 = p.c23 & 255, c2 = (p.c23 & 256) ? -1 : 1;
        if (keys[s] == VM_PT_EMPTY) {
            keys[s] = point; p.c0 = iloc; p.c1 = (unsigned)refloc; c2 = strand; c3 = k; p.q0 = q;
            order[n_pts++] = s;
        } else if (p.c0 + c3 >= iloc) {
            const int bonus = iloc - (p.c0 + c3) + k;
            if (bonus <= 0) continue;
            if (c3 + bonus < 20) {
                if (strand == 1) { c2 = 1; c3 += bonus; }
                else { p.c1 = (unsigned)refloc; c2 = -1; c3 += bonus; }
            } else {
                VmAnchor a; a.x = p.c0; a.y = p.c1; a.s = c2; a.l = c3;
                out[m++] = a;
                if (strand == 1) { p.c0 += c3; p.c1 += (unsigned)c3; c2 = 1; c3 = bonus; }
                else { p.c0 += c3; p.c1 = (unsigned)refloc; c2 = -1; c3 = bonus; }
            }
        } else {
            VmAnchor a; a.x = p.c0; a.y = p.c1; a.s = c2; a.l = c3;
            out[m++] = a;
            p.c0 = iloc; p.c1 = (unsigned)refloc; c2 = strand; c3 = k;
        }
        p.c23 = c3 | (c2 == 1 ? 0 : 256);
        vals[s] = p;
    }
    for (int q = 0; q < n_pts; ++q) {
        const VmPointVal p = vals[order[q]];
        VmAnchor a; a.x = p.c0; a.y = p.c1; a.s = (p.c23 & 256) ? -1 : 1; a.l = p.c23 & 255;
        out[m++] = a;
    }
    n_out[blockIdx.x] = m;
}

int vm_reseed_launch(const VmIndexDev &ix, const VmReseedJobDev *jobs_dev, int n_jobs, const uint8_t *reads_fwd,
                     const uint8_t *reads_rc, const int64_t *read_off, const int64_t *win_lo, const int64_t *win_hi,
                     const int32_t *gx, const int64_t *gy, void *hits, int32_t *n_hits, cudaStream_t stream)
{
    if (n_jobs <= 0) return 0;
    vm_reseed_hits_kernel<<<n_jobs, VM_RS_THREADS, 0, stream>>>(ix, jobs_dev, reads_fwd, reads_rc, read_off, win_lo, win_hi, gx, gy,
                                                                (VmHit *)hits, n_hits);
    return 1;
}

int vm_reseed_merge_launch(const VmReseedJobDev *jobs_dev, int n_jobs, const void *hits, const int32_t *n_hits, void *table,
                           int32_t *order, VmAnchor *out, int32_t *n_out, cudaStream_t stream)
{
    if (n_jobs <= 0) return 0;
    vm_reseed_merge_kernel<<<n_jobs, 32, 0, stream>>>(jobs_dev, (const VmHit *)hits, n_hits, (VmPoint *)table, order, out, n_out);
    // jobs that gave up (n_out = -1: one lane's table slice filled up) are replayed hit after hit; everyone else returns at once
    vm_reseed_merge_seq_kernel<<<n_jobs, 32, 0, stream>>>(jobs_dev, (const VmHit *)hits, n_hits, (VmPoint *)table, order, out, n_out);
    return 2;
}

size_t vm_reseed_hit_bytes() { return sizeof(VmHit); }
size_t vm_reseed_point_bytes() { return sizeof(VmPoint); }
