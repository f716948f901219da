// Chain extraction on the device: the pointer-chasing halves of hit2work_1 (mammap_clrnano.py:23581-23640)
// and of the local traceback (:27508-27527), run where S / P / S_arg already live, so that only the chains the
// host glue goes on to use cross PCIe (a few hundred anchors per read) instead of the whole DP state
// (32 bytes per anchor globally, 20 locally).
//
// Both kernels are one thread per read: the walks are serial by definition (a chain claims its anchors in
// descending-score order, each step is a dependent load), there are thousands of reads per launch, and the
// whole thing is bounded by L2 latency, not by throughput.  Each thread builds its result in its own slice
// of a scratch arena (the read's anchor range), then claims room in a dense output arena with one atomicAdd
// and copies it over.
#include "vm_extract.cuh"

// ---- global stage: primary chain + residual chains with score > 40, in discovery order ----
__global__ void __launch_bounds__(64) vm_extract_global_kernel(const int *__restrict__ ids, int n_ids, const int64_t *__restrict__ off,
                                                               const int32_t *__restrict__ cnt, const VmAnchor *__restrict__ a_all,
                                                               const double *__restrict__ S_all, const int32_t *__restrict__ P_all,
                                                               const int32_t *__restrict__ A_all, const int64_t *__restrict__ gmax,
                                                               double accept, uint8_t *used_all, VmAnchor *tmp_anc, double *tmp_S,
                                                               int32_t *tmp_len, double *tmp_score, VmExtractOut out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_ids) return;
    const int r = ids[t];
    const int64_t o = off[r];
    const int n = cnt[r];
    const int64_t g = gmax[r];
    VmExtractRec rec;
    rec.anc_off = 0; rec.n_anc = 0; rec.meta_off = 0; rec.n_chains = 0;
    if (n <= 0 || g < 0) { out.rec[r] = rec; return; }
    const VmAnchor *a = a_all + o;
    const double *S = S_all + o;
    const int32_t *P = P_all + o, *A = A_all + o;
    uint8_t *used = used_all + o;            // zeroed by the host
    VmAnchor *ta = tmp_anc + o;
    double *tS = tmp_S + o, *tscore = tmp_score + o;
    int32_t *tlen = tmp_len + o;
    int k = 0, c = 0;
    bool hit = false;
    {
        int take = (int)g;
        used[take] = 1;
        const double score = S[take];
        for (;;) {
            ta[k] = a[take];
            tS[k] = S[take];
            ++k;
            const int p = P[take];
            if (p == VM_NOPRE) break;
            take = p;
            used[take] = 1;
        }
        if (score > 40) { hit = true; tlen[c] = k; tscore[c] = score; ++c; }
        else k = 0;
    }
    const double scores = S[g];
    const double max_scores = scores > 0 ? scores : 0;
    if (!(hit && max_scores > accept)) { out.rec[r] = rec; return; }   // nothing below can change the verdict
    for (int q = n - 1; q >= 0; --q) {
        int take = A[q];
        if (used[take]) continue;
        const int k0 = k;
        used[take] = 1;
        double score = S[take];
        for (;;) {
            ta[k] = a[take];
            tS[k] = 0.0;
            ++k;
            const int p = P[take];
            if (p == VM_NOPRE) break;
            take = p;
            if (used[take]) { score = score - S[take]; break; }
            used[take] = 1;
        }
        if (score > 40) { tlen[c] = k - k0; tscore[c] = score; ++c; }
        else k = k0;
    }
    rec.n_anc = k;
    rec.n_chains = c;
    rec.anc_off = (long long)atomicAdd(out.n_anc_total, (unsigned long long)k);
    rec.meta_off = (long long)atomicAdd(out.n_chain_total, (unsigned long long)c);
    for (int x = 0; x < k; ++x) { out.anc[rec.anc_off + x] = ta[x]; out.S[rec.anc_off + x] = tS[x]; }
    for (int x = 0; x < c; ++x) { out.chain_len[rec.meta_off + x] = tlen[x]; out.chain_score[rec.meta_off + x] = tscore[x]; }
    out.rec[r] = rec;
}

// ---- local stage: the best chain, overlapping anchors trimmed, ASCENDING read order ----
__global__ void __launch_bounds__(64) vm_extract_local_kernel(const int *__restrict__ ids, int n_ids, const int64_t *__restrict__ off,
                                                              const int32_t *__restrict__ cnt, const VmAnchor *__restrict__ a_all,
                                                              const double *__restrict__ S_all, const int32_t *__restrict__ P_all,
                                                              const int64_t *__restrict__ gmax, VmAnchor *tmp_anc, VmExtractOut out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_ids) return;
    const int r = ids[t];
    const int64_t o = off[r];
    const int n = cnt[r];
    const int64_t g = gmax[r];
    VmExtractRec rec;
    rec.anc_off = 0; rec.n_anc = 0; rec.meta_off = 0; rec.n_chains = 0;
    if (n <= 0 || g < 0) { out.rec[r] = rec; return; }
    const VmAnchor *a = a_all + o;
    const int32_t *P = P_all + o;
    VmAnchor *ta = tmp_anc + o;
    int k = 0;
    int take = (int)g;
    VmAnchor pre = a[take];
    ta[k++] = pre;
    for (;;) {
        const int p = P[take];
        if (p == VM_NOPRE) break;
        take = p;
        const VmAnchor now = a[take];
        if (pre.x < now.x + now.l) {          // :27514-27522 -- the later anchor gives up the overlap
            const int ov = now.x + now.l - pre.x;
            VmAnchor tr;
            tr.x = pre.x + ov;
            tr.y = pre.s == 1 ? pre.y + (uint32_t)ov : pre.y;
            tr.s = pre.s;
            tr.l = pre.l - ov;
            ta[k - 1] = tr;
        }
        ta[k++] = now;
        pre = now;
    }
    rec.n_anc = k;
    rec.n_chains = 1;
    rec.anc_off = (long long)atomicAdd(out.n_anc_total, (unsigned long long)k);
    if (out.chain_score) out.chain_score[r] = S_all[o + g];       // g_max_scores of the local DP (:27506)
    for (int x = 0; x < k; ++x) out.anc[rec.anc_off + x] = ta[k - 1 - x];
    out.rec[r] = rec;
}

int vm_launch_extract_global(const int *ids_dev, int n_ids, const int64_t *off, const int32_t *cnt, const VmAnchor *sorted,
                             const double *S, const int32_t *P, const int32_t *S_arg, const int64_t *gmax, double accept,
                             uint8_t *used_zeroed, VmAnchor *tmp_anc, double *tmp_S, int32_t *tmp_len, double *tmp_score,
                             const VmExtractOut &out, cudaStream_t stream)
{
    if (n_ids <= 0) return 0;
    vm_extract_global_kernel<<<(n_ids + 63) / 64, 64, 0, stream>>>(ids_dev, n_ids, off, cnt, sorted, S, P, S_arg, gmax, accept, used_zeroed,
                                                                  tmp_anc, tmp_S, tmp_len, tmp_score, out);
    return 1;
}

int vm_launch_extract_local(const int *ids_dev, int n_ids, const int64_t *off, const int32_t *cnt, const VmAnchor *sorted,
                            const double *S, const int32_t *P, const int64_t *gmax, VmAnchor *tmp_anc, const VmExtractOut &out,
                            cudaStream_t stream)
{
    if (n_ids <= 0) return 0;
    vm_extract_local_kernel<<<(n_ids + 63) / 64, 64, 0, stream>>>(ids_dev, n_ids, off, cnt, sorted, S, P, gmax, tmp_anc, out);
    return 1;
}

// ---- rebuild_chain_break (:23437-23484) on the extracted local path: colinear sub-alignments ----
// pos2contig (:51-59): last contig whose start <= pos (the first one if pos precedes all)
__device__ __forceinline__ int vm_cid(const int64_t *__restrict__ starts, int n, long long pos)
{
    int lo = 0, hi = n;           // first index with starts[i] > pos
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (starts[mid] <= pos) lo = mid + 1; else hi = mid;
    }
    return lo > 0 ? lo - 1 : 0;
}

__global__ void __launch_bounds__(64) vm_rebuild_kernel(const int *__restrict__ ids, int n_ids, const int64_t *__restrict__ off,
                                                        const VmExtractRec *__restrict__ rec, const VmAnchor *__restrict__ path_all,
                                                        const int64_t *__restrict__ ctg_start, int n_ctg, int large_cost,
                                                        int small_alignment, VmAnchor *tmp_anc, int32_t *tmp_len, VmRebuildOut out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_ids) return;
    const int r = ids[t];
    VmRebuildRec rr;
    rr.anc_off = 0; rr.len_off = 0; rr.n_anc = 0; rr.n_al = 0;
    const int n = rec[r].n_anc;
    if (n <= 1) { out.rec[r] = rr; return; }
    const VmAnchor *raw = path_all + rec[r].anc_off;       // ascending read order
    VmAnchor *oa = tmp_anc + off[r];                         // scratch: the read's slice of the anchor arena
    int32_t *ol = tmp_len + off[r];
    int k = 0, na = 0, cur = 0;                              // anchors written, sub-alignments closed, length of the open one
    VmAnchor pre = raw[0];
    oa[k++] = pre;
    cur = 1;
    for (int i = 1; i < n; ++i) {
        const VmAnchor now = raw[i];
        if (pre.s == now.s) {
            const long long readgap = (long long)now.x - pre.x - pre.l;
            const long long refgap = pre.s == 1 ? (long long)now.y - (long long)pre.y - pre.l : (long long)pre.y - (long long)now.y - now.l;
            long long d = readgap - refgap;
            if (d < 0) d = -d;
            if (d <= large_cost && refgap >= -20 && readgap < 100 &&
                vm_cid(ctg_start, n_ctg, (long long)pre.y) == vm_cid(ctg_start, n_ctg, (long long)now.y)) {
                if (refgap >= 0 || readgap > 20) { oa[k++] = now; ++cur; pre = now; }
                continue;                                    // small negative refgap with readgap <= 20: the anchor is dropped
            }
        }
        // break: close the open sub-alignment
        if (cur == 1) { --k; cur = 0; }                      // singleton: dropped
        else { ol[na++] = cur; cur = 0; }
        if (na > 0) {                                        // the (new) last one must span >= small_alignment read bases
            const int len = ol[na - 1];
            if (oa[k - 1].x + oa[k - 1].l - oa[k - len].x < small_alignment) { k -= len; --na; }
        }
        oa[k++] = now;
        cur = 1;
        pre = now;
    }
    if (cur == 1) { --k; cur = 0; }
    else if (cur > 1) { ol[na++] = cur; cur = 0; }
    if (na > 0) {
        const int len = ol[na - 1];
        if (oa[k - 1].x + oa[k - 1].l - oa[k - len].x < small_alignment) { k -= len; --na; }
    }
    rr.n_anc = k;
    rr.n_al = na;
    if (na > 0) {
        rr.anc_off = (long long)atomicAdd(out.n_anc_total, (unsigned long long)k);
        rr.len_off = (long long)atomicAdd(out.n_len_total, (unsigned long long)na);
        for (int x = 0; x < k; ++x) out.anc[rr.anc_off + x] = oa[x];
        for (int x = 0; x < na; ++x) out.len[rr.len_off + x] = ol[x];
    }
    out.rec[r] = rr;
}

int vm_launch_rebuild(const int *ids_dev, int n_ids, const int64_t *off, const VmExtractRec *rec, const VmAnchor *path,
                      const int64_t *ctg_start_dev, int n_ctg, int large_cost, int small_alignment, VmAnchor *tmp_anc, int32_t *tmp_len,
                      const VmRebuildOut &out, cudaStream_t stream)
{
    if (n_ids <= 0) return 0;
    vm_rebuild_kernel<<<(n_ids + 63) / 64, 64, 0, stream>>>(ids_dev, n_ids, off, rec, path, ctg_start_dev, n_ctg, large_cost,
                                                           small_alignment, tmp_anc, tmp_len, out);
    return 1;
}
