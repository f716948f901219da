// Base-level kernels: global edit distance, z-drop edge extension, global dual-affine fill.
// Job descriptors shared between the host backend and the kernels (see vm_align.cu).
#pragma once
#include "vm_common.cuh"
#include <vector>

// one side of an alignment job: a slice of the reference or of a read, optionally reversed /
// complemented on the fly (the reference materialises reversed / reverse-complemented strings,
// mammap_clrnano.py:2372-2375, 2401-2404, 2469-2472, 2497-2500)
struct VmSeqSpec {
    int64_t lo;        // start of the slice inside its source
    int32_t len;
    int32_t src;       // 0 reference, 1 read forward, 2 read reverse complement
    int32_t reverse;
    int32_t comp;
};

struct VmAlnJobDev {
    VmSeqSpec t, q;    // target, query
    int32_t read;
    int32_t n_out;     // fill: number of CIGAR ops written
    int64_t out_off;   // fill: offset of this job's CIGAR ops; edit distance: band half-width k (-1: none)
    int64_t dir_off;   // fill: offset of the direction matrix (bytes)
    int64_t sc_off;    // fill: offset of global score scratch (ints), or -1 when shared memory is used
    int64_t result0;   // edit distance: distance; extension: q_e
    int64_t result1;   // extension: t_e
};

struct VmSeqSources {
    const uint8_t *ref;
    const uint8_t *reads_fwd;
    const uint8_t *reads_rc;
    const int64_t *read_off;
};

#ifdef __CUDACC__
#include "vm_index.cuh"
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

#endif

// Edit distance: the job's band half-width k travels in out_off (-1 = unbanded); result0 is the exact distance
// when it is <= k, otherwise some value > k.  Jobs are grouped by the number of register slots per lane.
#define VM_ED_NCLASS 10
extern const int VM_ED_CLASS_G[VM_ED_NCLASS];
int vm_ed_slots(int m, int n, long long band);
int vm_launch_edit_distance(VmAlnJobDev *jobs, const int *ids_dev, const int *class_start, const int *class_words, VmSeqSources src,
                            cudaStream_t stream);
// upper bound of the distance through the job's match segments (J.dir_off / J.n_out into segs_dev, 12-byte {q, t, l});
// ids_dev == nullptr: jobs 0 .. n_jobs - 1
int vm_launch_ed_upper(VmAlnJobDev *jobs, const int *ids_dev, int n_jobs, const void *segs_dev, VmSeqSources src, cudaStream_t stream);
// derive the match segments of every job from its sub-alignment's anchors on the device (anc_dev[J.dir_off .. + J.n_out))
int vm_launch_match_segments(VmAlnJobDev *jobs, int n_jobs, const VmAnchor *anc_dev, void *segs_dev, cudaStream_t stream);
int vm_launch_extend(VmAlnJobDev *jobs, int n_jobs, VmSeqSources src, cudaStream_t stream);

// ---- global fill (vm_fill.cu) ----
// Two jobs share a warp (one per half of every half2 register); b = -1: no partner.
struct VmFillPair { int32_t a, b; };
// One launch = one capacity class: R target rows per lane (32*R rows per band), pairs [pair_begin, pair_end)
struct VmFillLaunch {
    int R, multiband, pair_begin, pair_end, blocks;
    long long dir_words_per_warp, band_words_per_warp;
};
struct VmFillPlan {
    std::vector<VmFillPair> pairs;
    std::vector<VmFillLaunch> launches;
    size_t dir_words = 0, band_words = 0;   // scratch needed (uint32 words), one slice per launch
    double dir_bytes = 0;                   // direction bytes the launches will store (the kernel's dominant traffic)
};
// only_mask != nullptr: plan only the jobs j with only_mask[j] != 0
void vm_fill_plan(const VmAlnJobDev *jobs_host, int n_jobs, int sm_count, VmFillPlan &plan, int host_threads = 1,
                  const uint8_t *only_mask = nullptr);
// counters_dev: one zeroed int per launch.  The CIGAR ops of job j end up in dense_out[results[j].x .. + results[j].y)
// (results: uint2 per job, zero-initialised by the caller; dense_count: zeroed 64-bit bump allocator;
// cigar_scratch: per-job room of tlen + qlen + 2 ops at J.out_off).  Returns the number of kernel launches.
// Launches of fewer than VM_FILL_SMALL_BLOCKS blocks go round-robin (*side_rr) to the `side` streams (the caller
// orders them against main_stream with events); *dir_cursor / *band_cursor: words of the scratch arenas already
// handed to earlier launches (every launch gets its own slice).
#define VM_FILL_SMALL_BLOCKS 64
int vm_fill_launch(const VmFillPlan &plan, VmAlnJobDev *jobs_dev, const VmFillPair *pairs_dev, VmSeqSources src, int eqx,
                   uint32_t *dir_scratch, uint32_t *band_scratch, int *counters_dev, uint32_t *cigar_scratch, uint32_t *dense_out,
                   unsigned long long *dense_count, void *results, cudaStream_t main_stream, const cudaStream_t *side, int n_side,
                   int *side_rr, size_t *dir_cursor, size_t *band_cursor);

// ---- banded global fill with an optimality certificate (vm_fillb.cu) ----
#define VM_FB_NCLASS 12       // slot classes (vm_fillb.cu: VM_FB_CLASS), by band rows per anti-diagonal
// two jobs sharing a warp and a band of diagonals [kmin, kmax] (b = -1: no partner)
struct VmFillBandPair { int32_t a, b, kmin, kmax; };
struct VmFillBandLaunch {
    int cls, pair_begin, pair_end, blocks;    // slot class: (pairs per warp G, register slots per lane C), 32 / G * C band rows
    long long dir_words_per_warp;
};
struct VmFillBandPlan {
    std::vector<VmFillBandPair> pairs;
    std::vector<VmFillBandLaunch> launches;
    size_t dir_words = 0;
    double dir_bytes = 0;                   // direction bytes the launches will store
};
// false: the job is left to the full-matrix kernel (too small to gain, or too long for the staging)
bool vm_fillb_own_band(int tlen, int qlen, int &kmin, int &kmax);
// plans every job the banded kernel can take and clears its full_mask entry
void vm_fillb_plan(const VmAlnJobDev *jobs_host, int n_jobs, int sm_count, int host_threads, VmFillBandPlan &plan, uint8_t *full_mask);
// as vm_fill_launch; a job whose certificate fails gets results[j] = (0xffffffff, 0) and must be re-run unbanded
int vm_fillb_launch(const VmFillBandPlan &plan, VmAlnJobDev *jobs_dev, const VmFillBandPair *pairs_dev, VmSeqSources src, int eqx,
                    uint32_t *dir_scratch, int *counters_dev, uint32_t *cigar_scratch, uint32_t *dense_out,
                    unsigned long long *dense_count, void *results, cudaStream_t main_stream, const cudaStream_t *side, int n_side,
                    int *side_rr, size_t *dir_cursor);

// ---- the same plans made on the device (the jobs never visit the host) ----
#include "vm_ctx.cuh"
struct VmFbTable {          // per slot class 1..VM_FB_NCLASS of the banded kernel
    int32_t cnt[16], at[16], max_steps[16];
    unsigned long long sum_steps[16];
    int32_t n_live, n_pairs;
};
struct VmFfTable {          // per capacity class 0..26 of the full-matrix kernel
    int32_t n[32], pair_begin[32], max_q[32], max_t[32];
    int32_t n_live, n_pairs;
    double dir_bytes;
};
struct VmFillPlanBufs {
    VmDevBuf keys, bmin, bmax, start, cursor, order, tpairs, keys2, start2, cursor2, order2, small;
    VmPinnedBuf table;
    void release()
    {
        VmDevBuf *d[] = {&keys, &bmin, &bmax, &start, &cursor, &order, &tpairs, &keys2, &start2, &cursor2, &order2, &small};
        for (VmDevBuf *x : d) x->release();
        table.release();
    }
};
// Asynchronous part: plans every job the banded kernel can take.  pairs_out_dev (room for n_jobs pairs) receives the
// pairs bucketed by slot class; full_mask_dev[j] = 1 for valid jobs left to the full-matrix kernel; B.table (pinned host,
// a VmFbTable) is valid once the stream is synchronised.  Returns the number of kernel launches, < 0 on a CUDA error.
int vm_fillb_plan_dev(const VmAlnJobDev *jobs_dev, int n_jobs, VmFillPlanBufs &B, VmFillBandPair *pairs_out_dev, uint8_t *full_mask_dev,
                      cudaStream_t stream);
// launches of the plan from the table (pairs stay on the device: plan.pairs is left empty)
void vm_fillb_plan_finish(const VmFillPlanBufs &B, int sm_count, VmFillBandPlan &plan);
// Full-matrix kernel, jobs with only_mask_dev[j] != 0: pairs_out_dev gets the pairs by capacity class, B.table a VmFfTable.
int vm_fill_plan_dev(const VmAlnJobDev *jobs_dev, int n_jobs, const uint8_t *only_mask_dev, VmFillPlanBufs &B, VmFillPair *pairs_out_dev,
                     cudaStream_t stream);
void vm_fill_plan_finish(const VmFillPlanBufs &B, int sm_count, VmFillPlan &plan);
// sum over the valid jobs of tlen * qlen (cells) and tlen + qlen (bases), added to the two device doubles
int vm_launch_fill_stats(const VmAlnJobDev *jobs_dev, int n_jobs, double *cells_dev, double *bases_dev, cudaStream_t stream);
// mask_dev[j] = 1 for the jobs whose banded certificate failed (results[j].x == 0xffffffff), 0 otherwise; their number is
// added to *count_dev
int vm_launch_fill_redo_mask(const void *results_dev, int n_jobs, uint8_t *mask_dev, unsigned long long *count_dev, cudaStream_t stream);
