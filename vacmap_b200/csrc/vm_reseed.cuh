// Local re-seeding stage: job descriptor and launch entry points (see vm_reseed.cu).
#pragma once
#include "vm_common.cuh"
#include "vm_index.cuh"

struct VmReseedJobDev {
    int32_t read;
    int32_t need_reverse;   // the per-read driver swapped testseq / rc_testseq (:24063-24065)
    int32_t readstart, readend;
    int32_t n_win, n_guide;
    int64_t win_off;        // into the flattened window arrays
    int64_t g_off;          // into the flattened guide arrays
    int64_t hit_off;        // into the hit buffer
    int32_t hit_cap;        // hits the job may write; more are only counted
    int32_t pad0;
    int64_t dense_off;      // prefix sum of the actual hit counts: order[] uses it, out[] uses 2 * dense_off
    int64_t tab_off;        // diagonal table
    int32_t tab_size;       // power of two > number of hits
    int32_t pad;
};

int vm_reseed_launch(const VmIndexDev &ix, const VmReseedJobDev *jobs_dev, int n_jobs, const uint8_t *reads_fwd,
                     const uint8_t *reads_rc, const int64_t *read_off, const int64_t *win_lo, const int64_t *win_hi,
                     const int32_t *gx, const int64_t *gy, void *hits, int32_t *n_hits, cudaStream_t stream);
int vm_reseed_merge_launch(const VmReseedJobDev *jobs_dev, int n_jobs, const void *hits, const int32_t *n_hits, void *table,
                           int32_t *order, VmAnchor *out, int32_t *n_out, cudaStream_t stream);
size_t vm_reseed_hit_bytes();
size_t vm_reseed_point_bytes();
