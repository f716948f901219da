/*
 * vacmap_b200.h -- C ABI of libvacmap_b200.so (CUDA, sm_100a).
 *
 * Drop-in boundary for VACmap's per-read alignment hot path.  The reference has
 * no native code in-tree; its FFI seam is the un-vendored `vacmap_index` C
 * extension (seeding + k_cigar) plus numba-JIT functions.  Each entry point
 * below names the reference interface it replaces (file:line under
 * /root/reference/src/vacmap/, mode H module mammap_clrnano.py unless noted).
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are HOST pointers unless the
 *     name says `_dev`.  Inputs are borrowed for the duration of the call.
 *   - every function returns VM_OK (0) or a negative VM_ERR_* code and never
 *     throws across the ABI; vm_last_error(ctx) gives the message.
 *   - one vm_ctx per GPU / host thread; calls on one ctx are serialised on the
 *     ctx's CUDA stream.  There is NO CPU fallback: without a CUDA device
 *     vm_ctx_create fails with VM_ERR_NO_DEVICE.
 *   - anchors are the reference's int64[n][4] rows
 *     (readpos_start, refpos_global_leftmost, strand +1/-1, len)   (:23985)
 *   - batches are ragged: `off[n_reads+1]` gives row offsets into the
 *     concatenated per-read arrays.
 */
#ifndef VACMAP_B200_H
#define VACMAP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VM_OK 0
#define VM_ERR_CUDA (-1)
#define VM_ERR_NO_DEVICE (-2)
#define VM_ERR_ARG (-3)
#define VM_ERR_NOMEM (-4)
#define VM_ERR_STATE (-5)

#define VM_NOPRE (-9999999) /* chain-start marker in P[] (:24891) */

typedef struct vm_ctx vm_ctx;

/* ABI version of this header; bumped on any signature change. */
int vm_abi_version(void);

/* Create / destroy a per-GPU context (CUDA stream, scratch arenas, tables). */
int vm_ctx_create(int device, vm_ctx **out);
void vm_ctx_destroy(vm_ctx *ctx);
const char *vm_last_error(vm_ctx *ctx);
/* Number of kernels launched by this ctx since creation (bench `gpu_launches`). */
int64_t vm_kernel_launches(vm_ctx *ctx);

/*
 * Score tables.  The reference builds these with numpy at module import
 * (`extra` :15371-15376 float32; `readgapcost_list` :26567-26569 float32[100];
 * `log2cache` :27530 float64[100000]) or inside the njit functions with libm
 * log2 (`gapcost_list` :24843-24846 / :27317-27322, `large_readgapcost_list`
 * :28270-28275).  The host passes the module-level ones so device scores are
 * bit-identical to the host's; the in-function ones are rebuilt by the library
 * with libm, as numba does.
 */
int vm_set_tables(vm_ctx *ctx, const float *extra, int64_t n_extra,
                  const float *readgapcost, int64_t n_readgapcost,
                  const double *log2cache, int64_t n_log2cache);

/* Parameters of the chaining DPs (arguments of the reference njit functions). */
typedef struct vm_chain_params {
    int32_t kmersize;     /* 15 global (index k), 9 local */
    double skipcost;      /* golbal_skipcost / local_skipcost (vacmap:257-296) */
    int32_t maxdiff;      /* golbal_maxdiff 50 / local_maxdiff 30 */
    int32_t maxgap;       /* 1000 global (:23991); 99 H,S / 50 L local (:24061) */
    int32_t max_factor;   /* 1000 (:19367) opcount bail-out of the exact global DP */
    int32_t fast_t;       /* 5: bucket size above which the fast DP probes one member */
    int32_t large_readgap;/* 30 (:28587) multi-chain local DP */
    int32_t variant;      /* 0 global _d_all; 1 local _fine_list; 2 local _fine_list_mismatch */
} vm_chain_params;

/*
 * Global non-linear chaining of a batch of reads.
 * Replaces: hit2work_1's `np.argsort(one_mapinfo[:,0])` (:23572, numba quicksort
 * permutation) followed by get_optimal_chain_..._fine_list_d_all (:24828-25031)
 * and its fall-back get_optimal_chain_..._fine_list_d_fast_all (:25033-25339)
 * chosen exactly as hit2work_1 does (:23570-23579: n/read_len > 5, or the exact
 * DP bailed out).
 *
 *   anchors   int64[total][4], UNSORTED (as index.map() returned them)
 *   off       int64[n_reads+1]
 *   read_len  int32[n_reads]
 * outputs (caller-allocated, same raggedness):
 *   sorted    int64[total][4]  anchors after the argsort              (:23572)
 *   S         float64[total], P int32[total], S_arg int32[total]      (:24853-24885)
 *   g_max_index int64[n_reads]
 *   used_fast int32[n_reads]   1 when the fast DP produced the result
 *   kernel_ms (optional) device time of the kernels, CUDA events on the ctx stream
 */
int vm_chain_global_batch(vm_ctx *ctx, const vm_chain_params *prm, int64_t n_reads,
                          const int64_t *anchors, const int64_t *off, const int32_t *read_len,
                          int64_t *sorted, double *S, int32_t *P, int32_t *S_arg,
                          int64_t *g_max_index, int32_t *used_fast, float *kernel_ms);

/*
 * Device-resident three-step form of the same call, for measuring with inputs
 * already in HBM: upload once, run (timed) any number of times, download.
 */
int vm_chain_global_upload(vm_ctx *ctx, const vm_chain_params *prm, int64_t n_reads,
                           const int64_t *anchors, const int64_t *off, const int32_t *read_len);
int vm_chain_global_run(vm_ctx *ctx, float *kernel_ms);
int vm_chain_global_download(vm_ctx *ctx, int64_t *sorted, double *S, int32_t *P, int32_t *S_arg,
                             int64_t *g_max_index, int32_t *used_fast);
/* per-kernel device time of the last run (ms): [pack, sort, dp_exact, dp_fast] */
int vm_chain_global_times(vm_ctx *ctx, float *ms4);

#ifdef __cplusplus
}
#endif
#endif /* VACMAP_B200_H */
