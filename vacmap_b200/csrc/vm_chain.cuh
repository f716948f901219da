// Non-linear anchor chaining on sm_100a: kernels and launch plumbing.
//
// One warp owns one read.  The reference's DP
// (mammap_clrnano.py:24828-25031 global `_d_all`, 27305-27528 local `_fine_list`,
// 28250-28476 local `_fine_list_mismatch`) walks the predecessors of anchor i in
// DESCENDING score through the score-sorted index list S_arg and stops at the
// first one that can no longer win.  Here that walk is done 32 predecessors at a
// time: each lane scores one predecessor, an exclusive prefix-max over the lanes
// reproduces the reference's running `max_scores` at every step (so the break
// position and the `opcount` bail-out counter are exact), and an arg-max with
// lowest-lane tie-break reproduces "first strictly better predecessor wins".
// S (float64) and S_arg (int32) live in shared memory for reads up to
// VM_CHAIN_SMEM_CAP anchors and in the (L2-resident) output arrays beyond that.
// S_arg insertion replays the reference's binary search (`insertpoint_score`
// :19369-19387) from the two counts (#S < t, #S <= t) found with a warp-wide
// 32-ary search, then shifts the tail of S_arg one slot with all lanes.
#pragma once
#include "vm_common.cuh"

#define VM_CHAIN_SMEM_CAP 16384   // anchors; 12 B each -> 192 KB of the 227 KB
#define VM_CHAIN_ANCHOR_SMEM_MIN 2048   // capacity classes from here up to ..._MAX stage the anchors in shared memory as well
#define VM_CHAIN_ANCHOR_SMEM_MAX 6144   // 6144 * 28 B = 168 KB
#define VM_GCL_MAX 64             // gapcost_list entries kept in smem (maxdiff+1 <= 64)
#define VM_RGL_MAX 128            // read-gap cost entries kept in smem (maxgap+1 <= 128)

struct VmChainArgs {
    const VmAnchor *anchors;   // sorted, concatenated
    const int64_t *off;        // [n_reads] start of each read's anchors
    const int32_t *cnt;        // [n_reads] number of anchors
    double *S;                 // outputs, concatenated like anchors
    int32_t *P;
    int32_t *S_arg;
    int64_t *gmax;             // [n_reads] g_max_index, or -1 / -2 on bail-out
    int64_t *opcount;          // [n_reads]
    const float *extra;        // device copies of the host-built tables
    long long extra_size;      // len - 1
    const double *log2cache;
    long long log2cache_size;  // len - 1
    const double *gapcost_list;// device, maxdiff + 1 entries
    const float *rgcost;       // variant 1: readgapcost_list[100]; variant 2: large_readgapcost_list[maxgap+1]
    int n_rg;
    double skipcost;
    int maxdiff;
    int maxgap;
    int max_factor;
    // variant 3 (asm mode, linked global DP): per read the number of carried anchors in front of the batch (their S / P
    // already in S / P), and head[3 r .. 3 r + 3) = carried g_max_scores, g_max_index, prereadloc (mammap_asm.py:21713)
    const int32_t *pre_n;
    const double *head;
};

int vm_launch_chain_exact(int variant, const VmChainArgs &args, const int *read_ids_dev, int n_ids,
                          int cap, bool use_smem, cudaStream_t stream);
int vm_launch_chain_fast(int variant, const VmChainArgs &args, int fast_t, const int *read_ids_dev,
                         int n_ids, long long *scratch_i64, const int64_t *scratch_off,
                         cudaStream_t stream);
int vm_launch_pack(const int64_t *rows_dev, VmAnchor *out, long long total, cudaStream_t stream);
#define VM_SORT_SMEM_CAP 8192      // anchors; 16 B each (keys, perm, two stop lists) -> 128 KB
int vm_launch_sort_anchors(const VmAnchor *in, const int64_t *off, const int32_t *cnt, const int *read_ids_dev, int n_ids, int cap,
                           bool use_smem, int key_is_end, int32_t *perm, int32_t *gscratch, VmAnchor *sorted,
                           int64_t *sorted_rows, cudaStream_t stream);
