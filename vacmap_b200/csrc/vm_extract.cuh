// Device-side chain extraction (see vm_extract.cu): what the host glue needs of a chaining stage, compacted.
#pragma once
#include "vm_common.cuh"

// per read: where its extracted anchors / chain records sit in the dense output arrays
struct VmExtractRec {
    long long anc_off;
    long long meta_off;
    int32_t n_anc;
    int32_t n_chains;     // global stage: chains kept (0 = read not accepted); local stage: 1 when a path exists
};

struct VmExtractOut {
    VmExtractRec *rec;                    // [n_reads]
    VmAnchor *anc;                        // dense; global: chains back to back, each in DESCENDING read order;
                                          //        local: the trimmed best chain in ASCENDING read order
    double *S;                            // global: S of every anchor of the primary chain (0 for the others)
    int32_t *chain_len;                   // global: anchors per chain, discovery order (primary first)
    double *chain_score;                  // global: score per chain; local: [n_reads] score of the best chain (or null)
    unsigned long long *n_anc_total;      // bump allocators, zeroed by the host
    unsigned long long *n_chain_total;
};

int vm_launch_extract_global(const int *ids_dev, int n_ids, const int64_t *off, const int32_t *cnt, const VmAnchor *sorted,
                             const double *S, const int32_t *P, const int32_t *S_arg, const int64_t *gmax, double accept,
                             uint8_t *used_zeroed, VmAnchor *tmp_anc, double *tmp_S, int32_t *tmp_len, double *tmp_score,
                             const VmExtractOut &out, cudaStream_t stream);
int vm_launch_extract_local(const int *ids_dev, int n_ids, const int64_t *off, const int32_t *cnt, const VmAnchor *sorted,
                            const double *S, const int32_t *P, const int64_t *gmax, VmAnchor *tmp_anc, const VmExtractOut &out,
                            cudaStream_t stream);

// rebuild_chain_break on the device: per read, the colinear sub-alignments of its local path
struct VmRebuildRec {
    long long anc_off;    // into VmRebuildOut::anc: the anchors of all its sub-alignments, back to back
    long long len_off;    // into VmRebuildOut::len: anchors per sub-alignment
    int32_t n_anc;
    int32_t n_al;         // 0: the reference raises here (no sub-alignment left) -> the read has no records
};
struct VmRebuildOut {
    VmRebuildRec *rec;                    // [n_reads]
    VmAnchor *anc;
    int32_t *len;
    unsigned long long *n_anc_total;      // bump allocators, zeroed by the host
    unsigned long long *n_len_total;
};
int vm_launch_rebuild(const int *ids_dev, int n_ids, const int64_t *off, const VmExtractRec *rec, const VmAnchor *path,
                      const int64_t *ctg_start_dev, int n_ctg, int large_cost, int small_alignment, VmAnchor *tmp_anc, int32_t *tmp_len,
                      const VmRebuildOut &out, cudaStream_t stream);
