// Base-level kernels: global edit distance, z-drop edge extension, global dual-affine fill.
// Job descriptors shared between the host backend and the kernels (see vm_align.cu).
#pragma once
#include "vm_common.cuh"

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
    int64_t out_off;   // fill: offset of this job's CIGAR ops; edit distance / extension: unused
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

int vm_launch_edit_distance(VmAlnJobDev *jobs, const int *ids_dev, const int *class_start, const int *class_words, VmSeqSources src,
                            cudaStream_t stream);
int vm_launch_extend(VmAlnJobDev *jobs, int n_jobs, VmSeqSources src, cudaStream_t stream);
// dir: direction bytes, vm_fill_dir_bytes(tlen, qlen) per job at dir_off (8-byte aligned);
// band_scratch: 3 * qlen ints per job whose target exceeds vm_fill_band_rows() rows (sc_off, else -1)
size_t vm_fill_dir_bytes(int tlen, int qlen);
int vm_fill_band_rows();
int vm_launch_fill(VmAlnJobDev *jobs, int n_jobs, VmSeqSources src, int eqx, uint8_t *dir, int32_t *band_scratch,
                   uint32_t *cigar_out, cudaStream_t stream);
