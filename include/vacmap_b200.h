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
/* Page-locked host memory for the read batches handed to vm_align_submit: the host-to-device copy of a sub-batch is
 * then a DMA the worker's stream overlaps with the kernels of the other sub-batches; from pageable memory the driver
 * stages it through the calling thread at memcpy speed.  vm_host_alloc / vm_host_free own the buffer (vm_host_free accepts ctx == NULL);
 * vm_host_register / vm_host_unregister lock a buffer the caller owns (e.g. the read parser's).  No counterpart in the
 * reference (its reads stay in Python strings, vacmap:391-420). */
int vm_host_alloc(vm_ctx *ctx, int64_t bytes, void **out);
int vm_host_free(vm_ctx *ctx, void *p);
int vm_host_register(vm_ctx *ctx, void *p, int64_t bytes);
int vm_host_unregister(vm_ctx *ctx, void *p);

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
    int32_t variant;      /* 0 global _d_all; 1 local _fine_list; 2 local _fine_list_mismatch; 3 / 4 asm linked _d_all / _all (vm_chain_linked_batch) */
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

/*
 * asm mode: the linked global DP with carry-in, one job per (contig, anchor batch).
 * Replaces linked_get_optimal_chain_..._fine_list_d_all(g_max_scores, g_max_index, pre_S, pre_P, prereadloc,
 * one_mapinfo, kmersize, skipcost, maxdiff, maxgap) (mammap_asm.py:21687-21871) as called from
 * assembly_get_readmap_DP_test (:23244).
 *   anchors  int64[total][4]: per job the pre_n[j] anchors carried over from the previous batch followed by this
 *            batch's anchors sorted by read position (`linked_one_mapinfo`, :23232) -- already in DP order
 *   off      int64[n_jobs+1]; pre_n int32[n_jobs] (0: plain start, :21720-21729)
 *   head     float64[n_jobs][3] = g_max_scores, g_max_index, prereadloc of the call (ignored when pre_n == 0)
 *   S, P     IN: the first pre_n[j] entries of every job hold pre_S / pre_P; OUT: all of S, P (int32, -9999999 = start)
 *   S_arg    OUT int32[total]; g_max_index OUT int64[n_jobs]
 * prm->variant == 4 selects the second-round twin linked_..._fine_list_all (:21505-21686, called at :23343): same
 * arguments and carry, asm's read-gap cost on colinear pairs, no bail-out; any other value the first-round DP.
 *   used_fast OUT int32[n_jobs] (optional): 1 where the exact DP bailed out on opcount (:21754) and the result is the
 *            heuristic twin's, linked_..._fine_list_d_fast_all (:21872-22158) on the same arguments -- the caller's
 *            fall-back at :23246-23247 (S_arg is then its S_arg_i)
 */
int vm_chain_linked_batch(vm_ctx *ctx, const vm_chain_params *prm, int64_t n_jobs, const int64_t *anchors, const int64_t *off,
                          const int32_t *pre_n, const double *head, double *S, int32_t *P, int32_t *S_arg,
                          int64_t *g_max_index, int32_t *used_fast);

/*
 * Stage-level local chaining (parity tests).  Replaces get_optimal_chain_..._fine_list (variant 1,
 * clrnano:27305-27528), _fine_list_mismatch (variant 2, :28250-28476) and, with force_fast, their _fast twins
 * (:26938-27303, :27891-28248); without it the exact DP falls back to _fast as the reference does (:27380-27384).
 * presorted != 0: `anchors` are already ordered by read end (what the functions take); else they are sorted like
 * np.argsort(x + len) at :28585.  Outputs per read: score[r] = g_max_scores, the chain in path[path_off[r] ..
 * path_off[r+1]) as int64 rows, trimmed like :27508-27527, in ASCENDING read order (the reference's list reversed);
 * `path` needs room for off[n_reads] rows.
 */
int vm_chain_local_batch(vm_ctx *ctx, const vm_chain_params *prm, int32_t presorted, int32_t force_fast, int64_t n_reads,
                         const int64_t *anchors, const int64_t *off, const int32_t *read_len, double *score, int64_t *path,
                         int64_t *path_off, int32_t *used_fast);

/* ------------------------------------------------------------------------
 * Reference index.  Replaces `vacmap_index.Aligner(path, w=, k=)` (vacmap:344; a minimap2
 * .mmi built by `minimap2 -d`, vacmap:329-336) and its accessors `.k`, `.seq_offset`
 * (vacmap:358-361, clrnano:24098-24102) and `.seq(name)` (vacmap:363).  Sequences are
 * upper-cased and every non-ACGT base becomes N (what mappy's .seq() hands back).
 * Holds, in HBM: the reference (1 B/base), the (w,k)-minimizer hash table, and the
 * positions of every 9-mer of the reference for the local re-seeding stage.
 * ---------------------------------------------------------------------- */
typedef struct vm_index_handle vm_index_handle;

int vm_index_create(vm_ctx *ctx, int32_t n_contigs, const char *const *names, const char *const *seqs,
                    const int64_t *lens, int32_t w, int32_t k, vm_index_handle **out);
void vm_index_destroy(vm_index_handle *h);
/* The index is built ON THE DEVICE (what `minimap2 -d` does for the reference, vacmap:324-344, plus the 9-mer position
 * index of the local stage): chunk-parallel mm_sketch of the contigs, stable radix sort by hash, hash table by atomicCAS,
 * 9-mer codes radix-sorted.  The three entry points below let one process build it and the others use it:
 *   vm_index_arrays     the five device arrays of a built index (reference, hash table, occurrences, 9-mer positions,
 *                       9-mer offsets: ptrs[5], bytes[5]) and eight scalars (meta[8]: n_keys, n_occ, n_kpos, ht_slots,
 *                       mid_occ, w, k, reference length)
 *   vm_index_adopt      an index over device arrays the CALLER owns (e.g. filled by an NCCL broadcast); they must outlive
 *                       the handle
 *   vm_index_minimizers the distinct minimizer hashes (ascending), their counts and their occurrences on the host -- what
 *                       a minimap2 `.mmi` file stores (vacmap_b200/mmi.py writes it) */
int vm_index_arrays(vm_index_handle *h, const void **ptrs, int64_t *bytes, int64_t *meta);
int vm_index_adopt(vm_ctx *ctx, int32_t n_contigs, const char *const *names, const int64_t *lens, const void *const *ptrs,
                   const int64_t *bytes, const int64_t *meta, vm_index_handle **out);
int vm_index_minimizers(vm_index_handle *h, uint64_t *keys, int32_t *counts, uint64_t *occ);
/* np.argsort as numba compiles it (numba/misc/quicksort.py: median-of-three quicksort, insertion sort below 15 elements)
 * on int64 keys: order[n].  Host function; the Python mirrors of the reference's asm-mode code use it where the reference
 * sorts inside njit functions (mammap_asm.py:22754-22755). */
int vm_argsort_i64(const int64_t *keys, int64_t n, int64_t *order);
int vm_index_info(vm_index_handle *h, int32_t *k, int32_t *w, int32_t *n_contigs, int64_t *n_minimizers,
                  int64_t *n_keys, int32_t *mid_occ);
/* contig i: name, global start offset (`seq_offset[i][2]`), length, pointer to its (library-owned) sequence */
int vm_index_contig(vm_index_handle *h, int32_t i, const char **name, int64_t *start, int64_t *len, const char **seq);

/* ------------------------------------------------------------------------
 * Per-read alignment of a batch.  Replaces get_readmap_DP_test (clrnano:24023-24084) as
 * called by the worker loop get_list_of_readmap_stdout (clrnano:24117): seeding ->
 * global non-linear chaining -> local re-seeding + chaining -> extension -> records.
 * Parameters are the `option` dict keys the live path reads (vacmap:257-296) plus the
 * per-mode constants that differ between mammap_clrnano / _ccs / _sensitive.
 * ---------------------------------------------------------------------- */
typedef struct vm_align_params {
    double global_skipcost;   /* option['golbal_skipcost'] */
    double local_skipcost;    /* option['local_skipcost'] */
    double maxdivergence;     /* option['maxdivergence'] */
    double accept_score;      /* primary-chain threshold: 60 (H) / 40 (L, S)  (clrnano:23650) */
    int32_t global_maxdiff;   /* option['golbal_maxdiff'] */
    int32_t local_maxdiff;    /* option['local_maxdiff'] */
    int32_t check_num;        /* option['c'] "Top N clusters" */
    int32_t eqx;              /* option['eqx'] */
    int32_t hardclip;         /* option['H'] */
    int32_t nodiscard;        /* option['nodiscard'] */
    int32_t max_guides;       /* 5 (H) / 3 (L) / 0 = unlimited (S)  (clrnano:28581) */
    int32_t local_maxgap;     /* 99 (H, S) / 50 (L)  (clrnano:24061) */
    int32_t clamp40;          /* 1 in mode L: min(skipcost, 40) in the multi-chain local DP */
    int32_t host_threads;     /* host glue threads, 0 = all cores */
    int32_t workers;          /* size of the worker pool (own CUDA stream each), 0 = default (6), 1 = lock-step */
    int32_t chunk_reads;      /* reads per sub-batch, 0 = half of the batch (at least 512) */
} vm_align_params;

/* One alignment record = one row of `onemapinfolist` (clrnano:20760):
 * (readid, contig, strand, q_st, q_en, r_st, r_en, mapq, cigar). */
typedef struct vm_record {
    int32_t contig;      /* index into the contig table */
    int32_t strand;      /* +1 / -1 */
    int64_t q_st, q_en;  /* query span */
    int64_t r_st, r_en;  /* reference span, contig coordinates */
    int32_t mapq;
    int32_t cigar_len;   /* number of ops */
    int64_t cigar_off;   /* into the result's CIGAR arena */
} vm_record;

typedef struct vm_result vm_result;

/* seqs: upper-case bases of all reads concatenated; seq_off[n_reads+1].  The result arena is
 * library-allocated and freed with vm_result_free.  Reads that do not map, or that the
 * reference would drop through an exception (clrnano:24116-24125), simply have no records. */
int vm_align_batch(vm_ctx *ctx, vm_index_handle *index, const vm_align_params *prm, int64_t n_reads,
                   const char *seqs, const int64_t *seq_off, vm_result **out);
/* Asynchronous form: vm_align_submit queues the batch (cut into sub-batches for the context's worker pool, each
 * worker with its own CUDA stream) and returns at once; vm_align_wait blocks until it is done and hands back the
 * result (or the error).  A batch submitted while the previous one is still draining keeps the device busy across
 * batch boundaries.  seqs / seq_off must stay valid until vm_align_wait returns; resident != 0: the reads were put
 * in HBM by vm_reads_upload and must not be replaced while jobs that use them are in flight.
 * vm_align_batch == submit + wait.  The reference has no counterpart: its workers take one read at a time
 * (clrnano:24110-24117); this is the batch seam of the B200 design. */
typedef struct vm_job vm_job;
int vm_align_submit(vm_ctx *ctx, vm_index_handle *index, const vm_align_params *prm, int64_t n_reads, const char *seqs,
                    const int64_t *seq_off, int32_t resident, vm_job **out);
int vm_align_wait(vm_job *job, vm_result **out);
/* Two-step form for measuring with the reads already resident in HBM: vm_reads_upload copies the
 * batch (forward + reverse-complement strands) to the device, vm_align_resident then runs the
 * same pipeline on it (same arguments; seqs/seq_off are still needed by the host glue). */
int vm_reads_upload(vm_ctx *ctx, vm_index_handle *index, int64_t n_reads, const char *seqs, const int64_t *seq_off);
int vm_align_resident(vm_ctx *ctx, vm_index_handle *index, const vm_align_params *prm, int64_t n_reads,
                      const char *seqs, const int64_t *seq_off, vm_result **out);
/* Stage-level seeding: per read, what `index_object.map(seq, check_num, mid_occ=-1)` (clrnano:23985)
 * followed by get_reversed_chain_numpy_rough (clrnano:21202-21217) yields: int64 rows
 * (readpos, refpos_global, strand, len) in rows[row_off[r] .. row_off[r+1]) and the flip flag.
 * Returns VM_ERR_NOMEM (row_off filled) when `cap` rows are not enough. */
int vm_seed_batch_rows(vm_ctx *ctx, vm_index_handle *index, int32_t check_num, int64_t n_reads, const char *seqs,
                       const int64_t *seq_off, int64_t *rows, int64_t cap, int64_t *row_off, int32_t *need_reverse);

/* Stage-level local re-seeding: the 9-mer scan + same-diagonal merge of get_localmap_..._guide_1 (clrnano:23138-23344)
 * for guide jobs built by the caller (window construction :23095-23136 is host glue).  Job j: read job_read[j], given
 * already oriented (its `testseq`); read positions [readstart[j], readend[j]); reference windows
 * win_lo / win_hi[win_off[j] .. win_off[j+1]) as GLOBAL [lo, hi) in insertion order; guide points gx / gy[g_off[j] ..
 * g_off[j+1]) sorted by read position.  Output: the anchors, in the reference's emission order, as int64 rows
 * (readpos, refpos_global, strand, len) in rows[row_off[j] .. row_off[j+1]); VM_ERR_NOMEM (row_off filled) when `cap`
 * rows are not enough. */
int vm_local_reseed_batch(vm_ctx *ctx, vm_index_handle *index, int64_t n_reads, const char *seqs, const int64_t *seq_off,
                          int64_t n_jobs, const int32_t *job_read, const int32_t *readstart, const int32_t *readend,
                          const int64_t *win_off, const int64_t *win_lo, const int64_t *win_hi, const int64_t *g_off,
                          const int32_t *gx, const int64_t *gy, int64_t *rows, int64_t cap, int64_t *row_off);

/* Stage-level entry point of the base-level kernels on raw sequence pairs (parity tests).
 * kind 0: edlib.align(query, target, task='distance') (clrnano:19251)        -> out0[j] = distance
 * kind 3: as kind 0, computed inside the band |i - j| <= out1[j] (out1 is an INPUT): out0[j] = the exact distance
 *         when it is <= out1[j], else some value > out1[j] -- all the divergence filter (clrnano:19252) needs
 * kind 1: mp.k_cigar(t, q, 2,-4, 4,4,4,4, bw=100, zdropvalue=50) (clrnano:2381) -> out0 = q_e, out1 = t_e
 * kind 2: mp.k_cigar(t, q, 2,-4, 4,2,24,1, bw=-1, zdropvalue=-1, eqx) (clrnano:21554) -> out0[j] = number of
 *         CIGAR ops, written at cigar[cig_off[j]...], cig_off[j] = sum_{i<j} (tlen_i + qlen_i + 2). */
int vm_pairs_batch(vm_ctx *ctx, int32_t kind, int32_t eqx, int64_t n_pairs, const char *targets, const int64_t *t_off,
                   const char *queries, const int64_t *q_off, int64_t *out0, int64_t *out1, uint32_t *cigar);
/* ------------------------------------------------------------------------
 * SAM text of a batch's records on the library's host threads (no device needed): what the reference's emitter
 * get_bam_dict_str (clrnano:20841-21021; reassign_mapq :11661, mergecigar_ :4773, get_MD_CSshort/long :19012-19112,
 * P_alignmentstring :5391, output_functions.nm_from_cigar :300) writes for every read's onemapinfolist, byte for byte --
 * FLAG / primary by longest query span, SA, NM, MD, cs, CG, --H / --fakecigar, the tags of a FASTQ comment
 * (get_bam_dict_str_comments :21022).  A read whose NM / MD walk runs off a sequence emits nothing (the reference raises
 * and its worker swallows the read).  Inputs: the arrays of a vm_result, the reads (upper-case) with names and optional
 * qualities / comments (packed, [n_reads+1] offsets; NULL or an empty slice = absent), the contigs as the index gives them
 * (vm_index_contig).  Output: one buffer, the lines of read r at [offsets[r], offsets[r+1]).
 * ---------------------------------------------------------------------- */
typedef struct vm_sam_options {
    int32_t md;                /* --MD (MD and cs tags) */
    int32_t shortcs;           /* cs in its short form unless --cs=long */
    int32_t cigar2cg;          /* --L: CIGARs of more than 65 535 operations move to the CG tag */
    int32_t markunbalancetra;  /* --markunbalancetra: reassign_mapq */
    int32_t hardclip;          /* --H */
    int32_t fakecigar;         /* --fakecigar: short CIGARs in the SA tag */
    int32_t copycomments;      /* --copycomments */
    int32_t asm_mode;          /* -mode asm's emitter (iterator_get_bam_dict_str, mammap_asm.py:22757-22941): NM from the CIGAR alone,
                                  its primary-record rule, MAPQ written as 60 / 1 */
    const char *rg_id;         /* RG:Z tag of every record (NULL: none) */
} vm_sam_options;
typedef struct vm_text vm_text;
int vm_sam_batch(const vm_sam_options *opt, int32_t n_contigs, const char *const *contig_names, const char *const *contig_seqs,
                 const int64_t *contig_lens, int64_t n_reads, const int64_t *rec_off, const vm_record *recs, const uint32_t *cigar,
                 const char *seqs, const int64_t *seq_off, const char *names, const int64_t *name_off, const char *quals,
                 const int64_t *qual_off, const char *comments, const int64_t *comment_off, int32_t threads, vm_text **out);
const char *vm_text_data(vm_text *t);
int64_t vm_text_size(vm_text *t);
const int64_t *vm_text_offsets(vm_text *t);   /* [n_reads+1] */
void vm_text_free(vm_text *t);

int64_t vm_result_num_records(vm_result *r);
int64_t vm_result_num_cigar_ops(vm_result *r);
const int64_t *vm_result_read_offsets(vm_result *r);   /* [n_reads+1] into the record array */
const vm_record *vm_result_records(vm_result *r);
const uint32_t *vm_result_cigar(vm_result *r);         /* BAM encoding len<<4|op, ops MIDNSHP=X */
const char *vm_result_stage_times(vm_result *r);       /* "stage=ms;..." wall milliseconds per stage, "n_*" work counts, "c_*" branch counts */
/* Per-read status [n_reads]: why a read has the records it has.  The reference emits nothing for a read in several
 * distinguishable situations (SURVEY 8b "error convention"); they are told apart here instead of being merged into
 * "zero records":
 *   VM_READ_OK          >= 1 record
 *   VM_READ_FEW_ANCHORS <= 2 seed anchors (decode_hit, clrnano:23986-23988)
 *   VM_READ_LOW_SCORE   best global chain not above the mode's threshold (hit2work_1, clrnano:23650, 23711-23734)
 *   VM_READ_SHORT_LOCAL local chain of <= 1 anchor (clrnano:24067-24068)
 *   VM_READ_DROPPED     the reference raises inside the read and its worker swallows the exception
 *                       (clrnano:24116-24125): "Failed to compute CIGAR" (21559-21569), Cigar length check
 *                       (20779-20786), empty sub-alignment list, division by zero in the divergence filter
 *   VM_READ_NO_RECORDS  extend_func produced no record (clrnano:24075-24076)
 *   VM_READ_FAILED      the library could not process the read (beyond a kernel's size limits); no counterpart */
#define VM_READ_OK 0
#define VM_READ_FEW_ANCHORS 1
#define VM_READ_LOW_SCORE 2
#define VM_READ_SHORT_LOCAL 3
#define VM_READ_DROPPED 4
#define VM_READ_NO_RECORDS 5
#define VM_READ_FAILED 6
const int32_t *vm_result_read_status(vm_result *r);
void vm_result_free(vm_result *r);

#ifdef __cplusplus
}
#endif
#endif /* VACMAP_B200_H */
