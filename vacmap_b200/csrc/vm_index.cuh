// Reference index: minimizer hash table (global seeding) and 9-mer position index (local
// re-seeding), built on the host, resident in HBM.
//
// Replaces what the reference gets from `vacmap_index.Aligner(path, w, k)` (vacmap:344; a
// minimap2 .mmi built by `minimap2 -d`, vacmap:329-336) and, for the local stage, the per-read
// numba Dict of all 9-mers of the guide windows (mammap_clrnano.py:23073-23087, 23138-23140):
// on a 180 GB part the 9-mer positions of the WHOLE reference fit in HBM (4 B per base), so the
// per-read table build disappears and a window lookup becomes a range query in a sorted list.
#pragma once
#include "vm_common.cuh"
#include <string>
#include <vector>

#define VM_HT_EMPTY 0xffffffffffffffffULL
#define VM_K9 9
#define VM_K9_KEYS 1953125   // 5^9

struct VmHtSlot {
    uint64_t key;     // minimizer hash (VM_HT_EMPTY = free)
    uint32_t start;   // offset into occ
    uint32_t count;
};

struct VmIndexDev {           // device pointers, passed to kernels by value
    const VmHtSlot *ht;
    uint64_t ht_mask;
    const uint64_t *occ;      // global last-base position << 1 | strand, ascending per key
    const uint32_t *kpos;     // 9-mer start positions (global), sorted by (code, position)
    const int64_t *koff;      // VM_K9_KEYS + 1 offsets into kpos
    const uint8_t *ref;       // concatenated upper-case reference
    int64_t ref_len;
    int32_t w, k;
    int32_t mid_occ;
    // position directory of the long 9-mer runs (vm_index_build_buckets): krow[code] = row (or -1: short run, plain
    // binary search); kbk[row * (kb_n + 1) + b] = lower_bound(b << kb_shift) inside the code's run, b = 0 .. kb_n
    const int32_t *krow;
    const uint32_t *kbk;
    int32_t kb_shift, kb_n;
};

// first index in [b, e) of the code's position run with kpos >= lo (the window's start): a plain binary search over the
// whole run costs log2(run) dependent HBM round trips -- 14 on a GRCh38-sized reference, where a 9-mer occurs ~12 000
// times; with the directory it is one directory read and a search inside one position bucket (a sector or two)
__device__ __forceinline__ int64_t vm_kpos_lower_bound(const VmIndexDev &ix, int code, int64_t b, int64_t e, long long lo)
{
    int64_t l = b, h = e;
#ifdef __CUDA_ARCH__
    const int row = ix.krow ? ix.krow[code] : -1;
    if (row >= 0) {
        long long bk = lo <= 0 ? 0 : (lo >> ix.kb_shift);
        if (bk > ix.kb_n) bk = ix.kb_n;
        const uint32_t *t = ix.kbk + (size_t)row * (size_t)(ix.kb_n + 1);
        l = b + t[bk];
        h = bk < ix.kb_n ? b + t[bk + 1] : e;
    }
#endif
    while (l < h) {
        const int64_t mid = (l + h) >> 1;
        if ((long long)ix.kpos[mid] < lo) l = mid + 1; else h = mid;
    }
    return l;
}

struct VmIndex {
    int w = 10, k = 15;
    int mid_occ_default = 10;
    std::vector<std::string> names;
    std::vector<int64_t> ctg_start, ctg_len;
    std::string ref;                       // host copy (upper-case, non-ACGT -> N)
    std::vector<VmHtSlot> ht;
    std::vector<uint64_t> occ;
    std::vector<uint32_t> kpos;
    std::vector<int64_t> koff;
    int64_t n_keys = 0;
    int64_t n_occ = 0, n_kpos = 0;         // minimizer occurrences, 9-mer positions
    uint64_t ht_slots = 0;
    // device copies
    void *d_ht = nullptr, *d_occ = nullptr, *d_kpos = nullptr, *d_koff = nullptr, *d_ref = nullptr;
    void *d_ukeys = nullptr, *d_ucnt = nullptr;      // device build: unique minimizer hashes (ascending) and their counts
    void *d_krow = nullptr, *d_kbk = nullptr;        // position directory of the long 9-mer runs (always owned by the index)
    bool borrowed = false;                 // the device arrays belong to the caller (vm_index_adopt)
    VmIndexDev dev{};
};

// base -> 0..3 (ACGT/U, either case), 4 otherwise (minimap2 seq_nt4_table)
__host__ __device__ __forceinline__ int vm_nt4(unsigned char c)
{
#ifdef __CUDA_ARCH__
    // branch-free on the device: letters A C G T U (either case) by a bit mask, code from bits 1..2 of the ASCII code
    const unsigned u = (c & 0xDFu) - 'A';
    const unsigned x = ((c & 0xDFu) >> 1) & 3u;          // A 0, C 1, T/U 2, G 3
    const bool ok = u < 26u && ((0x00180045u >> u) & 1u);
    return ok ? (int)(x ^ (x >> 1)) : 4;
#endif
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': case 'U': case 'u': return 3;
    default: return 4;
    }
}

// minimap2 sketch.c hash64 (invertible integer hash)
__host__ __device__ __forceinline__ uint64_t vm_hash64(uint64_t key, uint64_t mask)
{
    key = (~key + (key << 21)) & mask;
    key = key ^ key >> 24;
    key = ((key + (key << 3)) + (key << 8)) & mask;
    key = key ^ key >> 14;
    key = ((key + (key << 2)) + (key << 4)) & mask;
    key = key ^ key >> 28;
    key = (key + (key << 31)) & mask;
    return key;
}

__host__ __device__ __forceinline__ uint64_t vm_ht_hash(uint64_t key)
{
    key *= 0x9E3779B97F4A7C15ULL;
    return key ^ (key >> 29);
}

// (w,k)-minimizer sketch of one sequence, minimap2 2.29 `mm_sketch` (non-HPC) semantics:
// every k-mer tying the window minimum is emitted, symmetric k-mers are skipped without
// occupying a window slot, an ambiguous base resets the run length but not the k-mer registers.
// emit(hash, last_base_pos << 1 | strand) is called in position order.  One caller = one
// sequential state machine (host: per contig; device: one thread per read).
// Range form: the machine starts FRESH at `begin` (i.e. as if the sequence started there) and runs
// to `end`; positions are reported in the coordinates of `str`.  observe(i, kind) is called for every
// base (kind 0 = counted k-mer position, 1 = ambiguous base, 2 = symmetric k-mer) and may return false
// to stop; `flush` = emit the pending minimum when the loop ran to `end` (minimap2 does at the end of
// the sequence).  Used with a warm-up by the chunk-parallel device sketch (vm_seed.cu).
template <typename Emit, typename Observe>
__host__ __device__ inline void vm_sketch_range(const unsigned char *str, int64_t begin, int64_t end, int w, int k, bool flush,
                                                Emit emit, Observe observe)
{
    const uint64_t shift1 = 2 * (uint64_t)(k - 1), mask = (1ULL << 2 * k) - 1;
    uint64_t kmer0 = 0, kmer1 = 0;
    uint64_t bufx[256], bufy[256];
    uint64_t minx = ~0ULL, miny = ~0ULL;
    int l = 0, buf_pos = 0, min_pos = 0, kmer_span = 0;
    for (int j = 0; j < w; ++j) { bufx[j] = ~0ULL; bufy[j] = ~0ULL; }
    bool stopped = false;
    for (int64_t i = begin; i < end; ++i) {
        const int c = vm_nt4(str[i]);
        uint64_t infox = ~0ULL, infoy = ~0ULL;
        if (c < 4) {
            kmer_span = l + 1 < k ? l + 1 : k;
            kmer0 = (kmer0 << 2 | (uint64_t)c) & mask;
            kmer1 = (kmer1 >> 2) | (3ULL ^ (uint64_t)c) << shift1;
            if (kmer0 == kmer1) {
                if (!observe(i, 2)) { stopped = true; break; }
                continue;
            }
            const int z = kmer0 < kmer1 ? 0 : 1;
            ++l;
            if (l >= k && kmer_span < 256) {
                infox = vm_hash64(z ? kmer1 : kmer0, mask) << 8 | (uint64_t)kmer_span;
                infoy = (uint64_t)i << 1 | (uint64_t)z;
            }
        } else { l = 0; kmer_span = 0; }
        bufx[buf_pos] = infox;
        bufy[buf_pos] = infoy;
        if (l == w + k - 1 && minx != ~0ULL) {
            for (int j = buf_pos + 1; j < w; ++j)
                if (minx == bufx[j] && bufy[j] != miny) emit(bufx[j] >> 8, bufy[j]);
            for (int j = 0; j < buf_pos; ++j)
                if (minx == bufx[j] && bufy[j] != miny) emit(bufx[j] >> 8, bufy[j]);
        }
        if (infox <= minx) {
            if (l >= w + k && minx != ~0ULL) emit(minx >> 8, miny);
            minx = infox; miny = infoy; min_pos = buf_pos;
        } else if (buf_pos == min_pos) {
            if (l >= w + k - 1 && minx != ~0ULL) emit(minx >> 8, miny);
            minx = ~0ULL;
            for (int j = buf_pos + 1; j < w; ++j)
                if (minx >= bufx[j]) { minx = bufx[j]; miny = bufy[j]; min_pos = j; }
            for (int j = 0; j <= buf_pos; ++j)
                if (minx >= bufx[j]) { minx = bufx[j]; miny = bufy[j]; min_pos = j; }
            if (l >= w + k - 1 && minx != ~0ULL) {
                for (int j = buf_pos + 1; j < w; ++j)
                    if (minx == bufx[j] && miny != bufy[j]) emit(bufx[j] >> 8, bufy[j]);
                for (int j = 0; j <= buf_pos; ++j)
                    if (minx == bufx[j] && miny != bufy[j]) emit(bufx[j] >> 8, bufy[j]);
            }
        }
        if (++buf_pos == w) buf_pos = 0;
        if (!observe(i, c < 4 ? 0 : 1)) { stopped = true; break; }
    }
    if (flush && !stopped && minx != ~0ULL) emit(minx >> 8, miny);
}

template <typename Emit>
__host__ __device__ inline void vm_sketch(const unsigned char *str, int64_t len, int w, int k, Emit emit)
{
    vm_sketch_range(str, 0, len, w, k, true, emit, [](int64_t, int) { return true; });
}

// 9-mer code over the 5-letter alphabet ACGTN (anything else reads as N)
__host__ __device__ __forceinline__ int vm_code5(unsigned char c) { return vm_nt4(c); }

VmIndex *vm_index_build_host(const std::vector<std::string> &names, const std::vector<std::string> &seqs, int w, int k);
// the same index built on the device (vm_index_gpu.cu); ix->ref holds the raw concatenated sequence on entry
int vm_index_build_device(VmIndex *ix, std::string &err);
int vm_index_upload(VmIndex *ix, std::string &err);
// position directory over ix->dev.kpos / koff (device arrays already in place); fills ix->dev.krow / kbk / kb_shift / kb_n
int vm_index_build_buckets(VmIndex *ix, std::string &err);
void vm_index_free(VmIndex *ix);
