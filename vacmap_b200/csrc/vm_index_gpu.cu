// Construction of the reference index on the device: what `minimap2 -d` does for the reference
// (vacmap:324-344) plus the 9-mer position index of the local stage, built where they will live.
//
//   minimizers   the contigs are sketched by the same exact chunk-parallel mm_sketch kernel that sketches reads
//                (vm_seed.cu), the per-chunk outputs are compacted by a prefix sum, sorted by hash with a STABLE radix
//                sort (positions arrive ascending, so every key's occurrence run stays position-sorted, as the host
//                build's (hash, position) sort leaves it), run-length encoded into keys / starts / counts and inserted
//                into the open-addressing table with atomicCAS (lookups do not depend on the insertion order);
//   mid_occ      minimap2's default occurrence cap: the (1 - 2e-4) quantile of the per-key counts, by sorting them;
//   9-mers       every position's 5-letter code, stable radix sort by code (21 bits), offsets from the run boundaries.
//
// The sorts / scans / run-length encoding are CUB device primitives (library plumbing of the one-off build, not part
// of the per-read path); everything else is this file's kernels.  Host copies are kept only of what the host glue
// reads (the normalised sequence).
#include "vm_index.cuh"
#include "vm_seed.cuh"
#include <cub/cub.cuh>

#define IX_OK(call)                                                                               \
    do {                                                                                          \
        cudaError_t _e = (call);                                                                  \
        if (_e != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(_e); return -1; } \
    } while (0)

// defined in vm_seed.cu
int vm_launch_sketch_chunks(const uint8_t *seq_dev, const int64_t *off_dev, const int64_t *chunk_off_dev, int n_seq, int64_t n_chunks, int w,
                            int k, uint64_t *mz_hash, uint32_t *mz_posz, int32_t *chunk_cnt, cudaStream_t stream);
int vm_sketch_chunk_size();

namespace {

__global__ void ixg_normalise_kernel(uint8_t *ref, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = vm_nt4(ref[i]);
    ref[i] = c == 0 ? 'A' : c == 1 ? 'C' : c == 2 ? 'G' : c == 3 ? 'T' : 'N';
}

// dense (hash, global position << 1 | strand) pairs from the per-chunk sketch outputs
__global__ void ixg_gather_kernel(const int64_t *__restrict__ off, const int64_t *__restrict__ chunk_off, int n_seq, int64_t n_chunks,
                                  int chunk, const int32_t *__restrict__ chunk_cnt, const int64_t *__restrict__ chunk_dst,
                                  const uint64_t *__restrict__ mz_hash, const uint32_t *__restrict__ mz_posz, uint64_t *keys,
                                  uint64_t *vals)
{
    const int64_t cid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cid >= n_chunks) return;
    int lo = 0, hi = n_seq;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (chunk_off[mid] <= cid) lo = mid; else hi = mid;
    }
    const int64_t base = off[lo];
    const int64_t src = base + (cid - chunk_off[lo]) * chunk;
    const int64_t dst = chunk_dst[cid];
    const int cnt = chunk_cnt[cid];
    for (int t = 0; t < cnt; ++t) {
        const uint32_t pz = mz_posz[src + t];
        keys[dst + t] = mz_hash[src + t];
        vals[dst + t] = ((uint64_t)(pz >> 1) + (uint64_t)base) << 1 | (uint64_t)(pz & 1u);
    }
}

__global__ void ixg_insert_kernel(const uint64_t *__restrict__ ukeys, const int64_t *__restrict__ ustart, const int32_t *__restrict__ ucnt,
                                  int64_t nk, VmHtSlot *ht, uint64_t mask)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nk) return;
    const uint64_t h = ukeys[i];
    uint64_t s = vm_ht_hash(h) & mask;
    for (;;) {
        const unsigned long long old = atomicCAS((unsigned long long *)&ht[s].key, (unsigned long long)VM_HT_EMPTY, (unsigned long long)h);
        if (old == (unsigned long long)VM_HT_EMPTY) break;
        s = (s + 1) & mask;
    }
    ht[s].start = (uint32_t)ustart[i];
    ht[s].count = (uint32_t)ucnt[i];
}

__global__ void ixg_ht_clear_kernel(VmHtSlot *ht, uint64_t slots)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= slots) return;
    VmHtSlot e;
    e.key = VM_HT_EMPTY; e.start = 0; e.count = 0;
    ht[i] = e;
}

// 5-letter code of the 9-mer starting at every position (all-N k-mers get the out-of-range code VM_K9_KEYS: they sort last)
__global__ void ixg_kmer_code_kernel(const uint8_t *__restrict__ ref, int64_t n_pos, uint32_t *codes, uint32_t *pos)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pos) return;
    uint32_t code = 0;
#pragma unroll
    for (int i = 0; i < VM_K9; ++i) code = code * 5u + (uint32_t)vm_code5(ref[p + i]);
    codes[p] = code == (uint32_t)(VM_K9_KEYS - 1) ? (uint32_t)VM_K9_KEYS : code;
    pos[p] = (uint32_t)p;
}

// koff[c] = first index of code >= c in the sorted code array (c in 0 .. VM_K9_KEYS)
__global__ void ixg_koff_kernel(const uint32_t *__restrict__ codes, int64_t n, int64_t *koff)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    const long long prev = i == 0 ? -1 : (long long)codes[i - 1];
    const long long cur = i == n ? (long long)VM_K9_KEYS : (long long)codes[i];
    for (long long c = prev + 1; c <= cur && c <= VM_K9_KEYS; ++c) koff[c] = i;
}

struct Tmp {        // a device allocation that frees itself
    void *p = nullptr;
    ~Tmp() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { if (p) cudaFree(p); p = nullptr; return cudaMalloc(&p, bytes ? bytes : 16); }
    template <typename T> T *as() { return (T *)p; }
    void *release() { void *q = p; p = nullptr; return q; }
};

} // namespace

// ix: names / ctg_start / ctg_len / w / k set, ix->ref = the concatenated sequence as given (any case, any letters).
// On return the device arrays and ix->dev are set, ix->ref is normalised (upper-case ACGTN), n_keys / n_occ / mid_occ filled.
int vm_index_build_device(VmIndex *ix, std::string &err)
{
    cudaStream_t stream = nullptr;
    const int64_t n = (int64_t)ix->ref.size();
    const int n_seq = (int)ix->names.size();
    const int w = ix->w, k = ix->k;
    // ---- reference on the device, normalised there ----
    IX_OK(cudaMalloc(&ix->d_ref, (size_t)n + 64));
    IX_OK(cudaMemcpy(ix->d_ref, ix->ref.data(), (size_t)n, cudaMemcpyHostToDevice));
    IX_OK(cudaMemset((uint8_t *)ix->d_ref + n, 'N', 64));
    if (n > 0) ixg_normalise_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>((uint8_t *)ix->d_ref, n);
    IX_OK(cudaMemcpy(&ix->ref[0], ix->d_ref, (size_t)n, cudaMemcpyDeviceToHost));
    // ---- minimizers ----
    const int chunk = vm_sketch_chunk_size();
    std::vector<int64_t> off((size_t)n_seq + 1), chunk_off((size_t)n_seq + 1, 0);
    for (int c = 0; c < n_seq; ++c) { off[(size_t)c] = ix->ctg_start[(size_t)c]; chunk_off[(size_t)c + 1] = chunk_off[(size_t)c] + (ix->ctg_len[(size_t)c] + chunk - 1) / chunk; }
    off[(size_t)n_seq] = n;
    const int64_t n_chunks = chunk_off[(size_t)n_seq];
    int64_t n_occ = 0, nk = 0;
    {
        Tmp d_off, d_choff, mz_hash, mz_posz, ccnt, cdst, scan_tmp;
        IX_OK(d_off.alloc(((size_t)n_seq + 1) * 8));
        IX_OK(d_choff.alloc(((size_t)n_seq + 1) * 8));
        IX_OK(cudaMemcpy(d_off.p, off.data(), ((size_t)n_seq + 1) * 8, cudaMemcpyHostToDevice));
        IX_OK(cudaMemcpy(d_choff.p, chunk_off.data(), ((size_t)n_seq + 1) * 8, cudaMemcpyHostToDevice));
        IX_OK(mz_hash.alloc((size_t)n * 8 + 64));
        IX_OK(mz_posz.alloc((size_t)n * 4 + 64));
        IX_OK(ccnt.alloc((size_t)(n_chunks + 1) * 4));
        IX_OK(cdst.alloc((size_t)(n_chunks + 1) * 8));
        IX_OK(cudaMemset(ccnt.p, 0, (size_t)(n_chunks + 1) * 4));
        vm_launch_sketch_chunks((const uint8_t *)ix->d_ref, d_off.as<int64_t>(), d_choff.as<int64_t>(), n_seq, n_chunks, w, k, mz_hash.as<uint64_t>(),
                                mz_posz.as<uint32_t>(), ccnt.as<int32_t>(), stream);
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, ccnt.as<int32_t>(), cdst.as<int64_t>(), (int)(n_chunks + 1), stream);
        IX_OK(scan_tmp.alloc(tb));
        IX_OK(cub::DeviceScan::ExclusiveSum(scan_tmp.p, tb, ccnt.as<int32_t>(), cdst.as<int64_t>(), (int)(n_chunks + 1), stream));
        IX_OK(cudaMemcpy(&n_occ, cdst.as<int64_t>() + n_chunks, 8, cudaMemcpyDeviceToHost));
        Tmp keys_a, vals_a, keys_b, vals_b, sort_tmp;
        IX_OK(keys_a.alloc((size_t)n_occ * 8));
        IX_OK(vals_a.alloc((size_t)n_occ * 8));
        if (n_chunks > 0)
            ixg_gather_kernel<<<(unsigned)((n_chunks + 127) / 128), 128, 0, stream>>>(d_off.as<int64_t>(), d_choff.as<int64_t>(), n_seq, n_chunks, chunk,
                                                                                     ccnt.as<int32_t>(), cdst.as<int64_t>(), mz_hash.as<uint64_t>(),
                                                                                     mz_posz.as<uint32_t>(), keys_a.as<uint64_t>(), vals_a.as<uint64_t>());
        IX_OK(cudaStreamSynchronize(stream));
        cudaFree(mz_hash.release());
        cudaFree(mz_posz.release());
        IX_OK(keys_b.alloc((size_t)n_occ * 8));
        IX_OK(vals_b.alloc((size_t)n_occ * 8));
        // stable LSD radix sort by hash (2k bits): occurrences of a key keep their ascending positions
        tb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, keys_a.as<uint64_t>(), keys_b.as<uint64_t>(), vals_a.as<uint64_t>(), vals_b.as<uint64_t>(), n_occ, 0,
                                        2 * k, stream);
        IX_OK(sort_tmp.alloc(tb));
        IX_OK(cub::DeviceRadixSort::SortPairs(sort_tmp.p, tb, keys_a.as<uint64_t>(), keys_b.as<uint64_t>(), vals_a.as<uint64_t>(), vals_b.as<uint64_t>(),
                                              n_occ, 0, 2 * k, stream));
        IX_OK(cudaStreamSynchronize(stream));
        cudaFree(keys_a.release());
        cudaFree(vals_a.release());
        cudaFree(sort_tmp.release());
        // unique keys, counts, starts
        Tmp ukeys, ucnt, ustart, nruns, rle_tmp;
        IX_OK(ukeys.alloc((size_t)n_occ * 8));
        IX_OK(ucnt.alloc((size_t)n_occ * 4));
        IX_OK(nruns.alloc(16));
        tb = 0;
        cub::DeviceRunLengthEncode::Encode(nullptr, tb, keys_b.as<uint64_t>(), ukeys.as<uint64_t>(), ucnt.as<int32_t>(), nruns.as<int64_t>(), n_occ, stream);
        IX_OK(rle_tmp.alloc(tb));
        IX_OK(cub::DeviceRunLengthEncode::Encode(rle_tmp.p, tb, keys_b.as<uint64_t>(), ukeys.as<uint64_t>(), ucnt.as<int32_t>(), nruns.as<int64_t>(), n_occ,
                                                 stream));
        IX_OK(cudaMemcpy(&nk, nruns.p, 8, cudaMemcpyDeviceToHost));
        if (n_occ == 0) nk = 0;
        IX_OK(ustart.alloc((size_t)(nk + 1) * 8));
        tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, ucnt.as<int32_t>(), ustart.as<int64_t>(), (int)nk, stream);
        Tmp scan2;
        IX_OK(scan2.alloc(tb));
        if (nk > 0) IX_OK(cub::DeviceScan::ExclusiveSum(scan2.p, tb, ucnt.as<int32_t>(), ustart.as<int64_t>(), (int)nk, stream));
        // the table
        uint64_t slots = 1024;
        while (slots < (uint64_t)nk * 2 + 16) slots <<= 1;
        IX_OK(cudaMalloc(&ix->d_ht, slots * sizeof(VmHtSlot)));
        ixg_ht_clear_kernel<<<(unsigned)((slots + 255) / 256), 256, 0, stream>>>((VmHtSlot *)ix->d_ht, slots);
        if (nk > 0)
            ixg_insert_kernel<<<(unsigned)((nk + 255) / 256), 256, 0, stream>>>(ukeys.as<uint64_t>(), ustart.as<int64_t>(), ucnt.as<int32_t>(), nk,
                                                                               (VmHtSlot *)ix->d_ht, slots - 1);
        ix->ht_slots = slots;
        // default occurrence cap: mm_idx_cal_max_occ(mi, 2e-4) then mm_mapopt_update's clamps [10, 1000000]
        ix->mid_occ_default = 10;
        if (nk > 0) {
            Tmp sorted_cnt, st;
            IX_OK(sorted_cnt.alloc((size_t)nk * 4));
            tb = 0;
            cub::DeviceRadixSort::SortKeys(nullptr, tb, ucnt.as<uint32_t>(), sorted_cnt.as<uint32_t>(), nk, 0, 32, stream);
            IX_OK(st.alloc(tb));
            IX_OK(cub::DeviceRadixSort::SortKeys(st.p, tb, ucnt.as<uint32_t>(), sorted_cnt.as<uint32_t>(), nk, 0, 32, stream));
            int64_t kth = (int64_t)((1. - 2e-4f) * (double)nk);
            if (kth >= nk) kth = nk - 1;
            uint32_t v = 0;
            IX_OK(cudaMemcpy(&v, sorted_cnt.as<uint32_t>() + kth, 4, cudaMemcpyDeviceToHost));
            int64_t thres = (int64_t)v + 1;
            thres = std::max<int64_t>(10, std::min<int64_t>(thres, 1000000));
            ix->mid_occ_default = (int)thres;
        }
        IX_OK(cudaStreamSynchronize(stream));
        ix->d_occ = vals_b.release();
        // the unique keys / counts stay for the .mmi writer (vm_index_minimizers); small next to the occurrences
        ix->d_ukeys = ukeys.release();
        ix->d_ucnt = ucnt.release();
        cudaFree(keys_b.release());
    }
    ix->n_keys = nk;
    ix->n_occ = n_occ;
    // ---- 9-mer positions ----
    const int64_t n_pos = n >= VM_K9 ? n - VM_K9 + 1 : 0;
    IX_OK(cudaMalloc(&ix->d_koff, ((size_t)VM_K9_KEYS + 1) * 8));
    IX_OK(cudaMemset(ix->d_koff, 0, ((size_t)VM_K9_KEYS + 1) * 8));
    int64_t n_valid = 0;
    {
        Tmp codes_a, pos_a, codes_b, pos_b, st;
        IX_OK(codes_a.alloc((size_t)n_pos * 4 + 16));
        IX_OK(pos_a.alloc((size_t)n_pos * 4 + 16));
        IX_OK(codes_b.alloc((size_t)n_pos * 4 + 16));
        IX_OK(pos_b.alloc((size_t)n_pos * 4 + 16));
        if (n_pos > 0) ixg_kmer_code_kernel<<<(unsigned)((n_pos + 255) / 256), 256, 0, stream>>>((const uint8_t *)ix->d_ref, n_pos, codes_a.as<uint32_t>(), pos_a.as<uint32_t>());
        size_t tb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, codes_a.as<uint32_t>(), codes_b.as<uint32_t>(), pos_a.as<uint32_t>(), pos_b.as<uint32_t>(), n_pos, 0, 21, stream);
        IX_OK(st.alloc(tb));
        if (n_pos > 0)
            IX_OK(cub::DeviceRadixSort::SortPairs(st.p, tb, codes_a.as<uint32_t>(), codes_b.as<uint32_t>(), pos_a.as<uint32_t>(), pos_b.as<uint32_t>(), n_pos, 0, 21,
                                                  stream));
        ixg_koff_kernel<<<(unsigned)((n_pos + 1 + 255) / 256), 256, 0, stream>>>(codes_b.as<uint32_t>(), n_pos, (int64_t *)ix->d_koff);
        IX_OK(cudaMemcpy(&n_valid, (int64_t *)ix->d_koff + VM_K9_KEYS, 8, cudaMemcpyDeviceToHost));
        IX_OK(cudaStreamSynchronize(stream));
        ix->d_kpos = pos_b.release();      // the all-N tail behind koff[VM_K9_KEYS] is never addressed
    }
    ix->n_kpos = n_valid;
    IX_OK(cudaGetLastError());
    ix->dev.ht = (const VmHtSlot *)ix->d_ht;
    ix->dev.ht_mask = ix->ht_slots - 1;
    ix->dev.occ = (const uint64_t *)ix->d_occ;
    ix->dev.kpos = (const uint32_t *)ix->d_kpos;
    ix->dev.koff = (const int64_t *)ix->d_koff;
    ix->dev.ref = (const uint8_t *)ix->d_ref;
    ix->dev.ref_len = n;
    ix->dev.w = w;
    ix->dev.k = k;
    ix->dev.mid_occ = ix->mid_occ_default;
    return vm_index_build_buckets(ix, err);
}

// ---------------------------------------------------------------------------
// Position directory of the long 9-mer runs.  One block per run: every entry whose bucket differs from its
// predecessor's writes its index into the buckets in between (lower_bound of each bucket start), the tail is the
// run length.
// ---------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) ixg_bucket_kernel(const uint32_t *__restrict__ kpos, const int64_t *__restrict__ koff,
                                                         const int32_t *__restrict__ row_code, int shift, int nb, uint32_t *__restrict__ kbk)
{
    const int code = row_code[blockIdx.x];
    const int64_t b = koff[code];
    const uint32_t len = (uint32_t)(koff[code + 1] - b);
    uint32_t *t = kbk + (size_t)blockIdx.x * (size_t)(nb + 1);
    const uint32_t *run = kpos + b;
    for (uint32_t i = threadIdx.x; i <= len; i += blockDim.x) {
        const int cur = i < len ? (int)(run[i] >> shift) : nb;            // past the end: every bucket left gets `len`
        const int prev = i > 0 ? (int)(run[i - 1] >> shift) : -1;
        for (int q = prev + 1; q <= cur; ++q) t[q] = i;
    }
}
} // namespace

#define VM_KB_MIN_RUN 128          // shorter runs: the plain binary search stays inside a few sectors anyway
#define VM_KB_MAX_BUCKETS 2048
#define VM_KB_MAX_BYTES (6ULL << 30)

int vm_index_build_buckets(VmIndex *ix, std::string &err)
{
    ix->dev.krow = nullptr; ix->dev.kbk = nullptr; ix->dev.kb_shift = 0; ix->dev.kb_n = 0;
    if (getenv("VM_NO_KPOS_DIRECTORY") || ix->dev.ref_len <= 0 || !ix->d_koff || !ix->d_kpos) return 0;
    std::vector<int64_t> koff((size_t)VM_K9_KEYS + 1);
    if (cudaMemcpy(koff.data(), ix->d_koff, koff.size() * 8, cudaMemcpyDeviceToHost) != cudaSuccess) { err = "position directory: cannot read koff"; return -1; }
    std::vector<int32_t> krow((size_t)VM_K9_KEYS, -1), row_code;
    const int64_t min_run = getenv("VM_KB_MIN_RUN") ? atoll(getenv("VM_KB_MIN_RUN")) : VM_KB_MIN_RUN;               // experiment knobs
    const int64_t max_buckets = getenv("VM_KB_MAX_BUCKETS") ? atoll(getenv("VM_KB_MAX_BUCKETS")) : VM_KB_MAX_BUCKETS;
    for (int c = 0; c < VM_K9_KEYS; ++c)
        if (koff[(size_t)c + 1] - koff[(size_t)c] >= min_run) { krow[(size_t)c] = (int32_t)row_code.size(); row_code.push_back(c); }
    if (row_code.empty()) return 0;
    int shift = 8;
    while (((ix->dev.ref_len >> shift) + 1) > max_buckets ||
           (unsigned long long)row_code.size() * (unsigned long long)((ix->dev.ref_len >> shift) + 2) * 4ULL > VM_KB_MAX_BYTES) ++shift;
    const int nb = (int)(ix->dev.ref_len >> shift) + 1;
    void *d_codes = nullptr;
    cudaError_t e = cudaMalloc(&ix->d_krow, krow.size() * 4);
    if (e == cudaSuccess) e = cudaMalloc(&ix->d_kbk, row_code.size() * (size_t)(nb + 1) * 4);
    if (e == cudaSuccess) e = cudaMalloc(&d_codes, row_code.size() * 4);
    if (e == cudaSuccess) e = cudaMemcpy(ix->d_krow, krow.data(), krow.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_codes, row_code.data(), row_code.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        ixg_bucket_kernel<<<(unsigned)row_code.size(), 256>>>((const uint32_t *)ix->d_kpos, (const int64_t *)ix->d_koff, (const int32_t *)d_codes, shift, nb,
                                                              (uint32_t *)ix->d_kbk);
        e = cudaDeviceSynchronize();
    }
    if (d_codes) cudaFree(d_codes);
    if (e != cudaSuccess) {
        // not enough memory for the directory: the plain search still works
        cudaGetLastError();
        if (ix->d_krow) { cudaFree(ix->d_krow); ix->d_krow = nullptr; }
        if (ix->d_kbk) { cudaFree(ix->d_kbk); ix->d_kbk = nullptr; }
        return 0;
    }
    ix->dev.krow = (const int32_t *)ix->d_krow;
    ix->dev.kbk = (const uint32_t *)ix->d_kbk;
    ix->dev.kb_shift = shift;
    ix->dev.kb_n = nb;
    return 0;
}
