// Minimizer seeding on the GPU: sketch -> hash-table lookup -> anchor expansion ->
// "top N clusters" filter -> majority-strand flip.
//
// Replaces `index_object.map(seq, check_num, mid_occ)` (mammap_clrnano.py:23985; un-vendored
// vacmap_index, restated from minimap2 -- see oracle/orc_index.c for the rules this must equal)
// and get_reversed_chain_numpy_rough (:21202-21217).
//
// Memory behaviour: the sketch streams each read once (1 B/base); the lookup is the
// HBM-latency-bound part (one random 16-byte slot probe per minimizer, then the occurrence
// run, 8 B each, contiguous); the expansion writes 16 B per anchor, coalesced per minimizer.
#include "vm_seed.cuh"

// ---- 1. sketch, chunk-parallel and exact ----
// mm_sketch is a sequential window machine, but its state is history-free after a short clean
// run: once k non-ambiguous bases have refilled the k-mer registers and w+k further positions
// were all "counted" (neither ambiguous nor a symmetric k-mer), the last w window slots, the
// current minimum (always the latest slot attaining the window minimum) and the saturated run
// length are functions of those positions alone.  So one thread per VM_SK_CHUNK positions
// starts a FRESH machine a warm-up before its chunk, keeps only emissions whose position falls
// inside its chunk, and runs on until w slots past the chunk end (every slot is emitted at the
// latest when it leaves the window).  If the warm-up zone is not clean (N's, symmetric k-mers)
// it is widened x4 until it is, or until it reaches the start of the read, where the machine is
// exact by definition.  Emissions are position-sorted, so concatenating the chunks in order
// reproduces the sequential output.
#define VM_SK_CHUNK 128

__global__ void vm_sketch_chunk_kernel(const uint8_t *__restrict__ reads, const int64_t *__restrict__ off,
                                       const int64_t *__restrict__ chunk_off, int n_reads, int64_t n_chunks, int w, int k,
                                       uint64_t *__restrict__ mz_hash, uint32_t *__restrict__ mz_posz,
                                       int32_t *__restrict__ chunk_cnt)
{
    const int64_t cid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cid >= n_chunks) return;
    int lo = 0, hi = n_reads;   // read owning this chunk
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (chunk_off[mid] <= cid) lo = mid; else hi = mid;
    }
    const int64_t base = off[lo];
    const int64_t L = off[lo + 1] - base;
    const int64_t c0 = (cid - chunk_off[lo]) * VM_SK_CHUNK;
    const int64_t c1 = c0 + VM_SK_CHUNK < L ? c0 + VM_SK_CHUNK : L;
    const unsigned char *str = reads + base;
    uint64_t *oh = mz_hash + base + c0;
    uint32_t *op = mz_posz + base + c0;
    int64_t warm = 2 * k + w + 8;
    int cnt = 0;
    for (;;) {
        const int64_t s = c0 - warm > 0 ? c0 - warm : 0;
        cnt = 0;
        bool regs_bad = false;
        int64_t last_bad = -1;
        int slots_after = 0;
        vm_sketch_range(str, s, L, w, k, true,
                        [&](uint64_t h, uint64_t y) {
                            const int64_t pos = (int64_t)(y >> 1);
                            if (pos >= c0 && pos < c1) { oh[cnt] = h; op[cnt] = (uint32_t)y; ++cnt; }
                        },
                        [&](int64_t i, int kind) {
                            if (i < c0) {
                                if (kind == 1 && i < s + k) regs_bad = true;
                                if (kind != 0) last_bad = i;
                            } else if (i >= c1) {
                                if (kind != 2 && ++slots_after > w) return false;
                            }
                            return true;
                        });
        // the first k positions of the clean run refill the registers (their symmetric test may still
        // see stale bits), the following w + k are judged with correct registers
        (void)regs_bad;
        if (s == 0 || last_bad < c0 - (w + 2 * k)) break;
        warm *= 4;
    }
    chunk_cnt[cid] = cnt;
}

// the same kernel for the contigs of a reference (vm_index_gpu.cu)
int vm_sketch_chunk_size() { return VM_SK_CHUNK; }
int vm_launch_sketch_chunks(const uint8_t *seq_dev, const int64_t *off_dev, const int64_t *chunk_off_dev, int n_seq, int64_t n_chunks, int w,
                            int k, uint64_t *mz_hash, uint32_t *mz_posz, int32_t *chunk_cnt, cudaStream_t stream)
{
    if (n_chunks <= 0) return 0;
    vm_sketch_chunk_kernel<<<(unsigned)((n_chunks + 127) / 128), 128, 0, stream>>>(seq_dev, off_dev, chunk_off_dev, n_seq, n_chunks, w, k, mz_hash,
                                                                                  mz_posz, chunk_cnt);
    return 1;
}

// compaction of the per-chunk outputs to the front of each read's slot range: one warp per read
__global__ void __launch_bounds__(32) vm_sketch_compact_kernel(const int64_t *__restrict__ off, const int64_t *__restrict__ chunk_off,
                                                               const int32_t *__restrict__ chunk_cnt, uint64_t *__restrict__ mz_hash,
                                                               uint32_t *__restrict__ mz_posz, int32_t *__restrict__ n_mz)
{
    const int r = blockIdx.x;
    const int lane = threadIdx.x;
    const int64_t base = off[r];
    const int64_t c_lo = chunk_off[r], c_hi = chunk_off[r + 1];
    int total = 0;
    for (int64_t c = c_lo; c < c_hi; ++c) {
        const int cnt = chunk_cnt[c];
        const int64_t src = base + (c - c_lo) * VM_SK_CHUNK, dst = base + total;
        if (src != dst) {
            for (int t0 = 0; t0 < cnt; t0 += 32) {
                const int t = t0 + lane;
                uint64_t h = 0;
                uint32_t p = 0;
                if (t < cnt) { h = mz_hash[src + t]; p = mz_posz[src + t]; }
                __syncwarp();
                if (t < cnt) { mz_hash[dst + t] = h; mz_posz[dst + t] = p; }
                __syncwarp();
            }
        }
        total += cnt;
    }
    if (lane == 0) n_mz[r] = total;
}

// ---- 2. lookup + per-read exclusive scan of hit counts: one warp per read ----
__global__ void __launch_bounds__(32) vm_seed_lookup_kernel(VmIndexDev ix, const int64_t *__restrict__ off,
                                                            const uint64_t *__restrict__ mz_hash,
                                                            const int32_t *__restrict__ n_mz, int mid_occ,
                                                            uint32_t *__restrict__ mz_start, uint32_t *__restrict__ mz_cnt,
                                                            uint32_t *__restrict__ mz_aoff, int32_t *__restrict__ n_anchor)
{
    const int r = blockIdx.x;
    const int lane = threadIdx.x;
    const int64_t base = off[r];
    const int m = n_mz[r];
    uint32_t running = 0;
    for (int t0 = 0; t0 < m; t0 += 32) {
        const int t = t0 + lane;
        uint32_t start = 0, cnt = 0;
        if (t < m) {
            const uint64_t h = mz_hash[base + t];
            uint64_t s = vm_ht_hash(h) & ix.ht_mask;
            for (;;) {
                const VmHtSlot slot = ix.ht[s];
                if (slot.key == h) { start = slot.start; cnt = slot.count; break; }
                if (slot.key == VM_HT_EMPTY) break;
                s = (s + 1) & ix.ht_mask;
            }
            if ((int64_t)cnt > (int64_t)mid_occ) cnt = 0;   // occurrence cap
            mz_start[base + t] = start;
            mz_cnt[base + t] = cnt;
        }
        uint32_t inc = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(VM_FULL, inc, d);
            if (lane >= d) inc += o;
        }
        if (t < m) mz_aoff[base + t] = running + inc - cnt;
        running += __shfl_sync(VM_FULL, inc, 31);
    }
    if (lane == 0) n_anchor[r] = (int32_t)running;
}

// ---- 3. expansion: one warp per read, anchors in (minimizer, occurrence) order ----
__global__ void __launch_bounds__(32) vm_seed_expand_kernel(VmIndexDev ix, const int64_t *__restrict__ off,
                                                            const uint32_t *__restrict__ mz_posz,
                                                            const int32_t *__restrict__ n_mz,
                                                            const uint32_t *__restrict__ mz_start,
                                                            const uint32_t *__restrict__ mz_cnt,
                                                            const uint32_t *__restrict__ mz_aoff,
                                                            const int64_t *__restrict__ a_off, VmAnchor *__restrict__ anchors)
{
    const int r = blockIdx.x;
    const int lane = threadIdx.x;
    const int64_t base = off[r];
    const int m = n_mz[r];
    VmAnchor *out = anchors + a_off[r];
    const int k = ix.k;
    for (int t = 0; t < m; ++t) {
        const uint32_t cnt = mz_cnt[base + t];
        if (cnt == 0) continue;
        const uint32_t start = mz_start[base + t], ao = mz_aoff[base + t], pz = mz_posz[base + t];
        const int qpos = (int)(pz >> 1), qz = (int)(pz & 1);
        for (uint32_t o = lane; o < cnt; o += 32) {
            const uint64_t y = ix.occ[start + o];
            VmAnchor a;
            a.x = qpos - k + 1;
            a.y = (uint32_t)((y >> 1) - (uint64_t)k + 1);
            a.s = ((int)(y & 1) == qz) ? 1 : -1;
            a.l = k;
            out[ao + o] = a;
        }
    }
}

// ---- 4. cluster filter + strand flip: one block per read ----
#define VM_CL_EMPTY 0x7fffffffffffffffLL
struct VmClSlot {
    long long key;
    int count;
    int first;
};

__global__ void vm_cl_init_kernel(VmClSlot *t, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { t[i].key = VM_CL_EMPTY; t[i].count = 0; t[i].first = 0x7fffffff; }
}

__device__ __forceinline__ long long vm_cluster_key(const VmAnchor &a)
{
    const long long diag = a.s == 1 ? (long long)a.y - a.x : (long long)a.y + a.x;
    long long q = diag / 5000;
    if ((diag % 5000 != 0) && (diag < 0)) --q;   // floor division
    return q * 2 + (a.s == 1 ? 0 : 1);
}

__global__ void __launch_bounds__(128) vm_seed_filter_kernel(const VmAnchor *__restrict__ in, const int64_t *__restrict__ a_off,
                                                             const int32_t *__restrict__ n_anchor, const int64_t *__restrict__ off,
                                                             int check_num, VmClSlot *__restrict__ table_all,
                                                             const int64_t *__restrict__ t_off, int *__restrict__ compact_all,
                                                             VmAnchor *__restrict__ out, int32_t *__restrict__ n_out,
                                                             int32_t *__restrict__ need_rev)
{
    __shared__ int s_cnt[4];
    __shared__ int s_scan[128];
    const int r = blockIdx.x;
    const int tid = threadIdx.x;
    const int n = n_anchor[r];
    const int64_t L = off[r + 1] - off[r];
    const VmAnchor *a = in + a_off[r];
    VmAnchor *o = out + a_off[r];
    VmClSlot *tab = table_all + t_off[r];
    const int tsize = (int)(t_off[r + 1] - t_off[r]);   // power of two >= 2n (0 when filtering is off)
    int *compact = compact_all + 2 * t_off[r];
    int *dropped = compact + tsize;                      // per table slot
    bool filter = false;
    if (tid < 4) s_cnt[tid] = 0;
    __syncthreads();
    if (check_num >= 0 && n > check_num && tsize > 0) {
        for (int i = tid; i < n; i += blockDim.x) {
            const long long key = vm_cluster_key(a[i]);
            unsigned long long h = (unsigned long long)key * 0x9E3779B97F4A7C15ULL;
            int s = (int)((h ^ (h >> 31)) & (unsigned long long)(tsize - 1));
            for (;;) {
                const long long prev = (long long)atomicCAS((unsigned long long *)&tab[s].key, (unsigned long long)VM_CL_EMPTY,
                                                            (unsigned long long)key);
                if (prev == VM_CL_EMPTY || prev == key) break;
                s = (s + 1) & (tsize - 1);
            }
            atomicAdd(&tab[s].count, 1);
            atomicMin(&tab[s].first, i);
        }
        __syncthreads();
        for (int s = tid; s < tsize; s += blockDim.x)
            if (tab[s].key != VM_CL_EMPTY) compact[atomicAdd(&s_cnt[0], 1)] = s;
        __syncthreads();
        const int C = s_cnt[0];
        if (C > check_num) {
            filter = true;
            // rank every cluster by (count desc, first appearance asc); keep the top check_num
            for (int ci = tid; ci < C; ci += blockDim.x) {
                const VmClSlot me = tab[compact[ci]];
                int rank = 0;
                for (int cj = 0; cj < C; ++cj) {
                    const VmClSlot ot = tab[compact[cj]];
                    if (ot.count > me.count || (ot.count == me.count && ot.first < me.first)) ++rank;
                }
                dropped[compact[ci]] = rank >= check_num ? 1 : 0;
            }
        }
        __syncthreads();
    }
    // pass 1: kept count and strand census
    int kept = 0, pos = 0, neg = 0;
    for (int i = tid; i < n; i += blockDim.x) {
        bool keep = true;
        if (filter) {
            const long long key = vm_cluster_key(a[i]);
            unsigned long long h = (unsigned long long)key * 0x9E3779B97F4A7C15ULL;
            int s = (int)((h ^ (h >> 31)) & (unsigned long long)(tsize - 1));
            while (tab[s].key != key) s = (s + 1) & (tsize - 1);
            keep = dropped[s] == 0;
        }
        if (keep) { ++kept; if (a[i].s == 1) ++pos; else ++neg; }
    }
    atomicAdd(&s_cnt[1], kept);
    atomicAdd(&s_cnt[2], pos);
    atomicAdd(&s_cnt[3], neg);
    __syncthreads();
    const int m = s_cnt[1];
    const bool flip = m >= 3 && s_cnt[3] > s_cnt[2];
    // pass 2: stable compaction (block scan per chunk), flipped rows are transformed and reversed
    int running = 0;
    for (int i0 = 0; i0 < n; i0 += blockDim.x) {
        const int i = i0 + tid;
        bool keep = false;
        VmAnchor v;
        if (i < n) {
            v = a[i];
            keep = true;
            if (filter) {
                const long long key = vm_cluster_key(v);
                unsigned long long h = (unsigned long long)key * 0x9E3779B97F4A7C15ULL;
                int s = (int)((h ^ (h >> 31)) & (unsigned long long)(tsize - 1));
                while (tab[s].key != key) s = (s + 1) & (tsize - 1);
                keep = dropped[s] == 0;
            }
        }
        s_scan[tid] = keep ? 1 : 0;
        __syncthreads();
        for (int d = 1; d < 128; d <<= 1) {
            const int t = tid >= d ? s_scan[tid - d] : 0;
            __syncthreads();
            s_scan[tid] += t;
            __syncthreads();
        }
        const int rank = running + s_scan[tid] - (keep ? 1 : 0);
        const int total = s_scan[127];
        __syncthreads();
        if (keep) {
            if (flip) {
                v.x = (int)(L - v.x - v.l);
                v.s = -v.s;
                o[m - 1 - rank] = v;
            } else o[rank] = v;
        }
        running += total;
    }
    if (tid == 0) { n_out[r] = m; need_rev[r] = flip ? 1 : 0; }
}

// ---------------------------------------------------------------------------
int vm_seed_batch(VmSeedBufs &B, const VmIndexDev &ix, const uint8_t *reads_dev, const int64_t *off_dev,
                  const std::vector<int64_t> &off_host, int check_num, int mid_occ, cudaStream_t stream,
                  std::vector<int32_t> &n_out, std::vector<int32_t> &need_rev, std::vector<int64_t> &a_off_host,
                  int64_t *launches, std::string &err)
{
    const int n = (int)off_host.size() - 1;
    n_out.assign(n, 0);
    need_rev.assign(n, 0);
    a_off_host.assign(n + 1, 0);
    if (n == 0) return 0;
    const size_t total = (size_t)off_host[n];
#define SEED_OK(call)                                                                  \
    do { cudaError_t _e = (call); if (_e != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(_e); return -1; } } while (0)
    SEED_OK(B.mz_hash.ensure(total * 8 + 64));
    SEED_OK(B.mz_posz.ensure(total * 4 + 64));
    SEED_OK(B.mz_start.ensure(total * 4 + 64));
    SEED_OK(B.mz_cnt.ensure(total * 4 + 64));
    SEED_OK(B.mz_aoff.ensure(total * 4 + 64));
    SEED_OK(B.n_mz.ensure((size_t)n * 4 + 64));
    SEED_OK(B.n_anchor.ensure((size_t)n * 4 + 64));
    SEED_OK(B.n_out.ensure((size_t)n * 4 + 64));
    SEED_OK(B.need_rev.ensure((size_t)n * 4 + 64));
    SEED_OK(B.a_off.ensure((size_t)(n + 1) * 8 + 64));
    SEED_OK(B.t_off.ensure((size_t)(n + 1) * 8 + 64));
    if (mid_occ < 0) mid_occ = ix.mid_occ;
    {
        std::vector<int64_t> chunk_off(n + 1, 0);
        for (int r = 0; r < n; ++r)
            chunk_off[r + 1] = chunk_off[r] + (off_host[r + 1] - off_host[r] + VM_SK_CHUNK - 1) / VM_SK_CHUNK;
        const int64_t n_chunks = chunk_off[n];
        SEED_OK(B.chunk_off.ensure((size_t)(n + 1) * 8 + 64));
        SEED_OK(B.chunk_cnt.ensure((size_t)n_chunks * 4 + 64));
        // (pageable source: cudaMemcpyAsync returns once it has been staged, so the vector may go out of scope -- no sync)
        SEED_OK(cudaMemcpyAsync(B.chunk_off.p, chunk_off.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, stream));
        if (n_chunks > 0)
            vm_sketch_chunk_kernel<<<(unsigned)((n_chunks + 63) / 64), 64, 0, stream>>>(
                reads_dev, off_dev, B.chunk_off.as<int64_t>(), n, n_chunks, ix.w, ix.k, B.mz_hash.as<uint64_t>(),
                B.mz_posz.as<uint32_t>(), B.chunk_cnt.as<int32_t>());
        vm_sketch_compact_kernel<<<n, 32, 0, stream>>>(off_dev, B.chunk_off.as<int64_t>(), B.chunk_cnt.as<int32_t>(),
                                                       B.mz_hash.as<uint64_t>(), B.mz_posz.as<uint32_t>(), B.n_mz.as<int32_t>());
        *launches += 1;
    }
    vm_seed_lookup_kernel<<<n, 32, 0, stream>>>(ix, off_dev, B.mz_hash.as<uint64_t>(), B.n_mz.as<int32_t>(), mid_occ,
                                                B.mz_start.as<uint32_t>(), B.mz_cnt.as<uint32_t>(), B.mz_aoff.as<uint32_t>(),
                                                B.n_anchor.as<int32_t>());
    *launches += 2;
    std::vector<int32_t> n_anchor(n);
    SEED_OK(vm_d2h_sync(B.pin, n_anchor.data(), B.n_anchor.p, (size_t)n * 4, stream));
    std::vector<int64_t> t_off(n + 1, 0);
    for (int r = 0; r < n; ++r) {
        a_off_host[r + 1] = a_off_host[r] + n_anchor[r];
        int64_t ts = 0;
        if (check_num >= 0 && n_anchor[r] > check_num) {
            ts = 64;
            while (ts < 2LL * n_anchor[r]) ts <<= 1;
        }
        t_off[r + 1] = t_off[r] + ts;
    }
    const size_t ta = (size_t)a_off_host[n];
    SEED_OK(B.raw.ensure(ta * 16 + 64));
    SEED_OK(B.out.ensure(ta * 16 + 64));
    SEED_OK(B.table.ensure((size_t)t_off[n] * sizeof(VmClSlot) + 64));
    SEED_OK(B.compact.ensure((size_t)t_off[n] * 8 + 64));
    SEED_OK(cudaMemcpyAsync(B.a_off.p, a_off_host.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, stream));
    SEED_OK(cudaMemcpyAsync(B.t_off.p, t_off.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, stream));
    if (t_off[n] > 0) {
        vm_cl_init_kernel<<<(unsigned)((t_off[n] + 255) / 256), 256, 0, stream>>>(B.table.as<VmClSlot>(), t_off[n]);
        *launches += 1;
    }
    vm_seed_expand_kernel<<<n, 32, 0, stream>>>(ix, off_dev, B.mz_posz.as<uint32_t>(), B.n_mz.as<int32_t>(),
                                                B.mz_start.as<uint32_t>(), B.mz_cnt.as<uint32_t>(), B.mz_aoff.as<uint32_t>(),
                                                B.a_off.as<int64_t>(), B.raw.as<VmAnchor>());
    vm_seed_filter_kernel<<<n, 128, 0, stream>>>(B.raw.as<VmAnchor>(), B.a_off.as<int64_t>(), B.n_anchor.as<int32_t>(), off_dev,
                                                 check_num, B.table.as<VmClSlot>(), B.t_off.as<int64_t>(), B.compact.as<int>(),
                                                 B.out.as<VmAnchor>(), B.n_out.as<int32_t>(), B.need_rev.as<int32_t>());
    *launches += 2;
    SEED_OK(vm_d2h_sync(B.pin, n_out.data(), B.n_out.p, (size_t)n * 4, stream, need_rev.data(), B.need_rev.p, (size_t)n * 4));
    SEED_OK(cudaGetLastError());
#undef SEED_OK
    return 0;
}
