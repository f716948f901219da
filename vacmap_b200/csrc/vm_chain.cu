// Non-linear anchor chaining kernels (see vm_chain.cuh for the design note).
#include "vm_chain.cuh"
#include "vm_sort.cuh"
#include <math_constants.h>

// ---------------------------------------------------------------------------
// pack: int64[total][4] rows -> 16-byte VmAnchor (coalesced: one thread per row)
// ---------------------------------------------------------------------------
__global__ void vm_pack_kernel(const longlong4 *__restrict__ rows, VmAnchor *__restrict__ out, long long total)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    longlong4 r = rows[i];
    VmAnchor a;
    a.x = (int32_t)r.x;
    a.y = (uint32_t)r.y;
    a.s = (int32_t)r.z;
    a.l = (int32_t)r.w;
    out[i] = a;
}

int vm_launch_pack(const int64_t *rows_dev, VmAnchor *out, long long total, cudaStream_t stream)
{
    if (total <= 0) return 0;
    int threads = 256;
    long long blocks = (total + threads - 1) / threads;
    vm_pack_kernel<<<(unsigned)blocks, threads, 0, stream>>>((const longlong4 *)rows_dev, out, total);
    return 1;
}

// ---------------------------------------------------------------------------
// anchor sort: warp-cooperative replay of numba's argsort (vm_sort.cuh), one warp
// per read.  Keys are read positions (global DP, reference :23572) or read end
// positions (local DP, :28585).  The sorted anchors are gathered in the same
// kernel (coalesced 16-byte stores; the random 16-byte loads hit L2).
// ---------------------------------------------------------------------------
template <int KEY_IS_END, bool SMEM>
__global__ void __launch_bounds__(32) vm_sort_anchors_kernel(const VmAnchor *__restrict__ in,
                                                             const int64_t *__restrict__ off,
                                                             const int32_t *__restrict__ cnt,
                                                             const int *__restrict__ read_ids, int cap,
                                                             int32_t *__restrict__ perm, int32_t *__restrict__ gscratch,
                                                             VmAnchor *__restrict__ sorted,
                                                             longlong4 *__restrict__ sorted_rows)
{
    extern __shared__ __align__(16) unsigned char vm_smem[];
    const int lane = threadIdx.x;
    const int rid = read_ids[blockIdx.x];
    const long long base = off[rid];
    const int n = cnt[rid];
    if (n <= 0) return;
    const VmAnchor *A = in + base;
    int *keys, *R, *Lpos, *Rpos;
    if (SMEM) {
        keys = (int *)vm_smem;
        R = keys + cap;
        Lpos = R + cap;
        Rpos = Lpos + cap;
    } else {
        keys = gscratch + 3 * base;
        Lpos = keys + n;
        Rpos = Lpos + n;
        R = perm + base;
    }
    for (int t = lane; t < n; t += 32) {
        const VmAnchor a = A[t];
        keys[t] = KEY_IS_END ? a.x + a.l : a.x;
    }
    __syncwarp();
    vm_warp_argsort_replay<int>(keys, R, Lpos, Rpos, n, lane);
    for (int t = lane; t < n; t += 32) {
        const int src = R[t];
        if (SMEM) perm[base + t] = src;
        const VmAnchor a = A[src];
        sorted[base + t] = a;
        if (sorted_rows) {
            longlong4 r;
            r.x = a.x; r.y = (long long)a.y; r.z = a.s; r.w = a.l;
            sorted_rows[base + t] = r;
        }
    }
}

int vm_launch_sort_anchors(const VmAnchor *in, const int64_t *off, const int32_t *cnt, const int *read_ids_dev, int n_ids, int cap,
                           bool use_smem, int key_is_end, int32_t *perm, int32_t *gscratch, VmAnchor *sorted,
                           int64_t *sorted_rows, cudaStream_t stream)
{
    if (n_ids <= 0) return 0;
    longlong4 *rows = (longlong4 *)sorted_rows;
    if (use_smem) {
        const size_t smem = (size_t)cap * 16;
        if (key_is_end) {
            vm_smem_optin(vm_sort_anchors_kernel<1, true>);
            vm_sort_anchors_kernel<1, true><<<n_ids, 32, smem, stream>>>(in, off, cnt, read_ids_dev, cap, perm, gscratch, sorted, rows);
        } else {
            vm_smem_optin(vm_sort_anchors_kernel<0, true>);
            vm_sort_anchors_kernel<0, true><<<n_ids, 32, smem, stream>>>(in, off, cnt, read_ids_dev, cap, perm, gscratch, sorted, rows);
        }
    } else {
        if (key_is_end) vm_sort_anchors_kernel<1, false><<<n_ids, 32, 0, stream>>>(in, off, cnt, read_ids_dev, cap, perm, gscratch, sorted, rows);
        else vm_sort_anchors_kernel<0, false><<<n_ids, 32, 0, stream>>>(in, off, cnt, read_ids_dev, cap, perm, gscratch, sorted, rows);
    }
    return 1;
}

// ---------------------------------------------------------------------------
// exact chaining DP, one warp per read
// ---------------------------------------------------------------------------
struct VmScoreCtx {
    double skipcost;
    int maxdiff;
    int maxgap;
    const double *gcl;   // smem
    const float *rgl;    // smem
    const float *extra;
    long long extra_size;
    const double *log2cache;
    long long log2cache_size;
};

// Candidate score of chaining anchor i after predecessor j.  The order of the
// float64 additions follows the reference expressions exactly
// (:24990, :24996 global; :27462, :27476-27480 local; :28416-28428 multi-chain).
template <int VARIANT>
__device__ __forceinline__ double vm_pair_score(const VmScoreCtx &c, const VmAnchor &ai, const VmAnchor &aj,
                                                double Sj, bool &skip)
{
    // 3 / 4: asm mode's linked DPs (mammap_asm.py:21687-21871 first round; :21505-21686 second round = the same
    // with a read-gap cost on colinear pairs)
    constexpr bool GLOBAL = VARIANT == 0 || VARIANT >= 3;
    constexpr bool READGAP = VARIANT == 1 || VARIANT == 2 || VARIANT == 4;
    int bonus, readgap;
    long long refgap;
    if (VARIANT >= 3) vm_pair_gaps_asm(ai, aj, bonus, readgap, refgap);
    else vm_pair_gaps(ai, aj, bonus, readgap, refgap);
    skip = false;
    if (!GLOBAL && (ai.x - aj.x - aj.l) < 0 && bonus <= 0) { skip = true; return -CUDART_INF; }
    long long gapcost = vm_llabs((long long)readgap - refgap);
    if (ai.s == aj.s && refgap >= 0 && readgap <= c.maxgap && gapcost <= (long long)c.maxdiff) {
        double t = Sj + (double)bonus;
        t = t - c.gcl[gapcost];
        if (READGAP) t = t - (double)c.rgl[readgap];
        return t;
    }
    if (GLOBAL) {
        if (gapcost > c.extra_size) gapcost = c.extra_size;
        double t = Sj - c.skipcost;
        t = t + (double)bonus;
        t = t - (double)__ldg(c.extra + gapcost);
        return t;
    } else if (VARIANT == 1) {
        if (gapcost > c.extra_size) gapcost = c.extra_size;
        const double e = (double)__ldg(c.extra + gapcost);
        double pen;
        if (ai.s != aj.s) pen = (50.0 < c.skipcost ? 50.0 : c.skipcost) + e;
        else pen = c.skipcost + e;
        double t = Sj + (double)bonus;
        return t - pen;
    } else {
        const long long g = gapcost < c.log2cache_size ? gapcost : c.log2cache_size;
        const double pen = c.skipcost + __ldg(c.log2cache + g);
        double t = Sj + (double)bonus;
        return t - pen;
    }
}

// number of q in [0, hi) with S[arg[q]] < target (LEQ: <= target); arg ascending in S.
template <bool LEQ>
__device__ __forceinline__ int vm_warp_count(const double *S, const int32_t *arg, int lo, int hi,
                                             double target, int lane)
{
    while (lo < hi) {
        const int span = hi - lo;
        const int step = (span + 31) >> 5;
        int p = lo + (lane + 1) * step - 1;
        if (p > hi - 1) p = hi - 1;
        const double v = S[arg[p]];
        const bool pred = LEQ ? (v <= target) : (v < target);
        const int c = __popc(__ballot_sync(VM_FULL, pred));
        if (c == 32) {
            lo = hi;
        } else {
            int pc = lo + (c + 1) * step - 1;
            if (pc > hi - 1) pc = hi - 1;
            int nlo = lo;
            if (c > 0) {
                int pp = lo + c * step - 1;
                if (pp > hi - 1) pp = hi - 1;
                nlo = pp + 1;
            }
            hi = pc;
            lo = nlo;
        }
    }
    return lo;
}

// Insert index k into arg[0..k) keeping it ascending in S; ties placed exactly
// where the reference's search puts them.
template <int VARIANT>
__device__ __forceinline__ void vm_insert_one(const double *S, int32_t *arg, int k, int lane)
{
    const double target = S[k];
    const double top = S[arg[k - 1]];
    int pos;
    if (VARIANT == 0 ? (top < target) : (top <= target)) {
        pos = k;                                  // new best (or tie, local): append
    } else if (VARIANT != 0) {
        // smallorequal2target_1d_point(...) + 1 (:27389) == upper bound
        pos = vm_warp_count<true>(S, arg, 0, k, target, lane);
    } else {
        // insertpoint_score (:19369-19387): replay its binary search knowing
        // a = #{S < target}, b = #{S <= target}
        const int b = vm_warp_count<true>(S, arg, 0, k, target, lane);
        if (b == 0) {
            pos = 0;
        } else {
            const int a = vm_warp_count<false>(S, arg, 0, b, target, lane);
            if (a == b) {
                pos = a;
            } else {
                int i = 0, j = k;
                pos = -1;
                while (i < j) {
                    const int mid = (i + j) >> 1;
                    if (mid < a) i = mid + 1;
                    else if (mid >= b) j = mid;
                    else { pos = mid + 1; break; }
                }
                if (pos < 0) pos = j;
            }
        }
    }
    // shift arg[pos..k) up by one, highest chunk first
    for (int hi = k; hi > pos; hi -= 32) {
        const int idx = hi - 1 - lane;
        int v = 0;
        if (idx >= pos) v = arg[idx];
        __syncwarp();
        if (idx >= pos) arg[idx + 1] = v;
    }
    __syncwarp();
    if (lane == 0) arg[pos] = k;
    __syncwarp();
}

// One bulk asynchronous copy (TMA engine, `cp.async.bulk`) of `bytes` (a multiple of 16) from global to this CTA's shared
// memory, completion signalled on an mbarrier; issued by one thread, awaited by the whole warp.
__device__ __forceinline__ void vm_bulk_load(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *mbar, int lane)
{
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst), bar = (unsigned)__cvta_generic_to_shared(mbar);
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(gmem_src),
                     "r"(bytes), "r"(bar)
                     : "memory");
    }
    __syncwarp();
    unsigned done = 0;
    while (!done) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar) : "memory");
    }
}

// SMEM: 0 = S / S_arg in the (L2-resident) output arrays; 1 = S / S_arg in shared memory; 2 = the read's anchors too, staged
// by one TMA bulk copy (the classes of the longest reads: a launch lasts as long as its slowest read, and every anchor of
// that read otherwise pays an L1 / L2 round trip for a[i] and one for each predecessor a[j] on the dependent path)
template <int VARIANT, int SMEM>
__global__ void __launch_bounds__(32) vm_chain_exact_kernel(VmChainArgs A, const int *__restrict__ read_ids, int cap)
{
    extern __shared__ __align__(16) unsigned char vm_smem[];
    const int lane = threadIdx.x;
    const int rid = read_ids[blockIdx.x];
    const long long base = A.off[rid];
    const int n = A.cnt[rid];
    const VmAnchor *__restrict__ a = A.anchors + base;

    double *gcl = (double *)vm_smem;
    float *rgl = (float *)(gcl + VM_GCL_MAX);
    double *S;
    int32_t *arg;
    if (SMEM) {
        S = (double *)(rgl + VM_RGL_MAX);
        arg = (int32_t *)(S + cap);
        if (SMEM == 2 && n > 0) {
            VmAnchor *sa = (VmAnchor *)(arg + cap);                       // 16-byte aligned: tables 1 KB, cap a multiple of 4
            unsigned long long *mbar = (unsigned long long *)(sa + cap);
            vm_bulk_load(sa, A.anchors + base, (unsigned)n * (unsigned)sizeof(VmAnchor), mbar, lane);
            a = sa;
        }
    } else {
        S = A.S + base;
        arg = A.S_arg + base;
    }
    constexpr bool GLOBAL = VARIANT == 0 || VARIANT >= 3;
    constexpr int INS = VARIANT >= 3 ? 0 : VARIANT;             // S_arg insertion rule (insertpoint_score for the read-start ordered DPs)
    for (int t = lane; t <= A.maxdiff && t < VM_GCL_MAX; t += 32) gcl[t] = A.gapcost_list[t];
    if (VARIANT == 1 || VARIANT == 2 || VARIANT == 4)
        for (int t = lane; t < A.n_rg && t < VM_RGL_MAX; t += 32) rgl[t] = A.rgcost[t];
    __syncwarp();
    if (n <= 0) {
        if (lane == 0) { A.gmax[rid] = -1; A.opcount[rid] = 0; }
        return;
    }

    VmScoreCtx c;
    c.skipcost = A.skipcost; c.maxdiff = A.maxdiff; c.maxgap = A.maxgap;
    c.gcl = gcl; c.rgl = rgl; c.extra = A.extra; c.extra_size = A.extra_size;
    c.log2cache = A.log2cache; c.log2cache_size = A.log2cache_size;

    int32_t *P = A.P + base;
    const VmAnchor a0 = a[0];
    int prekey = GLOBAL ? a0.x : a0.x + a0.l;
    if (VARIANT == 0) {
        // coverage of the first read position (:24865-24876): run length, capped at 20
        const bool eq = lane < n && a[lane].x == a0.x;
        const unsigned m = __ballot_sync(VM_FULL, eq);
        int run = (m == VM_FULL) ? 32 : (__ffs(~m) - 1);
        if (run > 20) run = 20;
        c.skipcost = A.skipcost + (double)run;
        c.maxdiff = A.maxdiff - run > 10 ? A.maxdiff - run : 10;
    }
    int testspace_en = 1;
    double g_max_scores = (double)a0.l;
    int g_max_index = 0;
    int i_first = 1;
    const int pre_n = VARIANT >= 3 ? A.pre_n[rid] : 0;
    if (pre_n > 0) {
        // carried anchors in front of the batch (:21713-21718): their scores / negated back-pointers are already in
        // S / P, only the first of them is in the test space until the read position first advances
        if (SMEM)
            for (int t = lane; t < pre_n; t += 32) S[t] = A.S[base + t];
        if (lane == 0) arg[0] = 0;
        g_max_scores = A.head[3 * rid];
        g_max_index = (int)A.head[3 * rid + 1];
        prekey = (int)A.head[3 * rid + 2];
        i_first = pre_n;
    } else if (lane == 0) { arg[0] = 0; S[0] = (double)a0.l; P[0] = VM_NOPRE; }
    __syncwarp();
    long long opcount = 0;
    long long result = 0;
    bool bailed = false;

    for (int i = i_first; i < n; ++i) {
        const VmAnchor ai = a[i];
        const int key = GLOBAL ? ai.x : ai.x + ai.l;
        if (prekey < key) {
            if (GLOBAL) {
                // (the second-round linked DP, variant 4, has no bail-out)
                if (VARIANT != 4 && ((double)opcount / (double)i) > (double)A.max_factor) { bailed = true; result = -1; break; }
            } else {
                if (opcount > 100000 && ((double)opcount / (double)prekey) > 1000.0) { bailed = true; result = -2; break; }
            }
            for (int k = testspace_en; k < i; ++k) vm_insert_one<INS>(S, arg, k, lane);
            testspace_en = i;
            if (VARIANT == 0) {
                const bool eq = (i + lane) < n && a[i + lane].x == ai.x;
                const unsigned m = __ballot_sync(VM_FULL, eq);
                int run = (m == VM_FULL) ? 32 : (__ffs(~m) - 1);
                if (run > 20) run = 20;
                c.skipcost = A.skipcost + (double)run;
                c.maxdiff = A.maxdiff - run > 10 ? A.maxdiff - run : 10;
            }
            prekey = key;
        }
        double max_scores = (double)ai.l;
        int pre_index = VM_NOPRE;
        const double li = (double)ai.l;
        for (int top = testspace_en; top > 0; top -= 32) {
            const int q = top - 1 - lane;
            const bool valid = q >= 0;
            int j = 0;
            double Sj = 0.0, t = -CUDART_INF;
            bool skip = false;
            if (valid) {
                j = arg[q];
                Sj = S[j];
                const VmAnchor aj = a[j];
                t = vm_pair_score<VARIANT>(c, ai, aj, Sj, skip);
            }
            // ---- break position and winner of the chunk ----
            // The reference visits the lanes in order, keeps a running maximum `exc`, stops at the first lane whose S can
            // no longer win (S <= exc - l_i), and the first strictly better candidate wins.  Fast path: one warp-wide
            // arg-max (two REDUX on an order-preserving 64-bit key + a ballot).  With X = max(max_scores, M), M the
            // chunk's maximum at lane bl, every lane behind bl has exc = X exactly and every lane has exc <= X, and S is
            // non-increasing along the lanes -- so the lanes that stop under X form a suffix starting at first', and when
            // first' > bl it is the reference's break position exactly (lanes before it do not stop even under X, lanes
            // from it on see exc = X).  Only when the stop would fall at or before the winner is the exact running maximum
            // needed (prefix-max ladder, as before).
            unsigned long long ukey = (unsigned long long)__double_as_longlong(t);
            ukey = (ukey >> 63) ? ~ukey : (ukey | 0x8000000000000000ULL);
            const unsigned khi = (unsigned)(ukey >> 32), klo = (unsigned)ukey;
            const unsigned mh = __reduce_max_sync(VM_FULL, khi);
            const unsigned ml = __reduce_max_sync(VM_FULL, khi == mh ? klo : 0u);
            const unsigned wmask = __ballot_sync(VM_FULL, khi == mh && klo == ml);
            int bl = __ffs(wmask) - 1;                                // lowest lane attaining the maximum
            unsigned long long mkey = (unsigned long long)mh << 32 | ml;
            mkey = (mkey >> 63) ? (mkey & 0x7fffffffffffffffULL) : ~mkey;
            double bt = __longlong_as_double((long long)mkey);
            const double X = bt > max_scores ? bt : max_scores;
            const unsigned vm = __ballot_sync(VM_FULL, valid);
            const int nvalid = __popc(vm);
            unsigned bm;
            if (GLOBAL) bm = __ballot_sync(VM_FULL, valid && !(Sj > (X - li)));
            else bm = __ballot_sync(VM_FULL, valid && (Sj < (X - li)));
            int first = bm ? (__ffs(bm) - 1) : nvalid;
            if (first <= bl) {
                // exact running max the reference would hold just before visiting each lane's j
                double inc = t;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const double o = __shfl_up_sync(VM_FULL, inc, d);
                    if (lane >= d && o > inc) inc = o;
                }
                double exc = __shfl_up_sync(VM_FULL, inc, 1);
                if (lane == 0 || !(exc > max_scores)) exc = max_scores;
                bool brk;
                if (GLOBAL) brk = valid && !(Sj > (exc - li));         // :24949 / else break :25003
                else brk = valid && (Sj < (exc - li));                 // :27413
                bm = __ballot_sync(VM_FULL, brk);
                first = bm ? (__ffs(bm) - 1) : nvalid;
                // lanes at/after the break are never evaluated: arg-max over the lanes before it, lowest lane wins ties
                double tt = lane >= first ? -CUDART_INF : t;
                bt = tt;
                bl = lane;
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) {
                    const double ot = __shfl_xor_sync(VM_FULL, bt, d);
                    const int ol = __shfl_xor_sync(VM_FULL, bl, d);
                    if (ot > bt || (ot == bt && ol < bl)) { bt = ot; bl = ol; }
                }
            }
            if (GLOBAL) opcount += first;                            // counted inside the if (:24951)
            else opcount += bm ? first + 1 : nvalid;                 // counted before the test (:27410)
            const int bj = __shfl_sync(VM_FULL, j, bl);
            if (bt > max_scores) { max_scores = bt; pre_index = bj; }
            if (bm) break;
        }
        if (lane == 0) { S[i] = max_scores; P[i] = pre_index; }
        __syncwarp();
        if (max_scores > g_max_scores) { g_max_scores = max_scores; g_max_index = i; }
    }
    if (!bailed) {
        for (int k = testspace_en; k < n; ++k) vm_insert_one<INS>(S, arg, k, lane);
        result = g_max_index;
        if (SMEM) {
            double *So = A.S + base;
            int32_t *Ao = A.S_arg + base;
            for (int t = lane; t < n; t += 32) { So[t] = S[t]; Ao[t] = arg[t]; }
        }
    }
    if (lane == 0) { A.gmax[rid] = result; A.opcount[rid] = opcount; }
}

template <int VARIANT>
static int vm_launch_exact_v(const VmChainArgs &args, const int *ids, int n_ids, int cap, bool use_smem,
                             cudaStream_t stream)
{
    if (n_ids <= 0) return 0;
    // Where S / S_arg (and the anchors) of a read live decides both the latency of one DP step and how many reads an SM
    // holds (one warp per read): everything in shared memory is the fastest step but, for the long-read classes, leaves
    // one or two warps per SM.  That is right for the handful of long reads of an ONT batch (the launch lasts as long
    // as its slowest read) and wrong when the whole batch is long (HiFi, -k 19: every read has ~5 000 anchors; 148
    // resident warps took 132 ms per 7 500 reads).  Pick the placement with the smallest estimated duration:
    // waves of resident warps x relative step latency.
    static const int n_sm = [] { int dev = 0, v = 148; cudaGetDevice(&dev); cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev); return v; }();
    static const int forced = getenv("VM_CHAIN_MODE") ? atoi(getenv("VM_CHAIN_MODE")) : -1;      // experiment knob
    const size_t tables = VM_GCL_MAX * 8 + VM_RGL_MAX * 4;
    const size_t smem_of[3] = {tables, tables + (size_t)cap * 12, tables + (size_t)cap * 28 + 16};
    const bool ok[3] = {true, use_smem, use_smem && cap >= VM_CHAIN_ANCHOR_SMEM_MIN && cap <= VM_CHAIN_ANCHOR_SMEM_MAX};
    const double lat[3] = {2.2, 1.25, 1.0};
    int mode = 0;
    double best = 0;
    for (int m = 0; m < 3; ++m) {
        if (!ok[m]) continue;
        size_t per_sm = (size_t)233472 / (smem_of[m] + 1024);
        if (per_sm > 32) per_sm = 32;
        if (per_sm < 1) per_sm = 1;
        const double waves = (double)n_ids / (double)(per_sm * (size_t)n_sm);
        const double t = (waves > 1.0 ? waves : 1.0) * lat[m];
        if (m == 0 || t <= best) { best = t; mode = m; }
    }
    if (forced >= 0 && forced <= 2 && ok[forced]) mode = forced;
    if (mode == 2) {
        // anchors staged in shared memory too (28 B per anchor + the mbarrier)
        vm_smem_optin(vm_chain_exact_kernel<VARIANT, 2>);
        vm_chain_exact_kernel<VARIANT, 2><<<n_ids, 32, smem_of[2], stream>>>(args, ids, cap);
    } else if (mode == 1) {
        vm_smem_optin(vm_chain_exact_kernel<VARIANT, 1>);
        vm_chain_exact_kernel<VARIANT, 1><<<n_ids, 32, smem_of[1], stream>>>(args, ids, cap);
    } else {
        vm_chain_exact_kernel<VARIANT, 0><<<n_ids, 32, smem_of[0], stream>>>(args, ids, cap);
    }
    return 1;
}

int vm_launch_chain_exact(int variant, const VmChainArgs &args, const int *read_ids_dev, int n_ids,
                          int cap, bool use_smem, cudaStream_t stream)
{
    switch (variant) {
    case 0: return vm_launch_exact_v<0>(args, read_ids_dev, n_ids, cap, use_smem, stream);
    case 1: return vm_launch_exact_v<1>(args, read_ids_dev, n_ids, cap, use_smem, stream);
    case 3: return vm_launch_exact_v<3>(args, read_ids_dev, n_ids, cap, use_smem, stream);
    case 4: return vm_launch_exact_v<4>(args, read_ids_dev, n_ids, cap, use_smem, stream);
    default: return vm_launch_exact_v<2>(args, read_ids_dev, n_ids, cap, use_smem, stream);
    }
}

// ---------------------------------------------------------------------------
// heuristic ("fast") DP: `_d_fast_all` :25033-25339 and the local `_fast`
// fall-backs :26938-27303, :27891-28248.  The algorithm is defined by its
// integer-score buckets and closest-diagonal probe, so it is replayed step by
// step: warp-uniform scalar control flow (every lane computes the same values,
// lane 0 stores), with the S_arg_i tail shift done by all lanes.
// scratch per read (int64): Si[n], target[n], count[cnt]
// ---------------------------------------------------------------------------
__device__ __forceinline__ int vm_ips_distance(const long long *Si, long long target, int k, const int32_t *arg,
                                               long long tdist, const long long *dist)
{
    int i = 0, j = k;
    if (Si[arg[0]] > target) return 0;
    if (Si[arg[k - 1]] < target) return k;
    while (i < j) {
        const int mid = (i + j) >> 1;
        const long long now = Si[arg[mid]];
        if (now < target) i = mid + 1;
        else if (now > target) j = mid;
        else {
            const long long nd = dist[arg[mid]];
            if (nd < tdist) i = mid + 1;
            else if (nd > tdist) j = mid;
            else return mid + 1;
        }
    }
    return j;
}

__device__ __forceinline__ int vm_closest_distance(long long tdist, const long long *dist, const int32_t *arg,
                                                   int st, int en)
{
    int i = st, j = en;
    if (dist[arg[i]] >= tdist) return i;
    if (dist[arg[j - 1]] <= tdist) return j - 1;
    while (i < j) {
        const int mid = (i + j) >> 1;
        const long long nd = dist[arg[mid]];
        if (nd < tdist) i = mid + 1;
        else if (nd > tdist) j = mid;
        else return mid;
    }
    if ((tdist - dist[arg[j - 1]]) < (dist[arg[j]] - tdist)) return j - 1;
    return j;
}

__device__ __forceinline__ void vm_shift_insert(int32_t *arg, int pos, int k, int lane)
{
    for (int hi = k; hi > pos; hi -= 32) {
        const int idx = hi - 1 - lane;
        int v = 0;
        if (idx >= pos) v = arg[idx];
        __syncwarp();
        if (idx >= pos) arg[idx + 1] = v;
    }
    __syncwarp();
    if (lane == 0) arg[pos] = k;
    __syncwarp();
}

template <int VARIANT>
__global__ void __launch_bounds__(32) vm_chain_fast_kernel(VmChainArgs A, int fast_t, const int *__restrict__ read_ids,
                                                           long long *__restrict__ scratch,
                                                           const int64_t *__restrict__ scratch_off)
{
    extern __shared__ __align__(16) unsigned char vm_smem[];
    const int lane = threadIdx.x;
    const int rid = read_ids[blockIdx.x];
    const long long base = A.off[rid];
    const int n = A.cnt[rid];
    const VmAnchor *__restrict__ a = A.anchors + base;
    double *gcl = (double *)vm_smem;
    float *rgl = (float *)(gcl + VM_GCL_MAX);
    constexpr bool GLOBAL = VARIANT == 0 || VARIANT == 3;      // 3: asm mode's linked twin (mammap_asm.py:21872-22158)
    for (int t = lane; t <= A.maxdiff && t < VM_GCL_MAX; t += 32) gcl[t] = A.gapcost_list[t];
    if (!GLOBAL)
        for (int t = lane; t < A.n_rg && t < VM_RGL_MAX; t += 32) rgl[t] = A.rgcost[t];
    __syncwarp();
    if (n <= 0) {
        if (lane == 0) A.gmax[rid] = -1;
        return;
    }
    double *S = A.S + base;
    int32_t *P = A.P + base;
    int32_t *arg = A.S_arg + base;
    const long long sbase = scratch_off[blockIdx.x];
    const long long cnt_size = scratch_off[blockIdx.x + 1] - sbase - 2LL * n;
    long long *Si = scratch + sbase;
    long long *target = Si + n;
    long long *count = target + n;

    VmScoreCtx c;
    c.skipcost = A.skipcost; c.maxdiff = A.maxdiff; c.maxgap = A.maxgap;
    c.gcl = gcl; c.rgl = rgl; c.extra = A.extra; c.extra_size = A.extra_size;
    c.log2cache = A.log2cache; c.log2cache_size = A.log2cache_size;

    const int lastpos = a[n - 1].x;
    const long long readlength = (long long)lastpos + 1000;
    for (int t = lane; t < n; t += 32) {
        const VmAnchor v = a[t];
        if (v.s == 1) target[t] = (long long)v.y - v.x + readlength;
        else target[t] = -((long long)v.y + v.x + readlength);
    }
    for (long long t = lane; t < cnt_size; t += 32) count[t] = 0;
    __syncwarp();

    const VmAnchor a0 = a[0];
    int prekey = GLOBAL ? a0.x : a0.x + a0.l;
    int testspace_en_i = 1;
    double g_max_scores = (double)a0.l;
    int g_max_index = 0;
    long long max_score_i = 0;
    int i_first = 1;
    const int pre_n = VARIANT == 3 ? A.pre_n[rid] : 0;
    if (pre_n > 0) {
        // carried prefix (:21903-21914): S / P already hold pre_S / pre_P; only its first anchor is in the test space
        for (int t = lane; t < pre_n; t += 32) Si[t] = (long long)S[t];
        __syncwarp();
        max_score_i = Si[0];
        if (lane == 0) {
            arg[0] = 0;
            if (max_score_i >= 0 && max_score_i < cnt_size) count[max_score_i] = 1;
        }
        g_max_scores = A.head[3 * rid];
        g_max_index = (int)A.head[3 * rid + 1];
        prekey = (int)A.head[3 * rid + 2];
        i_first = pre_n;
    } else if (lane == 0) {
        arg[0] = 0; S[0] = (double)a0.l; Si[0] = a0.l; P[0] = VM_NOPRE;
        if (a0.l < cnt_size) count[a0.l] = 1;
    }
    __syncwarp();

    for (int i = i_first; i < n; ++i) {
        const VmAnchor ai = a[i];
        double max_scores = (double)ai.l;
        int pre_index = VM_NOPRE;
        const int key = GLOBAL ? ai.x : ai.x + ai.l;
        if (prekey < key) {
            for (int k = testspace_en_i; k < i; ++k) {
                const long long sk = Si[k];
                if (lane == 0 && sk < cnt_size) count[sk] += 1;
                if (sk > max_score_i) max_score_i = sk;
                const int loc = vm_ips_distance(Si, sk, k, arg, target[k], target);
                vm_shift_insert(arg, loc, k, lane);
            }
            testspace_en_i = i;
            if (VARIANT == 0) {
                const bool eq = (i + lane) < n && a[i + lane].x == ai.x;
                const unsigned m = __ballot_sync(VM_FULL, eq);
                int run = (m == VM_FULL) ? 32 : (__ffs(~m) - 1);
                if (run > 20) run = 20;
                c.skipcost = A.skipcost + (double)run;
                c.maxdiff = A.maxdiff - run > 10 ? A.maxdiff - run : 10;
            }
            prekey = key;
        }
        long long c_score_i = max_score_i;
        int en_loc = testspace_en_i;
        const long long f_kmersize = (long long)ai.l + 1;
        while ((double)c_score_i > (max_scores - (double)f_kmersize)) {
            const long long now_count = (c_score_i >= 0 && c_score_i < cnt_size) ? count[c_score_i] : 0;
            if (now_count == 0) { --c_score_i; continue; }
            const int st_loc = en_loc - (int)now_count;
            if (now_count > fast_t) {
                const int j = arg[vm_closest_distance(target[i], target, arg, st_loc, en_loc)];
                bool skip;
                const double t = vm_pair_score<VARIANT>(c, ai, a[j], S[j], skip);
                if (!skip && t > max_scores) { max_scores = t; pre_index = j; }
            } else {
                for (int q = en_loc - 1; q >= st_loc; --q) {
                    const int j = arg[q];
                    bool skip;
                    const double t = vm_pair_score<VARIANT>(c, ai, a[j], S[j], skip);
                    if (!skip && t > max_scores) { max_scores = t; pre_index = j; }
                }
            }
            en_loc = st_loc;
            --c_score_i;
        }
        if (lane == 0) { S[i] = max_scores; Si[i] = (long long)max_scores; P[i] = pre_index; }
        __syncwarp();
        if (max_scores > g_max_scores) { g_max_scores = max_scores; g_max_index = i; }
    }
    for (int k = testspace_en_i; k < n; ++k) {
        const long long sk = Si[k];
        if (lane == 0 && sk < cnt_size) count[sk] += 1;
        const int loc = vm_ips_distance(Si, sk, k, arg, target[k], target);
        vm_shift_insert(arg, loc, k, lane);
    }
    if (lane == 0) A.gmax[rid] = g_max_index;
}

int vm_launch_chain_fast(int variant, const VmChainArgs &args, int fast_t, const int *read_ids_dev,
                         int n_ids, long long *scratch_i64, const int64_t *scratch_off,
                         cudaStream_t stream)
{
    if (n_ids <= 0) return 0;
    size_t smem = VM_GCL_MAX * 8 + VM_RGL_MAX * 4;
    switch (variant) {
    case 0: vm_chain_fast_kernel<0><<<n_ids, 32, smem, stream>>>(args, fast_t, read_ids_dev, scratch_i64, scratch_off); break;
    case 1: vm_chain_fast_kernel<1><<<n_ids, 32, smem, stream>>>(args, fast_t, read_ids_dev, scratch_i64, scratch_off); break;
    case 3: vm_chain_fast_kernel<3><<<n_ids, 32, smem, stream>>>(args, fast_t, read_ids_dev, scratch_i64, scratch_off); break;
    default: vm_chain_fast_kernel<2><<<n_ids, 32, smem, stream>>>(args, fast_t, read_ids_dev, scratch_i64, scratch_off); break;
    }
    return 1;
}
