// Warp-cooperative, bit-exact replay of numba's argsort (numba/misc/quicksort.py).
//
// The reference sorts anchors with numba's np.argsort / list.sort(key=) -- an
// UNSTABLE quicksort (median-of-3 of low/mid/high, pivot stashed at `high`, Hoare
// scans with strict `<`, insertion sort below 15 elements).  The order it leaves
// equal keys in feeds the chaining DP's tie-breaking, so the permutation has to
// be reproduced exactly (mammap_clrnano.py:23572, 28585, 23103, 23183, 23652).
//
// The partition is NOT replayed step by step.  The k-th swap of a Hoare scan
// exchanges the k-th element from the left that is not `< pivot` with the k-th
// element from the right that is not `> pivot`, for as long as the former lies
// left of the latter, and both stop lists depend only on the ORIGINAL contents
// of the segment.  So one warp (1) compacts both stop lists with ballots,
// (2) performs all swaps of the partition concurrently, and (3) derives the
// final pivot slot in closed form.  Segments are independent, so the order in
// which they are processed is irrelevant to the result.  Sub-15 segments are
// finished with a stable rank sort (= the insertion sort's result).
#pragma once
#include "vm_common.cuh"

template <typename K>
__device__ __forceinline__ int vm_warp_hoare(const K *keys, int *R, int *Lpos, int *Rpos, int low, int high,
                                             K pivot, int lane)
{
    const int len = high - low;   // candidate slots low .. high-1 (pivot is parked at `high`)
    const unsigned ltmask = (1u << lane) - 1u;
    int nL = 0, nR = 0;
    for (int b = 0; b < len; b += 32) {
        const int p = low + b + lane;
        const bool ge = (b + lane) < len && !(keys[R[p]] < pivot);
        const unsigned m = __ballot_sync(VM_FULL, ge);
        if (ge) Lpos[nL + __popc(m & ltmask)] = p;
        nL += __popc(m);
    }
    for (int b = 0; b < len; b += 32) {
        const int p = high - 1 - b - lane;
        const bool le = (b + lane) < len && !(pivot < keys[R[p]]);
        const unsigned m = __ballot_sync(VM_FULL, le);
        if (le) Rpos[nR + __popc(m & ltmask)] = p;
        nR += __popc(m);
    }
    __syncwarp();
    int K_ = 0;
    const int kmax = nL < nR ? nL : nR;
    for (int b = 0; b < kmax; b += 32) {
        const int k = b + lane;
        int lp = 0, rp = 0;
        bool ok = false;
        if (k < kmax) {
            lp = Lpos[k];
            rp = Rpos[k];
            ok = lp < rp;
        }
        const unsigned m = __ballot_sync(VM_FULL, ok);
        if (ok) {
            const int t = R[lp];
            R[lp] = R[rp];
            R[rp] = t;
        }
        const int c = __popc(m);
        K_ += c;
        if (c < 32) break;
    }
    __syncwarp();
    int i_final;
    if (K_ < nL && (K_ == 0 || Lpos[K_] < Rpos[K_ - 1])) i_final = Lpos[K_];
    else if (K_ > 0) i_final = Rpos[K_ - 1];
    else i_final = high;
    return i_final;
}

template <typename K>
__device__ __forceinline__ void vm_warp_small_sort(const K *keys, int *R, int low, int high, int lane)
{
    const int cnt = high - low + 1;
    if (cnt < 2) return;
    const int id = lane < cnt ? R[low + lane] : 0;
    const K key = keys[id];
    int rank = 0;
    for (int m = 0; m < cnt; ++m) {
        const K km = __shfl_sync(VM_FULL, key, m);
        if (km < key || (!(key < km) && m < lane)) ++rank;
    }
    __syncwarp();
    if (lane < cnt) R[low + rank] = id;
    __syncwarp();
}

// keys[id] for id in [0,n); R receives the permutation; Lpos/Rpos: n ints of scratch each.
template <typename K>
__device__ __forceinline__ void vm_warp_argsort_replay(const K *keys, int *R, int *Lpos, int *Rpos, int n, int lane)
{
    for (int t = lane; t < n; t += 32) R[t] = t;
    __syncwarp();
    if (n < 2) return;
    int stack_lo[64], stack_hi[64];
    int sp = 1;
    stack_lo[0] = 0;
    stack_hi[0] = n - 1;
    while (sp > 0) {
        --sp;
        int low = stack_lo[sp], high = stack_hi[sp];
        while (high - low >= 15) {
            const int mid = (low + high) >> 1;
            // median of three {low, mid, high}: same compare/swap sequence as the reference
            int rl = R[low], rm = R[mid], rh = R[high];
            int t;
            if (keys[rm] < keys[rl]) { t = rl; rl = rm; rm = t; }
            if (keys[rh] < keys[rm]) { t = rh; rh = rm; rm = t; }
            if (keys[rm] < keys[rl]) { t = rl; rl = rm; rm = t; }
            const K pivot = keys[rm];
            __syncwarp();
            // park the pivot at `high`
            if (lane == 0) { R[low] = rl; R[mid] = rh; R[high] = rm; }
            __syncwarp();
            const int i = vm_warp_hoare<K>(keys, R, Lpos, Rpos, low, high, pivot, lane);
            if (lane == 0) { const int ti = R[i]; R[i] = R[high]; R[high] = ti; }
            __syncwarp();
            if (high - i > i - low) {
                if (high > i && sp < 64) { stack_lo[sp] = i + 1; stack_hi[sp] = high; ++sp; }
                high = i - 1;
            } else {
                if (i > low && sp < 64) { stack_lo[sp] = low; stack_hi[sp] = i - 1; ++sp; }
                low = i + 1;
            }
        }
        vm_warp_small_sort<K>(keys, R, low, high, lane);
    }
}
