// Device counting sort (see vm_devsort.cuh).
#include "vm_devsort.cuh"

namespace {

__global__ void vm_bs_hist_kernel(const int32_t *__restrict__ keys, int n, int32_t *cnt)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int k = keys[j];
    if (k >= 0) atomicAdd(cnt + k, 1);
}

// one block: exclusive scan of cnt[n_key] into start[n_key + 1]; cnt is zeroed (it becomes the scatter cursor)
__global__ void __launch_bounds__(1024) vm_bs_scan_kernel(int32_t *cnt, int n_key, int32_t *start)
{
    __shared__ int32_t part[1024];
    const int t = threadIdx.x;
    const int per = (n_key + 1023) / 1024;
    const int lo = t * per, hi = min(lo + per, n_key);
    int32_t s = 0;
    for (int k = lo; k < hi; ++k) s += cnt[k];
    part[t] = s;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        const int32_t v = t >= d ? part[t - d] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    int32_t run = part[t] - s;      // exclusive
    for (int k = lo; k < hi; ++k) {
        const int32_t c = cnt[k];
        start[k] = run;
        cnt[k] = 0;
        run += c;
    }
    if (t == 1023) start[n_key] = part[1023];
}

__global__ void vm_bs_scatter_kernel(const int32_t *__restrict__ keys, int n, const int32_t *__restrict__ start, int32_t *cursor,
                                     int32_t *order)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int k = keys[j];
    if (k >= 0) order[start[k] + atomicAdd(cursor + k, 1)] = j;
}

} // namespace

int vm_bucket_sort(const int32_t *keys, int n, int n_key, int32_t *start, int32_t *cursor, int32_t *order, cudaStream_t stream)
{
    cudaMemsetAsync(cursor, 0, (size_t)n_key * 4, stream);
    int launches = 0;
    if (n > 0) { vm_bs_hist_kernel<<<(n + 255) / 256, 256, 0, stream>>>(keys, n, cursor); ++launches; }
    vm_bs_scan_kernel<<<1, 1024, 0, stream>>>(cursor, n_key, start);
    ++launches;
    if (n > 0) { vm_bs_scatter_kernel<<<(n + 255) / 256, 256, 0, stream>>>(keys, n, start, cursor, order); ++launches; }
    return launches;
}
