// CUDA side of the device-resident extension stage: generic launchers for the per-read functors of vm_dgrun.hpp and
// the warp-collective z-drop extension they call.  The per-read logic itself is in vm_dglue.hpp.
#pragma once
#include "vm_dgrun.hpp"
#include "vm_extend.cuh"

static_assert(sizeof(vmd::A32) == sizeof(VmAnchor), "anchor layout");
static_assert(sizeof(vmd::Job) == sizeof(VmAlnJobDev), "job layout");
static_assert(sizeof(vmd::Spec) == sizeof(VmSeqSpec), "sequence spec layout");
static_assert(sizeof(vmd::Rec) == sizeof(vm_record), "record layout");

// the reference's k_cigar(2,-4,4,4,4,4,bw=100,zdropvalue=50) for the job (target, query) of read `read`
struct VmDgExt {
    VmSeqSources S;
    VmExtSmem *M;
    int32_t read;
    __host__ __device__ void operator()(const vmd::Spec &t, const vmd::Spec &q, int32_t &q_e, int32_t &t_e)
    {
#if defined(__CUDA_ARCH__)
        VmSeqSpec ts, qs;
        ts.lo = t.lo; ts.len = t.len; ts.src = t.src; ts.reverse = t.reverse; ts.comp = t.comp;
        qs.lo = q.lo; qs.len = q.len; qs.src = q.src; qs.reverse = q.reverse; qs.comp = q.comp;
        const VmSeqView T = vm_view(S, ts, read), Q = vm_view(S, qs, read);
        int qe, te;
        vm_extend_warp(T, Q, *M, qe, te);
        q_e = qe; t_e = te;
#else
        (void)t; (void)q; q_e = 0; t_e = 0;
#endif
    }
};

template <typename F>
__global__ void __launch_bounds__(64) vm_dg_item_kernel(int64_t n, F f)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) f(t);
}

// one warp per item; every lane runs the functor with the same data (see vm_dgrun.hpp)
template <typename F>
__global__ void __launch_bounds__(128) vm_dg_warp_kernel(int64_t n, F f, VmSeqSources S)
{
    __shared__ VmExtSmem M[4];
    const int warp = threadIdx.x >> 5;
    const int64_t t = (int64_t)blockIdx.x * 4 + warp;
    if (t >= n) return;
    VmDgExt ext;
    ext.S = S; ext.M = &M[warp]; ext.read = 0;
    f(t, ext, (threadIdx.x & 31) == 0);
}

template <typename F>
__global__ void __launch_bounds__(128) vm_dg_write_kernel(const int32_t *__restrict__ ids, int64_t n, F f)
{
    const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t < n) f.run(ids[t], threadIdx.x & 31, 32);
}
