// Cell update of the global dual-affine fill, shared by the full-matrix kernel (vm_fill.cu) and the banded
// kernel (vm_fillb.cu): ksw2's u/v/x/y difference recurrences in fp16, two jobs per half2 lane.
#pragma once
#include "vm_align.cuh"
#include <cuda_fp16.h>

namespace {

struct VmGapPar2 {
    int match, mismatch, q1, e1, q2, e2;
};
__host__ __device__ constexpr VmGapPar2 vm_fill_par() { return VmGapPar2{2, -4, 4, 2, 24, 1}; }

// H on the boundary row / column after `len` gap bases (0 for len == 0)
__device__ __forceinline__ int vm_hb(int len)
{
    constexpr VmGapPar2 g = vm_fill_par();
    if (len <= 0) return 0;
    const int a = -(g.q1 + g.e1 * len), b = -(g.q2 + g.e2 * len);
    return a > b ? a : b;
}

__device__ __forceinline__ __half2 vm_h2(unsigned bits) { return *reinterpret_cast<__half2 *>(&bits); }
__device__ __forceinline__ unsigned vm_u32(__half2 h) { return *reinterpret_cast<unsigned *>(&h); }
__device__ __forceinline__ __half2 vm_h2i(int v) { return __half2half2(__int2half_rn(v)); }
#define VM_H2C(x) __floats2half2_rn((float)(x), (float)(x))

// fp16 bit pattern of a base code; N (and padding) is NaN so that both the ordered == and the ordered != test
// fail and the substitution score becomes 0
__device__ __forceinline__ unsigned vm_code_half(int c)
{
    return c == 0 ? 0x0000u : c == 1 ? 0x3c00u : c == 2 ? 0x4000u : c == 3 ? 0x4200u : 0x7fffu;
}

// The same codes, high byte only (their low bytes are zero; N = 0x7f00 is still a NaN): two jobs in 16 bits of
// shared memory, spread to the half2 operand with one byte permute
__device__ __forceinline__ unsigned vm_code_byte(int c)
{
    return c == 0 ? 0x00u : c == 1 ? 0x3cu : c == 2 ? 0x40u : c == 3 ? 0x42u : 0x7fu;
}
__device__ __forceinline__ __half2 vm_codes_half2(unsigned two_bytes) { return vm_h2(__byte_perm(two_bytes, 0u, 0x1404)); }

// One cell for both jobs.  in: v = v(i-1,j), x1/x2 = x(i-1,j) from the cell above; u/y1/y2 = from the cell to the
// left.  out: the same quantities for (i,j), and the direction bits of job A in byte 0 / job B in byte 2.
// EQBIT: bit 7 = the bases are equal (the full-matrix kernel's traceback reads it; the banded kernel's compares the
// staged bases itself and saves the instruction).
template <bool EQBIT = true>
__device__ __forceinline__ void vm_cell2(__half2 tc, __half2 qc, __half2 &v, __half2 &x1, __half2 &x2, __half2 &u, __half2 &y1,
                                         __half2 &y2, unsigned &dir)
{
    constexpr VmGapPar2 g = vm_fill_par();
    const unsigned eqm = __heq2_mask(tc, qc);
    const __half2 ne1 = __hne2(tc, qc);
    __half2 z = __hfma2(ne1, VM_H2C(g.mismatch), vm_h2(eqm & (g.match == 2 ? 0x40004000u : 0u)));
    const __half2 a = __hadd2(x1, v), b = __hadd2(y1, u), a2 = __hadd2(x2, v), b2 = __hadd2(y2, u);
    unsigned m, dl;
    m = __hgt2_mask(a, z);  z = __hmax2(z, a);  dl = m & 0x00010001u;
    m = __hgt2_mask(b, z);  z = __hmax2(z, b);  dl = (dl & ~m) | (m & 0x00020002u);
    m = __hgt2_mask(a2, z); z = __hmax2(z, a2); dl = (dl & ~m) | (m & 0x00030003u);
    m = __hgt2_mask(b2, z); z = __hmax2(z, b2); dl = (dl & ~m) | (m & 0x00040004u);
    const __half2 un = __hsub2(z, v), vn = __hsub2(z, u);
    const __half2 one = VM_H2C(1), zero = VM_H2C(0);
    const __half2 t1 = __hsub2(VM_H2C(g.q1), z);          // -(z - q1)
    const __half2 ap = __hfma2_relu(one, a, t1), bp = __hfma2_relu(one, b, t1);
    const __half2 t2 = __hsub2(VM_H2C(g.q2), z);
    const __half2 a2p = __hfma2_relu(one, a2, t2), b2p = __hfma2_relu(one, b2, t2);
    unsigned d = EQBIT ? (dl | (eqm & 0x00800080u)) : dl;
    d |= __hgt2_mask(ap, zero) & 0x00080008u;
    d |= __hgt2_mask(bp, zero) & 0x00100010u;
    d |= __hgt2_mask(a2p, zero) & 0x00200020u;
    d |= __hgt2_mask(b2p, zero) & 0x00400040u;
    x1 = __hsub2(ap, VM_H2C(g.q1 + g.e1));
    y1 = __hsub2(bp, VM_H2C(g.q1 + g.e1));
    x2 = __hsub2(a2p, VM_H2C(g.q2 + g.e2));
    y2 = __hsub2(b2p, VM_H2C(g.q2 + g.e2));
    u = un;
    v = vn;
    dir = d;
}

} // namespace
