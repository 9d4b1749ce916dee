// mma_tf32.cuh -- warp-level m16n8k8 TF32 tensor-core products with the 3xTF32 split
// (x = hi + lo, both TF32; a.b ~= a_lo.b_hi + a_hi.b_lo + a_hi.b_hi, fp32 accumulate), which keeps
// fp32-level accuracy (the dropped lo.lo term and the split residuals are ~2^-20 relative) while moving the two D x W
// contractions of the lin backward off the FMA pipe.
//
// Fragment layout of mma.sync.m16n8k8 (g = lane >> 2, t = lane & 3):
//   A (16 x 8, row):  a0 (g, t)   a1 (g + 8, t)   a2 (g, t + 4)   a3 (g + 8, t + 4)
//   B ( 8 x 8, col):  b0 (k = t, n = g)           b1 (k = t + 4, n = g)
//   C (16 x 8):       c0 (g, 2t)  c1 (g, 2t + 1)  c2 (g + 8, 2t)  c3 (g + 8, 2t + 1)
#pragma once
#include <stdint.h>

namespace gdn {

// x = hi + lo with hi = x rounded to TF32 (round-half-away on the 13 dropped mantissa bits: an integer
// add and a mask; cvt.rna.tf32.f32 is emulated in several instructions on this architecture) and
// lo = x - hi, exact in fp32 and handed over as it is: the tensor core reads only the TF32 bits of an
// operand, i.e. truncates lo to its own top 10 mantissa bits.  |x - hi - tf32(lo)| <= 2^-21 |x|, and
// because hi is rounded (not truncated) lo has either sign, so the residual does not accumulate as a bias
// over a long dot product (a truncating split measurably did: 1.7e-4 on the smallest gradients).
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
// round-to-nearest split (both parts proper TF32 values): for operands that are split rarely
__device__ __forceinline__ void split_tf32_rn(float x, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(x - __uint_as_float(hi)));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// c += a.b with both operands split; the small terms go first
__device__ __forceinline__ void mma_3xtf32(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                           uint32_t bh0, uint32_t bh1, uint32_t bl0, uint32_t bl1) {
    mma_tf32(c, al, bh0, bh1);
    mma_tf32(c, ah, bl0, bl1);
    mma_tf32(c, ah, bh0, bh1);
}
// four 8 x 4 fp32 sub-matrices (8 rows of 16 bytes each) -> one register each: lane (g, t) receives
// element (g, t) of matrix j in r[j].  Lane l supplies the row address of matrix l >> 3, row l & 7.
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* row_ptr) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(row_ptr);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

}  // namespace gdn
