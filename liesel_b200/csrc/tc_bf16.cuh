// BF16 correction passes of the split-precision tensor-core product (shared by tc_probe.cu, which
// validates them stand-alone, and dense_tc.cu):
//   a.b ~= hi(a).hi(b) [kind::tf32]  +  bf16(lo(a)).bf16(b) + bf16(a).bf16(lo(b)) [kind::f16]
// hi(x) = x with the low 13 mantissa bits cleared, lo(x) = x - hi(x) (exact). The corrections are
// 2^-11 of the product, so BF16's 2^-9 relative rounding (round to nearest even) leaves about
// 2^-19 per product, the level of the dropped lo.lo term of the all-TF32 3-pass scheme; kind::f16
// runs at twice the rate (and half the energy) of kind::tf32.
// Shared-memory tiles are K-major with 16 elements per row: fp32 rows are 64 bytes (SWIZZLE_64B:
// 16-byte chunk index ^= (row >> 1) & 3), bf16 rows 32 bytes (SWIZZLE_32B: chunk ^= (row >> 2) & 1).
#pragma once
#include <cuda_bf16.h>

namespace lslb {
namespace tcx {

typedef unsigned long long u64x;

__device__ __forceinline__ u64x desc_k_sw32(const void* smem) {  // rows of 32 bytes, SBO = 8 rows
  const unsigned addr = (unsigned)__cvta_generic_to_shared(smem);
  return (u64x)((addr >> 4) & 0x3FFF) | ((u64x)(256 >> 4) << 32) | ((u64x)1 << 46) | ((u64x)6 << 61);
}
__device__ __forceinline__ unsigned idesc_bf16(int M, int N) {  // D = fp32, A = B = bf16, both K-major
  return (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(unsigned tmem_d, u64x adesc, u64x bdesc, unsigned idesc, unsigned accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ unsigned pack_bf16(float a, float b) {  // a -> low half (lower address)
  unsigned r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ float hi_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ unsigned short bf16_bits(float x) { return __bfloat16_as_ushort(__float2bfloat16_rn(x)); }

// fp32 tile [rows][16] (SWIZZLE_64B) -> hi in place; bf16(x), bf16(x - hi) tiles [rows][16] (SWIZZLE_32B)
__device__ __forceinline__ void split_tile_bf16(float* h, unsigned char* xb, unsigned char* xlb, int rows, int t,
                                                int nthreads) {
  for (int item = t; item < rows * 2; item += nthreads) {
    const int r = item >> 1, half = item & 1;
    const int sw64 = (r >> 1) & 3, sw32 = (r >> 2) & 1;
    float4* row = reinterpret_cast<float4*>(h + r * 16);
    const float4 v0 = row[(2 * half) ^ sw64], v1 = row[(2 * half + 1) ^ sw64];
    const float x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    float hi[8], lo[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      hi[i] = hi_tf32(x[i]);
      lo[i] = x[i] - hi[i];
    }
    row[(2 * half) ^ sw64] = make_float4(hi[0], hi[1], hi[2], hi[3]);
    row[(2 * half + 1) ^ sw64] = make_float4(hi[4], hi[5], hi[6], hi[7]);
    const int off = r * 32 + ((half ^ sw32) << 4);
    *reinterpret_cast<uint4*>(xb + off) =
        make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
    *reinterpret_cast<uint4*>(xlb + off) =
        make_uint4(pack_bf16(lo[0], lo[1]), pack_bf16(lo[2], lo[3]), pack_bf16(lo[4], lo[5]), pack_bf16(lo[6], lo[7]));
  }
}

}  // namespace tcx
}  // namespace lslb
