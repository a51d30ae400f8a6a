// TOOLS LIBRARY (libliesel_b200_bench.so, not part of the product library).
// tcgen05 building block for the dense path (probe + unit under test for the S4 tensor-core kernels):
// D[M=128 x N] (fp32, TMEM) = A[128 x K] . B[N x K]^T with both operands K-major fp32 in global
// memory, loaded by TMA (128-byte swizzle) in K-chunks of 32 floats, multiplied by
// tcgen05.mma.kind::tf32. With `split = 1` the product is the 3-pass split
//   hi(A).hi(B) + lo(A).hi(B) + hi(A).lo(B),   hi(x) = x with the low 13 mantissa bits cleared,
// which restores fp32-level accuracy (operands are split on the fly in shared memory: the split is
// element-wise, hence independent of the swizzled layout).
// One CTA, 4 warps: warp 0 lane 0 = TMA producer + MMA issuer (the probe is not pipelined),
// all 4 warps = epilogue (tcgen05.ld, 32 TMEM lanes per warp).
#include <cuda.h>

#include "common.cuh"
#include "liesel_b200_bench.h"
#include "tc_bf16.cuh"

namespace lslb {

typedef unsigned long long u64t;

__device__ __forceinline__ unsigned tsmem(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, u64t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(tsmem(dst)), "l"(map), "r"(c0), "r"(c1), "r"(tsmem(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_init1(u64t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tsmem(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect(u64t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tsmem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(u64t* bar, unsigned parity) {
  unsigned ok = 0;
  while (!ok) {
    asm volatile(
        "{\n.reg .pred q;\nmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\nselp.u32 %0, 1, 0, q;\n}\n"
        : "=r"(ok)
        : "r"(tsmem(bar)), "r"(parity)
        : "memory");
  }
}
// K-major operand tile, 128-byte swizzle, rows of exactly 128 bytes (32 fp32):
// SBO = 8 rows * 128 B, LBO unused, version 1 (sm_100), layout SWIZZLE_128B (= 2)
// With rows of 64 bytes (16 fp32) the layout is SWIZZLE_64B (= 4) and SBO = 8 rows * 64 B.
template <int PKT>
__device__ __forceinline__ u64t umma_desc_k(const void* smem, unsigned k_byte_offset) {
  const unsigned addr = tsmem(smem) + k_byte_offset;
  u64t d = 0;
  d |= (u64t)((addr >> 4) & 0x3FFF);
  d |= (u64t)((8 * PKT * 4) >> 4) << 32;
  d |= (u64t)1 << 46;
  d |= (u64t)(PKT == 32 ? 2 : 4) << 61;
  return d;
}
// instruction descriptor: D = fp32, A = B = tf32, both K-major, dense
__device__ __forceinline__ unsigned umma_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(unsigned tmem_d, u64t adesc, u64t bdesc, unsigned idesc, unsigned accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(u64t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tsmem(bar))
               : "memory");
}

template <int N, int PK>  // PK = K chunk (floats): one swizzle row of 128 (PK = 32) or 64 bytes (PK = 16)
__global__ void __launch_bounds__(128, 1)
tc_probe_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int K,
                int split, float* __restrict__ D) {
  extern __shared__ unsigned char smraw_[];
  // 128-byte-swizzled tiles must start on a 1024-byte boundary of the shared address space
  unsigned char* smraw = smraw_ + ((1024u - (tsmem(smraw_) & 1023u)) & 1023u);
  float* sA = reinterpret_cast<float*>(smraw);               // [128][32] swizzled
  float* sB = sA + 128 * PK;                                 // [N][32]
  float* sAl = sB + N * PK;                                  // lo parts
  float* sBl = sAl + 128 * PK;
  __shared__ __align__(8) u64t bar_tma, bar_mma;
  __shared__ unsigned tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) {
    mbar_init1(&bar_tma, 1);
    mbar_init1(&bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tsmem(&tmem_base)),
                 "r"(N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const unsigned tmem_d = tmem_base;
  const unsigned idesc = umma_idesc_tf32(128, N);

  const int nchunks = K / PK;
  for (int kc = 0; kc < nchunks; ++kc) {
    if (tid == 0) {
      mbar_expect(&bar_tma, (128 + N) * PK * 4);
      tma_load_2d(sA, &mapA, kc * PK, 0, &bar_tma);
      tma_load_2d(sB, &mapB, kc * PK, 0, &bar_tma);
    }
    mbar_wait_parity(&bar_tma, kc & 1);
    if (split) {  // element-wise hi/lo split in place (layout-agnostic)
      for (int i = tid; i < 128 * PK; i += 128) {
        const float v = sA[i];
        const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        sA[i] = h;
        sAl[i] = v - h;
      }
      for (int i = tid; i < N * PK; i += 128) {
        const float v = sB[i];
        const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        sB[i] = h;
        sBl[i] = v - h;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async proxy (UMMA)
    }
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
      for (int k8 = 0; k8 < PK / 8; ++k8) {
        const unsigned kb = k8 * 32;  // 8 tf32 = 32 bytes
        umma_tf32(tmem_d, umma_desc_k<PK>(sA, kb), umma_desc_k<PK>(sB, kb), idesc, (kc | k8) != 0);
        if (split) {
          umma_tf32(tmem_d, umma_desc_k<PK>(sAl, kb), umma_desc_k<PK>(sB, kb), idesc, 1);
          umma_tf32(tmem_d, umma_desc_k<PK>(sA, kb), umma_desc_k<PK>(sBl, kb), idesc, 1);
        }
      }
      umma_commit(&bar_mma);
    }
    mbar_wait_parity(&bar_mma, kc & 1);  // smem may be overwritten by the next chunk
    asm volatile("tcgen05.fence::after_thread_sync;");
  }
  // epilogue: warp w reads TMEM lanes 32w..32w+31 (= rows), 32 columns at a time
  const int row = tid;
  for (int c0 = 0; c0 < N; c0 += 32) {
    unsigned r[32];
    const unsigned taddr = tmem_d + ((unsigned)(warp * 32) << 16) + (unsigned)c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) D[(size_t)row * N + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(N));
  }
}

// ---------------------------------------------------------------------------------------------
// Mixed-precision 3-pass product: hi(A).hi(B) in TF32 plus the two CORRECTION passes in BF16
//   bf16(lo(A)).bf16(B) + bf16(A).bf16(lo(B))       (kind::f16, twice the rate of kind::tf32)
// accumulated into the same fp32 TMEM accumulator. The corrections are 2^-11 of the result, so
// BF16's 2^-9 relative rounding leaves 2^-19..2^-20 per product. K chunks of 16 floats: fp32 tiles
// have 64-byte rows (SWIZZLE_64B), bf16 tiles 32-byte rows (SWIZZLE_32B). The bf16 tiles are
// produced in shared memory from the fp32 tile at explicit swizzled addresses (this is the code
// the K2 kernel of dense_tc.cu uses for the residual tile; helpers in tc_bf16.cuh).
// ---------------------------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(128, 1)
tc_probe_bf16_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, int K,
                     float* __restrict__ D) {
  constexpr int PK = 16;
  extern __shared__ unsigned char smraw_[];
  unsigned char* smraw = smraw_ + ((1024u - (tsmem(smraw_) & 1023u)) & 1023u);
  float* sA = reinterpret_cast<float*>(smraw);                    // [128][16] fp32, SW64
  float* sB = sA + 128 * PK;                                      // [N][16]
  unsigned char* sAb = reinterpret_cast<unsigned char*>(sB + N * PK);  // [128][16] bf16, SW32
  unsigned char* sAlb = sAb + 128 * 32;
  unsigned char* sBb = sAlb + 128 * 32;
  unsigned char* sBlb = sBb + N * 32;
  __shared__ __align__(8) u64t bar_tma, bar_mma;
  __shared__ unsigned tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init1(&bar_tma, 1);
    mbar_init1(&bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tsmem(&tmem_base)),
                 "r"(N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const unsigned tmem_d = tmem_base;
  const unsigned idesc = umma_idesc_tf32(128, N), idesc16 = tcx::idesc_bf16(128, N);
  for (int kc = 0; kc < K / PK; ++kc) {
    if (tid == 0) {
      mbar_expect(&bar_tma, (128 + N) * PK * 4);
      tma_load_2d(sA, &mapA, kc * PK, 0, &bar_tma);
      tma_load_2d(sB, &mapB, kc * PK, 0, &bar_tma);
    }
    mbar_wait_parity(&bar_tma, kc & 1);
    tcx::split_tile_bf16(sA, sAb, sAlb, 128, tid, 128);
    tcx::split_tile_bf16(sB, sBb, sBlb, N, tid, 128);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
      for (int k8 = 0; k8 < PK / 8; ++k8)
        umma_tf32(tmem_d, umma_desc_k<PK>(sA, k8 * 32), umma_desc_k<PK>(sB, k8 * 32), idesc, (kc | k8) != 0);
      tcx::mma_bf16(tmem_d, tcx::desc_k_sw32(sAlb), tcx::desc_k_sw32(sBb), idesc16, 1);
      tcx::mma_bf16(tmem_d, tcx::desc_k_sw32(sAb), tcx::desc_k_sw32(sBlb), idesc16, 1);
      umma_commit(&bar_mma);
    }
    mbar_wait_parity(&bar_mma, kc & 1);
    asm volatile("tcgen05.fence::after_thread_sync;");
  }
  const int row = tid;
  for (int c0 = 0; c0 < N; c0 += 32) {
    unsigned r[32];
    const unsigned taddr = tmem_d + ((unsigned)(warp * 32) << 16) + (unsigned)c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) D[(size_t)row * N + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(N));
  }
}

// tensor-map builders: tmap.cu (product library)
int make_tmap_f32(CUtensorMap* map, const float* base, long long rows, long long cols, long long ld,
                  int box_rows, int kc);

}  // namespace lslb

// D [128 x N] = A [128 x K] . B [N x K]^T ; N in {64, 128, 256}; K multiple of 32
extern "C" int lslb_debug_tc_probe(const float* A, const float* B, int N, int K, int split, float* D,
                                   lslb_stream_t stream) {
  using namespace lslb;
  if (!A || !B || !D || K < 32 || (K % 32) != 0) return LSLB_ERR_INVALID;
  const int PKH = (split & 6) ? 16 : 32;  // bit 1 of `split`: 16-float chunks, 64-byte swizzle
  const bool bf16corr = (split & 4) != 0;  // bit 2: TF32 main pass + two BF16 correction passes
  split &= 1;
  CUtensorMap mA, mB;
  if (make_tmap_f32(&mA, A, 128, K, K, 128, PKH) != LSLB_OK) return LSLB_ERR_CUDA;
  if (make_tmap_f32(&mB, B, N, K, K, N, PKH) != LSLB_OK) return LSLB_ERR_CUDA;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t smem = (size_t)2 * (128 + N) * PKH * 4 + 1024;
#define LSLB_TC_GO(NN)                                                                                  \
  do {                                                                                                  \
    if (PKH == 32) {                                                                                    \
      auto kern = tc_probe_kernel<NN, 32>;                                                              \
      LSLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
      kern<<<1, 128, smem, s>>>(mA, mB, K, split, D);                                                   \
    } else {                                                                                            \
      auto kern = tc_probe_kernel<NN, 16>;                                                              \
      LSLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
      kern<<<1, 128, smem, s>>>(mA, mB, K, split, D);                                                   \
    }                                                                                                   \
  } while (0)
  if (bf16corr) {
    const size_t smem16 = (size_t)(128 + N) * 16 * 4 + (size_t)2 * (128 + N) * 32 + 1024;
#define LSLB_TC_GO16(NN)                                                                                \
  do {                                                                                                  \
    auto kern = tc_probe_bf16_kernel<NN>;                                                               \
    LSLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16));   \
    kern<<<1, 128, smem16, s>>>(mA, mB, K, D);                                                          \
  } while (0)
    if (N == 64) LSLB_TC_GO16(64);
    else if (N == 128) LSLB_TC_GO16(128);
    else if (N == 256) LSLB_TC_GO16(256);
    else return LSLB_ERR_INVALID;
#undef LSLB_TC_GO16
    LSLB_LAUNCH_CHECK();
    return LSLB_OK;
  }
  if (N == 64) LSLB_TC_GO(64);
  else if (N == 128) LSLB_TC_GO(128);
  else if (N == 256) LSLB_TC_GO(256);
  else return LSLB_ERR_INVALID;
#undef LSLB_TC_GO
  LSLB_LAUNCH_CHECK();
  return LSLB_OK;
}
