// TOOLS LIBRARY (libliesel_b200_bench.so, not part of the product library): pipe-peak
// micro-benchmarks (FP32 FMA, packed FFMA2, MUFU ex2, shared-memory bandwidth, register-file probes)
// that provide the MEASURED roofline denominators MEASURED_PEAKS.json does not contain.
#include "common.cuh"
#include "liesel_b200_bench.h"

namespace lslb {
// 8 independent dependency chains per thread; every FMA has a live result
__global__ void __launch_bounds__(256) fma_peak_kernel(int iters, float a, float b, float* sink) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5,
        x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
  }
  float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 123.456f) sink[0] = s;
}

// same work with the packed fma.rn.f32x2 (two lanes' worth of FP32 FMA per instruction)
__global__ void __launch_bounds__(256) ffma2_peak_kernel(int iters, float a, float b, float* sink) {
  unsigned long long x[4], av, bv;
  asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float lo = threadIdx.x + 2 * k, hi = lo + 1;
    asm("mov.b64 %0, {%1, %2};" : "=l"(x[k]) : "f"(lo), "f"(hi));
  }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[k]) : "l"(av), "l"(bv));
    }
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[k]));
    s += lo + hi;
  }
  if (s == 123.456f) sink[0] = s;
}

// packed and scalar FMAs interleaved 1:1 (instruction count): can the two FP32 datapaths of an
// SMSP run an FFMA2 and an FFMA side by side?
__global__ void __launch_bounds__(256) mixed_peak_kernel(int iters, float a, float b, float* sink) {
  unsigned long long x[4], av, bv;
  asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
  float y0 = threadIdx.x, y1 = y0 + 1, y2 = y0 + 2, y3 = y0 + 3;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float lo = threadIdx.x + 2 * k, hi = lo + 1;
    asm("mov.b64 %0, {%1, %2};" : "=l"(x[k]) : "f"(lo), "f"(hi));
  }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[0]) : "l"(av), "l"(bv));
      asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(y0) : "f"(a), "f"(b));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[1]) : "l"(av), "l"(bv));
      asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(y1) : "f"(a), "f"(b));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[2]) : "l"(av), "l"(bv));
      asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(y2) : "f"(a), "f"(b));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[3]) : "l"(av), "l"(bv));
      asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(y3) : "f"(a), "f"(b));
    }
  }
  float s = y0 + y1 + y2 + y3;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[k]));
    s += lo + hi;
  }
  if (s == 123.456f) sink[0] = s;
}

// Register-file bandwidth probes. x_k = fma(a_k, b_k, x_k) over 8 accumulators:
//   SHARED_B = false: three distinct register operands per instruction
//   SHARED_B = true : b is the same register in every instruction (operand-reuse cache can serve it)
template <bool PACKED, bool SHARED_B>
__global__ void __launch_bounds__(256) rf_probe_kernel(int iters, float a0, float b0, float* sink) {
  if constexpr (PACKED) {
    unsigned long long x[8], a[8], b[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float lo = threadIdx.x + 2 * k, hi = lo + 1;
      asm("mov.b64 %0, {%1, %2};" : "=l"(x[k]) : "f"(lo), "f"(hi));
      float av = a0 + 1e-6f * k, bv = b0 + 1e-3f * (SHARED_B ? 0 : k);
      asm("mov.b64 %0, {%1, %1};" : "=l"(a[k]) : "f"(av));
      asm("mov.b64 %0, {%1, %1};" : "=l"(b[k]) : "f"(bv));
    }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(x[k]) : "l"(a[k]), "l"(b[SHARED_B ? 0 : k]));
      }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float lo, hi;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[k]));
      s += lo + hi;
    }
    if (s == 123.456f) sink[0] = s;
  } else {
    float x[8], a[8], b[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      x[k] = threadIdx.x + k;
      a[k] = a0 + 1e-6f * k;
      b[k] = b0 + 1e-3f * (SHARED_B ? 0 : k);
    }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(x[k]) : "f"(a[k]), "f"(b[SHARED_B ? 0 : k]));
      }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += x[k];
    if (s == 123.456f) sink[0] = s;
  }
}

__global__ void __launch_bounds__(256) mufu_peak_kernel(int iters, float a, float* sink) {
  float x0 = threadIdx.x * 1e-3f, x1 = x0 + .1f, x2 = x0 + .2f, x3 = x0 + .3f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = exp2f(x0) - a; x1 = exp2f(x1) - a; x2 = exp2f(x2) - a; x3 = exp2f(x3) - a;
    }
  }
  float s = x0 + x1 + x2 + x3;
  if (s == 123.456f) sink[0] = s;
}

// every warp reads one conflict-free 128-byte row of shared memory per LDS.32
__global__ void __launch_bounds__(256) lds_peak_kernel(int iters, float* sink) {
  __shared__ float sm[256 * 8];
  for (int i = threadIdx.x; i < 256 * 8; i += 256) sm[i] = i;
  __syncthreads();
  float acc = 0.f;
  int idx = threadIdx.x;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += sm[(idx + 256 * u) & 2047];
    idx = (idx + 32) & 255;
  }
  if (acc == 123.456f) sink[0] = acc;
}
}  // namespace lslb

using namespace lslb;

// kind: 0 FFMA, 1 FFMA2 (f32x2), 2 MUFU.EX2, 3 LDS.  *work receives the number of operations one
// launch performs (FMA lanes, ex2 calls, or 4-byte shared loads).
extern "C" int lslb_debug_pipe_peak(int kind, int blocks, int iters, float* sink, double* work,
                                    lslb_stream_t stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const double threads = (double)blocks * 256.0;
  switch (kind) {
    case 0:
      fma_peak_kernel<<<blocks, 256, 0, s>>>(iters, 1.0001f, 0.5f, sink);
      *work = threads * 64.0 * iters;
      break;
    case 1:
      ffma2_peak_kernel<<<blocks, 256, 0, s>>>(iters, 1.0001f, 0.5f, sink);
      *work = threads * 64.0 * iters;
      break;
    case 2:
      mufu_peak_kernel<<<blocks, 256, 0, s>>>(iters, 1.0f, sink);
      *work = threads * 32.0 * iters;
      break;
    case 3:
      lds_peak_kernel<<<blocks, 256, 0, s>>>(iters, sink);
      *work = threads * 8.0 * iters;
      break;
    case 4:  // per unrolled step: 4 FFMA2 (8 FMA lanes-ops) + 4 FFMA = 12 FMAs per thread, x8 unroll
      mixed_peak_kernel<<<blocks, 256, 0, s>>>(iters, 1.0001f, 0.5f, sink);
      *work = threads * 96.0 * iters;
      break;
    case 5:  // FFMA2, three distinct register-pair operands
      rf_probe_kernel<true, false><<<blocks, 256, 0, s>>>(iters, 1.0001f, 0.5f, sink);
      *work = threads * 64.0 * iters;
      break;
    case 6:  // FFMA2, one operand shared by consecutive instructions
      rf_probe_kernel<true, true><<<blocks, 256, 0, s>>>(iters, 1.0001f, 0.5f, sink);
      *work = threads * 64.0 * iters;
      break;
    case 7:  // FFMA, three distinct register operands
      rf_probe_kernel<false, false><<<blocks, 256, 0, s>>>(iters, 1.0001f, 0.5f, sink);
      *work = threads * 32.0 * iters;
      break;
    case 8:  // FFMA, one operand shared
      rf_probe_kernel<false, true><<<blocks, 256, 0, s>>>(iters, 1.0001f, 0.5f, sink);
      *work = threads * 32.0 * iters;
      break;
    default: return LSLB_ERR_INVALID;
  }
  LSLB_LAUNCH_CHECK();
  return LSLB_OK;
}
