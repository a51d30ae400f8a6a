// Packed-FP32 helpers and the TMA-staged row tile shared by the two-chains-per-thread kernels.
#pragma once
#include "common.cuh"

namespace lslb {

constexpr int XR = 128;     // rows per tile
constexpr int XSTAGES = 3;  // ring depth
typedef unsigned long long u64;

__device__ __forceinline__ unsigned x_smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ u64 pk(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ u64 bc(float v) {  // {v, v}: ptxas folds this into a .F32 broadcast operand
  u64 r;
  asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void upk(u64 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float ex2f(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ float rcpf(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float lg2f(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

template <int NS>
struct XRaw {  // what the TMA engine writes: the caller's SoA arrays, one tile
  float w[NS][XR * 4];
  float y[XR];
  int start[NS][XR];
};


// One CTA's view of the 3-stage TMA ring over 128-row tiles (see banded_x2.cu for the protocol).
template <int NS>
struct XPipe {
  XRaw<NS>* raw;
  u64* full;
  const EvalParams<float>* p;
  __device__ __forceinline__ int tile_rows(int t) const {
    long long rem = p->n - (long long)t * XR;
    return rem >= XR ? XR : (int)rem;
  }
  __device__ __forceinline__ void init(int tid) {
    if (tid == 0) {
#pragma unroll
      for (int s = 0; s < XSTAGES; ++s)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(x_smem_u32(&full[s])));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
  }
  __device__ __forceinline__ void bulk(void* dst, const void* src, unsigned bytes, u64* bar) const {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            x_smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(x_smem_u32(bar))
        : "memory");
  }
  __device__ __forceinline__ void issue(int t, int s) const {
    const long long r0 = (long long)t * XR;
    constexpr unsigned bytes = XR * 4 + NS * (XR * 16 + XR * 4);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(x_smem_u32(&full[s])),
                 "r"(bytes)
                 : "memory");
    bulk(raw[s].y, p->y + r0, XR * 4, &full[s]);
#pragma unroll
    for (int m = 0; m < NS; ++m) {
      bulk(raw[s].w[m], p->sp[m].w + 4 * r0, XR * 16, &full[s]);
      bulk(raw[s].start[m], p->sp[m].start + r0, XR * 4, &full[s]);
    }
  }
  __device__ __forceinline__ void wait_full(int s, unsigned parity) const {
    unsigned ok = 0;
    while (!ok) {
      asm volatile(
          "{\n.reg .pred q;\nmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\nselp.u32 %0, 1, 0, q;\n}\n"
          : "=r"(ok)
          : "r"(x_smem_u32(&full[s])), "r"(parity)
          : "memory");
    }
  }
  // ragged last tile: ordinary loads (bulk copies need 16-byte multiples); caller syncs afterwards
  __device__ __forceinline__ void load_ragged(int t, int s, int rows, int tid, int nthreads) const {
    const long long r0 = (long long)t * XR;
    XRaw<NS>& D = raw[s];
    for (int r = tid; r < rows; r += nthreads) {
      D.y[r] = p->y[r0 + r];
#pragma unroll
      for (int m = 0; m < NS; ++m) {
        D.start[m][r] = p->sp[m].start[r0 + r];
#pragma unroll
        for (int u = 0; u < 4; ++u) D.w[m][4 * r + u] = p->sp[m].w[4 * (r0 + r) + u];
      }
    }
  }
};

}  // namespace lslb
