// Prologue kernels shared by the logp/grad and IWLS passes (each TU instantiates its own copy).
#pragma once
#include "common.cuh"

namespace lslb {

// =======================================================================================
// gather / scatter helpers
// =======================================================================================
// params [C x P] -> pT [Psmall x Cpad] (coefficient-major, chains contiguous). Blocks below nb_real: one
// thread per (real row, chain). Blocks from nb_real on: the VIRTUAL rows (minus the centring term
// cbar . beta of every centred block of a predictor), one warp per (virtual row, chain) with the lanes
// over the coefficients - coalesced loads and a fixed shuffle tree instead of a serial chain of
// ncoef x nblocks dependent double FMAs on strided loads (12 us at the S2 size).
template <typename T>
__global__ void gather_small_kernel(const __grid_constant__ FinalizeParams fp,
                                    const T* __restrict__ params, T* __restrict__ pT, int nb_real,
                                    int* __restrict__ zero_ints, int nz) {
  // (also clears the per-chain-group counters of this call's finalize kernel: same stream, earlier launch)
  if (zero_ints != nullptr && blockIdx.x == 0)
    for (int i = threadIdx.x; i < nz; i += blockDim.x) zero_ints[i] = 0;
  if ((int)blockIdx.x >= nb_real) {
    const int wpb = blockDim.x >> 5;
    const int per_v = (fp.Cpad + wpb - 1) / wpb;  // blocks per virtual row
    const int vb = blockIdx.x - nb_real;
    const int j = fp.Preal + vb / per_v;
    const int c = (vb % per_v) * wpb + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (c >= fp.Cpad || j >= fp.Psmall) return;
    double acc = 0.0;
    if (c < fp.C) {
      for (int b = 0; b < fp.nblocks; ++b) {
        const PriorP& pr = fp.pr[b];
        if (pr.vrow != j) continue;
        const T* cb = reinterpret_cast<const T*>(pr.center);
        const T* x = params + (long long)c * fp.P + pr.off;
        for (int t = lane; t < pr.ncoef; t += 32) acc += (double)cb[t] * (double)x[t];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) pT[(long long)j * fp.Cpad + c] = (T)(-acc);
    return;
  }
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)fp.Preal * fp.Cpad;
  if (idx >= total) return;
  int c = (int)(idx % fp.Cpad);
  int j = (int)(idx / fp.Cpad);
  T v = T(0);
  if (c < fp.C) {
    for (int b = 0; b < fp.nblocks; ++b) {
      const PriorP& pr = fp.pr[b];
      if (pr.soff >= 0 && j >= pr.soff && j < pr.soff + pr.ncoef) {
        v = params[(long long)c * fp.P + pr.off + (j - pr.soff)];
        break;
      }
    }
  }
  pT[idx] = v;
}

template <typename T>
inline int launch_gather_small(const FinalizeParams& fp, const T* params, T* pT, cudaStream_t stream,
                               int* zero_ints = nullptr, int nz = 0) {
  const long long total = (long long)fp.Preal * fp.Cpad;
  const int nb_real = (int)((total + 255) / 256);
  const int nv = fp.Psmall - fp.Preal;
  const int nb = nb_real + nv * ((fp.Cpad + 7) / 8);
  if (nb > 0 || zero_ints != nullptr) {
    gather_small_kernel<T><<<(unsigned)(nb > 0 ? nb : 1), 256, 0, stream>>>(fp, params, pT, nb_real, zero_ints, nz);
    LSLB_LAUNCH_CHECK();
  }
  return LSLB_OK;
}

// params[:, goff:goff+G] ([C x G] view, row stride P)  ->  bT [G x Cpad]; also initialises the
// gradient buffer with the prior gradient and emits per-tile prior log-density partials.
template <typename T>
__global__ void group_init_kernel(const T* __restrict__ params, int P, int goff, int G, int C,
                                  int Cpad, T m, T inv_s2, bool with_prior, T* __restrict__ bT,
                                  T* __restrict__ gbT, T* __restrict__ gpl) {
  __shared__ T tile[32][33];
  __shared__ T red[8][32];
  int g0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int cl = ty + 8 * k;
    int g = g0 + tx, c = c0 + cl;
    tile[cl][tx] = (g < G && c < C) ? params[(long long)c * P + goff + g] : T(0);
  }
  __syncthreads();
  T acc = T(0);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int gl = ty + 8 * k;
    int g = g0 + gl, c = c0 + tx;
    if (g < G && c < Cpad) {
      T b = tile[tx][gl];
      bT[(long long)g * Cpad + c] = b;
      T d = b - m;
      gbT[(long long)g * Cpad + c] = with_prior ? -d * inv_s2 : T(0);
      if (c < C) acc += T(-0.5) * d * d * inv_s2;
    }
  }
  red[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && c0 + tx < Cpad) {
    T s = T(0);
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][tx];
    gpl[(long long)blockIdx.x * Cpad + c0 + tx] = s;
  }
}

}  // namespace lslb
