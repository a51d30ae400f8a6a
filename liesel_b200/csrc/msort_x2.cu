// Several P-splines of DIFFERENT covariates, few chains (S2: 5 cubic smooths x 20 coefficients,
// n = 1e5, 64 chains; float32) - the chunk-sorted kernel (round 2). EXPERIMENT, OFF BY DEFAULT
// (LSLB_MSORT=1 switches it on at plan creation): parity-green on the GPU (all S2 / family / centring
// tests, BASELINE-size S2 against the fp64 oracle) but SLOWER than msmooth_x2.cu at S2 - 122 us
// against 90 us (profiles/r02_s2_msort_full.md): 12.5 instructions per row, smooth and direction as
// designed, but the walk is a serial chain global load (local row) -> shared load (eta) -> FMA ->
// shared store with no overlap between iterations, runs of ~5 rows per warp segment, and 17
// barriers per 169-row sub-chunk; 16 warps issue 33 % of the time (long-scoreboard 3.5, barrier 2.2,
// wait 2.3 stall cycles per issued instruction). Kept in-tree because it removes the indirect branch
// and is the base for a version with the sorted rows staged in shared memory; see DESIGN.md section 4.
//
// What was wrong with msmooth_x2 (ncu, profiles/r02_s2_msmooth_full.md): a warp kept one smooth's 20
// coefficients and 20 gradient accumulators in registers and reached them through `switch (span)` -
// one constant-table load + indirect branch + reconvergence PER ROW AND SMOOTH, 384 warp instructions
// per row of which 40 are the useful packed FMAs; 15 warps per SM could not hide the dependent
// latencies (issue slots 42 % busy, FMA pipe 20 %): 90 us against a pipe bound of 8 us.
//
// Here the knot span is made loop-invariant instead. The plan cuts the rows into sub-chunks of at
// most MSR_SC rows and stores, for every smooth, each sub-chunk's rows SORTED BY THAT SMOOTH'S SPAN:
// the weights in sorted order, the local row of every sorted position and the run offsets per span
// (msort_prepare). A CTA (16 warps; lane = two chains, packed f32x2) owns a few sub-chunks and keeps
// the sub-chunk's predictors in shared memory, eta[pred][row][64 chains]. Per sub-chunk:
//   forward, smooth by smooth : the 16 warps split the sorted positions evenly; inside a run of equal
//                               span the four coefficients sit in FIXED registers, a row costs two
//                               broadcast loads (local row, weights), 4 FMA2 and eta[row] += f;
//   likelihood                : rows in natural order, eta -> residuals in place;
//   backward, smooth by smooth: same walk, g[0..3] += w_u * residual[row] in fixed registers, added to
//                               the warp's PRIVATE coefficient slice in shared memory when the run
//                               ends; after a barrier the 16 slices are summed in warp order into the
//                               CTA's accumulators (deterministic: no atomics, fixed orders).
// No dynamic register indexing, no indirect branches; 3 ns + 2 barriers per sub-chunk.
#include <stdlib.h>

#include <vector>

#include "common.cuh"
#include "x2_common.cuh"

namespace lslb {

constexpr int MSR_WARPS = 16;
constexpr int MSR_KMAX = 24;    // coefficients per smooth (a warp's private slice holds one smooth)
constexpr int MSR_KTOT = 128;   // spline coefficients of all smooths together (CTA accumulators)
constexpr int MSR_SC = 176;     // rows per sub-chunk (two predictor tiles of 176 x 64 chains = 88 KB)
constexpr int MSR_TC = 64;      // chains per CTA
constexpr int MSR_OFFS = 32;    // ints per (sub-chunk, smooth) in the run-offset table

struct MsortP {
  const float* w[LSLB_MAX_SPL];            // [n][4] weights, sub-chunk-sorted order
  const unsigned short* idx[LSLB_MAX_SPL]; // [n] local row of each sorted position
  const int* off;                          // [nsub][ns][MSR_OFFS] positions at which each span's run starts
  const long long* start;                  // [nsub + 1] first row of each sub-chunk
  int nsub;
};

template <int FAM, bool GRAD>
__global__ void __launch_bounds__(MSR_WARPS * 32, 1)
msort_x2_kernel(const __grid_constant__ EvalParams<float> p, const __grid_constant__ MsortP q, int sub_per_cta) {
  extern __shared__ __align__(16) unsigned char smraw[];
  u64* eta0 = reinterpret_cast<u64*>(smraw);                 // [MSR_SC][32]
  u64* eta1 = eta0 + MSR_SC * 32;                            // [MSR_SC][32]
  u64* GW = eta1 + MSR_SC * 32;                              // [MSR_WARPS][MSR_KMAX][32]
  u64* GC = GW + MSR_WARPS * MSR_KMAX * 32;                  // [MSR_KTOT][32]
  int* offs = reinterpret_cast<int*>(GC + MSR_KTOT * 32);    // [LSLB_MAX_SPL][MSR_OFFS]
  int* coff = offs + LSLB_MAX_SPL * MSR_OFFS;                // [LSLB_MAX_SPL] offset of each smooth in GC

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int c = blockIdx.y * MSR_TC + 2 * lane;  // first of this thread's two chains (< Cpad)
  const size_t ldc = (size_t)p.Cpad;
  auto ld2 = [&](int row) -> u64 { return *reinterpret_cast<const u64*>(p.pT + (size_t)row * ldc + c); };
  const u64 ZERO = pk(0.f, 0.f);

  u64 o0 = ZERO, o1 = ZERO;
  for (int k = 0; k < p.ni; ++k) {
    const u64 v = ld2(p.ic[k].soff);
    if (p.ic[k].pred) o1 = add2(o1, v); else o0 = add2(o0, v);
  }
  if (tid == 0) {
    int a = 0;
    for (int m = 0; m < LSLB_MAX_SPL; ++m) {
      coff[m] = a;
      if (m < p.ns) a += p.sp[m].k;
    }
  }
  if constexpr (GRAD) {
    for (int e = tid; e < MSR_KTOT * 32; e += MSR_WARPS * 32) GC[e] = ZERO;
  }
  float lpa_hi = 0.f, lpa_lo = 0.f, lpb_hi = 0.f, lpb_lo = 0.f;
  u64 sr0 = ZERO, sr1 = ZERO;

  const int sub0 = blockIdx.x * sub_per_cta;
  for (int sb = sub0; sb < sub0 + sub_per_cta && sb < q.nsub; ++sb) {
    const long long r0 = q.start[sb];
    const int rows = (int)(q.start[sb + 1] - r0);
    __syncthreads();  // the previous sub-chunk is finished: eta, offs and GW are free
    for (int e = tid; e < p.ns * MSR_OFFS; e += MSR_WARPS * 32) offs[e] = q.off[(size_t)sb * p.ns * MSR_OFFS + e];
    for (int e = tid; e < rows * 32; e += MSR_WARPS * 32) {
      eta0[e] = ZERO;
      eta1[e] = ZERO;
    }
    __syncthreads();
    const int i0 = (rows * wid) / MSR_WARPS, i1 = (rows * (wid + 1)) / MSR_WARPS;  // this warp's sorted positions

    // ---- forward: eta[pred(m)][row] += sum_u w_u b[span + u] ---------------------------------------
    for (int m = 0; m < p.ns; ++m) {
      u64* eta = p.sp[m].pred ? eta1 : eta0;
      const float4* W = reinterpret_cast<const float4*>(q.w[m]) + r0;
      const unsigned short* IX = q.idx[m] + r0;
      const int* of = offs + m * MSR_OFFS;
      const int so = p.sp[m].soff;
      if (i0 < i1) {
        int s = 0;
        while (of[s + 1] <= i0) ++s;
        int i = i0;
        while (i < i1) {
          const int e = min(i1, of[s + 1]);
          if (e > i) {
            const u64 b0 = ld2(so + s), b1 = ld2(so + s + 1), b2 = ld2(so + s + 2), b3 = ld2(so + s + 3);
            for (; i + 4 <= e; i += 4) {
              int ix[4];
              float4 wv[4];
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                ix[t] = IX[i + t];
                wv[t] = __ldg(&W[i + t]);
              }
              u64 ev[4];
#pragma unroll
              for (int t = 0; t < 4; ++t) ev[t] = eta[ix[t] * 32 + lane];
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                u64 f = fma2(bc(wv[t].x), b0, ev[t]);
                f = fma2(bc(wv[t].y), b1, f);
                f = fma2(bc(wv[t].z), b2, f);
                f = fma2(bc(wv[t].w), b3, f);
                eta[ix[t] * 32 + lane] = f;
              }
            }
            for (; i < e; ++i) {
              const int ix = IX[i];
              const float4 wv = __ldg(&W[i]);
              u64 f = fma2(bc(wv.x), b0, eta[ix * 32 + lane]);
              f = fma2(bc(wv.y), b1, f);
              f = fma2(bc(wv.z), b2, f);
              f = fma2(bc(wv.w), b3, f);
              eta[ix * 32 + lane] = f;
            }
          }
          ++s;
        }
      }
      __syncthreads();  // the next smooth's warps touch the same rows of eta
    }

    // ---- likelihood: rows in natural order, predictors -> residuals in place -----------------------
    float lta = 0.f, ltb = 0.f;
    for (int r = wid; r < rows; r += MSR_WARPS) {
      const u64 e0 = add2(eta0[r * 32 + lane], o0), e1 = add2(eta1[r * 32 + lane], o1);
      float e0a, e0b, e1a, e1b;
      upk(e0, e0a, e0b);
      upk(e1, e1a, e1b);
      const float yv = p.y[r0 + r];
      float lla, llb, d0a, d0b, d1a, d1b, u0, u1;
      lik_row<float, FAM, false>(yv, e0a, e1a, p.fixed_inv_scale, lla, d0a, d1a, u0, u1);
      lik_row<float, FAM, false>(yv, e0b, e1b, p.fixed_inv_scale, llb, d0b, d1b, u0, u1);
      lta += lla;
      ltb += llb;
      if constexpr (GRAD) {
        const u64 d0 = pk(d0a, d0b);
        sr0 = add2(sr0, d0);
        eta0[r * 32 + lane] = d0;
        if constexpr (FAM == FAM_NORMAL_LS) {
          const u64 d1 = pk(d1a, d1b);
          sr1 = add2(sr1, d1);
          eta1[r * 32 + lane] = d1;
        }
      }
    }
    kahan_add(lpa_hi, lpa_lo, lta);
    kahan_add(lpb_hi, lpb_lo, ltb);

    // ---- backward: g[span + u] += w_u * residual[row] -----------------------------------------------
    if constexpr (GRAD) {
      for (int m = 0; m < p.ns; ++m) {
        const int km = p.sp[m].k;
        u64* mine = GW + (size_t)wid * MSR_KMAX * 32;
        __syncthreads();  // residuals are complete (m == 0) / the fold of smooth m - 1 has read every slice
        for (int j = 0; j < km; ++j) mine[j * 32 + lane] = ZERO;  // own slice, used by this warp only until the next barrier
        const u64* res = p.sp[m].pred ? eta1 : eta0;
        const float4* W = reinterpret_cast<const float4*>(q.w[m]) + r0;
        const unsigned short* IX = q.idx[m] + r0;
        const int* of = offs + m * MSR_OFFS;
        if (i0 < i1) {
          int s = 0;
          while (of[s + 1] <= i0) ++s;
          int i = i0;
          while (i < i1) {
            const int e = min(i1, of[s + 1]);
            if (e > i) {
              u64 g0 = ZERO, g1 = ZERO, g2 = ZERO, g3 = ZERO;
              for (; i + 4 <= e; i += 4) {
                float4 wv[4];
                u64 rv[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  wv[t] = __ldg(&W[i + t]);
                  rv[t] = res[(int)IX[i + t] * 32 + lane];
                }
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  g0 = fma2(bc(wv[t].x), rv[t], g0);
                  g1 = fma2(bc(wv[t].y), rv[t], g1);
                  g2 = fma2(bc(wv[t].z), rv[t], g2);
                  g3 = fma2(bc(wv[t].w), rv[t], g3);
                }
              }
              for (; i < e; ++i) {
                const float4 wv = __ldg(&W[i]);
                const u64 rv = res[(int)IX[i] * 32 + lane];
                g0 = fma2(bc(wv.x), rv, g0);
                g1 = fma2(bc(wv.y), rv, g1);
                g2 = fma2(bc(wv.z), rv, g2);
                g3 = fma2(bc(wv.w), rv, g3);
              }
              mine[(s + 0) * 32 + lane] = add2(mine[(s + 0) * 32 + lane], g0);
              mine[(s + 1) * 32 + lane] = add2(mine[(s + 1) * 32 + lane], g1);
              mine[(s + 2) * 32 + lane] = add2(mine[(s + 2) * 32 + lane], g2);
              mine[(s + 3) * 32 + lane] = add2(mine[(s + 3) * 32 + lane], g3);
            }
            ++s;
          }
        }
        __syncthreads();  // every warp's slice of smooth m is complete
        for (int j = wid; j < km; j += MSR_WARPS) {  // fold the 16 slices in warp order
          u64 acc = GC[(coff[m] + j) * 32 + lane];
#pragma unroll
          for (int w = 0; w < MSR_WARPS; ++w) acc = add2(acc, GW[((size_t)w * MSR_KMAX + j) * 32 + lane]);
          GC[(coff[m] + j) * 32 + lane] = acc;
        }
      }
    }
  }
  __syncthreads();  // GC is complete; GW is free: reuse it as the reduction buffer

  // ---- combine the warps in a fixed order and publish this CTA's partial sums -----------------------
  float* out = p.partial + (size_t)blockIdx.x * (p.Psmall + LSLB_NACC) * ldc + c;
  auto at2 = [&](int row) -> u64* { return reinterpret_cast<u64*>(out + (size_t)row * ldc); };
  u64* red = GW;  // [MSR_WARPS][4][32]
  red[(wid * 4 + 0) * 32 + lane] = sr0;
  red[(wid * 4 + 1) * 32 + lane] = sr1;
  red[(wid * 4 + 2) * 32 + lane] = pk(lpa_hi, lpb_hi);
  red[(wid * 4 + 3) * 32 + lane] = pk(lpa_lo, lpb_lo);
  __syncthreads();
  if constexpr (GRAD) {
    for (int m = 0; m < p.ns; ++m)
      for (int j = wid; j < p.sp[m].k; j += MSR_WARPS) *at2(p.sp[m].soff + j) = GC[(coff[m] + j) * 32 + lane];
  }
  if (wid == 0) {
    if constexpr (GRAD) {
      u64 a0 = ZERO, a1 = ZERO;
      for (int w = 0; w < MSR_WARPS; ++w) {
        a0 = add2(a0, red[(w * 4 + 0) * 32 + lane]);
        a1 = add2(a1, red[(w * 4 + 1) * 32 + lane]);
      }
      for (int k = 0; k < p.ni; ++k) *at2(p.ic[k].soff) = p.ic[k].pred ? a1 : a0;
    }
    float ha = 0.f, la = 0.f, hb = 0.f, lb = 0.f;
    for (int w = 0; w < MSR_WARPS; ++w) {
      float xa, xb, ya, yb;
      upk(red[(w * 4 + 2) * 32 + lane], xa, xb);
      upk(red[(w * 4 + 3) * 32 + lane], ya, yb);
      kahan_add(ha, la, xa);
      kahan_add(ha, la, -ya);
      kahan_add(hb, lb, xb);
      kahan_add(hb, lb, -yb);
    }
    *at2(p.Psmall) = pk(ha, hb);
    *at2(p.Psmall + 1) = pk(la, lb);
  }
}

// ---- plan side: sub-chunks and their per-smooth sorted orders ---------------------------------------
// One CTA per sub-chunk: stable counting sort of the sub-chunk's rows by span (quadratic ranking, the
// sub-chunk has at most MSR_SC rows; plan creation only).
__global__ void msort_build_kernel(const int* __restrict__ start, const float* __restrict__ w,
                                   const long long* __restrict__ sub_start, int nspan, int m, int ns,
                                   unsigned short* __restrict__ idx_out, float* __restrict__ w_out,
                                   int* __restrict__ off_out) {
  __shared__ int sp[MSR_SC];
  const int sb = blockIdx.x, t = threadIdx.x;
  const long long r0 = sub_start[sb];
  const int rows = (int)(sub_start[sb + 1] - r0);
  if (t < rows) sp[t] = start[r0 + t];
  __syncthreads();
  if (t < rows) {
    const int s = sp[t];
    int rank = 0;
    for (int r = 0; r < rows; ++r) rank += (sp[r] < s) || (sp[r] == s && r < t);
    idx_out[r0 + rank] = (unsigned short)t;
    const float4 v = *reinterpret_cast<const float4*>(w + 4 * (r0 + t));
    *reinterpret_cast<float4*>(w_out + 4 * (r0 + rank)) = v;
  }
  if (t < MSR_OFFS) {
    int below = 0;
    if (t <= nspan) {
      for (int r = 0; r < rows; ++r) below += sp[r] < t;
    } else {
      below = rows;
    }
    off_out[((size_t)sb * ns + m) * MSR_OFFS + t] = below;
  }
}

static bool msort_structurally_ok(const lslb_plan* plan) {
  if (plan->desc.dtype != LSLB_F32 || plan->variant != VAR_BANDED) return false;
  if (plan->ns < 2 || plan->ns > LSLB_MAX_SPL) return false;
  int ktot = 0;
  for (int s = 0; s < plan->ns; ++s) {
    const lslb_block_desc& d = plan->blk[plan->sp[s]];
    if (d.ncoef > MSR_KMAX || d.ncoef < 4) return false;
    if ((reinterpret_cast<uintptr_t>(d.data) & 15) != 0) return false;
    ktot += d.ncoef;
  }
  if (ktot > MSR_KTOT) return false;
  return plan->desc.n >= 2048;
}

// called by lslb_plan_create once the variant is known
int msort_prepare(lslb_plan* plan) {
  plan->ms_nsub = 0;
  const char* e = getenv("LSLB_MSORT");
  if (!(e && atoi(e) != 0)) return LSLB_OK;  // opt-in experiment (see the header)
  if (!msort_structurally_ok(plan) || banded_fast_eligible(plan)) return LSLB_OK;
  const long long n = plan->desc.n;
  const int want = (int)((n + MSR_SC - 1) / MSR_SC);
  const int ctas = want < plan->sm_count ? want : plan->sm_count;
  const int nsub = (want + ctas - 1) / ctas * ctas;  // a multiple of the row-chunk grid
  std::vector<long long> st(nsub + 1);
  for (int s = 0; s <= nsub; ++s) st[s] = (long long)(((__int128)n * s) / nsub);
  cudaError_t err = cudaMalloc(&plan->d_ms_start, sizeof(long long) * (nsub + 1));
  if (err == cudaSuccess)
    err = cudaMemcpy(plan->d_ms_start, st.data(), sizeof(long long) * (nsub + 1), cudaMemcpyHostToDevice);
  if (err == cudaSuccess) err = cudaMalloc(&plan->d_ms_off, sizeof(int) * (size_t)nsub * plan->ns * MSR_OFFS);
  for (int m = 0; m < plan->ns && err == cudaSuccess; ++m) {
    const lslb_block_desc& d = plan->blk[plan->sp[m]];
    err = cudaMalloc(&plan->d_ms_w[m], sizeof(float) * 4 * (size_t)n);
    if (err == cudaSuccess) err = cudaMalloc(&plan->d_ms_idx[m], sizeof(unsigned short) * (size_t)n);
    if (err == cudaSuccess)
      msort_build_kernel<<<nsub, 256>>>(d.index, (const float*)d.data, plan->d_ms_start, d.ncoef - 3, m, plan->ns,
                                        plan->d_ms_idx[m], plan->d_ms_w[m], plan->d_ms_off);
  }
  if (err == cudaSuccess) err = cudaDeviceSynchronize();
  if (err == cudaSuccess) err = cudaGetLastError();
  if (err != cudaSuccess) {
    set_cuda_error(err, "chunk-sorted smooth orders");
    return LSLB_ERR_CUDA;
  }
  plan->ms_nsub = nsub;
  return LSLB_OK;
}

void msort_release(lslb_plan* plan) {
  if (plan->d_ms_start) cudaFree(plan->d_ms_start);
  if (plan->d_ms_off) cudaFree(plan->d_ms_off);
  for (int m = 0; m < LSLB_MAX_SPL; ++m) {
    if (plan->d_ms_w[m]) cudaFree(plan->d_ms_w[m]);
    if (plan->d_ms_idx[m]) cudaFree(plan->d_ms_idx[m]);
  }
}

bool msort_x2_eligible(const lslb_plan* plan) { return plan->ms_nsub > 0; }

LaunchCfg msort_x2_cfg(const lslb_plan* plan, int C) {
  LaunchCfg k;
  k.TC = MSR_TC;
  k.Cpad = (C + MSR_TC - 1) / MSR_TC * MSR_TC;
  k.ntiles = k.Cpad / MSR_TC;
  k.nchunks = plan->ms_nsub < plan->sm_count ? plan->ms_nsub : plan->sm_count;  // row chunks = grid.x
  k.fstride = plan->ms_nsub / k.nchunks;                                         // sub-chunks per CTA
  return k;
}

static size_t msort_smem_bytes() {
  return (size_t)(2 * MSR_SC * 32 + MSR_WARPS * MSR_KMAX * 32 + MSR_KTOT * 32) * 8 + LSLB_MAX_SPL * MSR_OFFS * 4 +
         LSLB_MAX_SPL * 4;
}

template <int FAM>
static int msort_go(const lslb_plan* plan, const EvalParams<float>& ep, const MsortP& q, const LaunchCfg& k, bool g,
                    cudaStream_t st) {
  const size_t smem = msort_smem_bytes();
  dim3 grid(k.nchunks, k.ntiles), block(MSR_WARPS * 32);
  if (g) {
    auto kern = msort_x2_kernel<FAM, true>;
    LSLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, block, smem, st>>>(ep, q, k.fstride);
  } else {
    auto kern = msort_x2_kernel<FAM, false>;
    LSLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, block, smem, st>>>(ep, q, k.fstride);
  }
  LSLB_LAUNCH_CHECK();
  return LSLB_OK;
}

int launch_msort_x2(const lslb_plan* plan, const EvalParams<float>& ep, const LaunchCfg& k, bool g,
                    cudaStream_t st) {
  MsortP q;
  for (int m = 0; m < LSLB_MAX_SPL; ++m) {
    q.w[m] = plan->d_ms_w[m];
    q.idx[m] = plan->d_ms_idx[m];
  }
  q.off = plan->d_ms_off;
  q.start = plan->d_ms_start;
  q.nsub = plan->ms_nsub;
  switch (plan->fam) {
    case FAM_NORMAL_LS: return msort_go<FAM_NORMAL_LS>(plan, ep, q, k, g, st);
    case FAM_NORMAL_FIXED: return msort_go<FAM_NORMAL_FIXED>(plan, ep, q, k, g, st);
    case FAM_BERNOULLI: return msort_go<FAM_BERNOULLI>(plan, ep, q, k, g, st);
    case FAM_POISSON: return msort_go<FAM_POISSON>(plan, ep, q, k, g, st);
  }
  return LSLB_ERR_INVALID;
}

}  // namespace lslb
