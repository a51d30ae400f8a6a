// Several P-splines of DIFFERENT covariates, few chains (S2: 5 cubic smooths x 20 coefficients,
// n = 1e5, 64 chains; float32).
//
// Rows can be span-sorted for one covariate only, so the run-length kernels (banded_x2.cu) do not
// apply, and the generic kernel (logp_grad.cu: thread = chain, gradient columns in shared memory)
// pays four shared-memory read-modify-writes per row and smooth plus a 100-row prologue/epilogue per
// 84-row CTA: 0.15 ms per call against a pipe bound of 8 us.
//
// Here a WARP owns one smooth (and one slice of the rows) and keeps that smooth's k <= 20
// coefficients AND its k gradient accumulators in registers for TWO chains per thread (packed
// f32x2): a register file cannot be indexed dynamically, but the knot span of a row is
// warp-uniform, so `switch (span)` jumps to code with static register indices. Per block of 64
// rows (staged with cp.async, two stages):
//   A  warp (m, j):  f_m[r] = sum_u w_u b[span + u]        -> shared F[r][m][64 chains]
//   C  all warps  :  eta = offsets + sum_m f_m[r]; response log-density and residuals
//                    (row r handled by warp r mod W)        -> shared R[r][predictor][64 chains]
//   E  warp (m, j):  g[span + u] += w_u R[r]                (registers)
// with two barriers per block. One CTA per SM and chain tile; at the end the row slices of a smooth
// are summed in slice order through shared memory, so results are bitwise reproducible. Shared-memory
// traffic per row, smooth and 64 chains: one 512-byte store + two 512-byte loads instead of twelve.
//
// Measured (round 2, profiles/r02_s2_msmooth_full.md): 90 us for S2 (the generic kernel: 110), the call
// 0.124 ms. 384 warp instructions per row of which 40 are packed FMAs: every row and smooth pays a
// constant-table load + indirect branch + reconvergence, and 15 warps per SM cannot hide the dependent
// latencies (issue slots 42 % busy, FMA pipe 20 %; stalls: wait 3.0, barrier 1.4, branch resolving 1.1
// per issued instruction). Grouping four rows per dispatch to batch the shared-memory loads was slower
// (101 us: spills at the 128-register cap); so was the chunk-sorted layout of msort_x2.cu (122 us).
#include "common.cuh"
#include "x2_common.cuh"

namespace lslb {

constexpr int MS_RB = 64;     // rows per staged block
constexpr int MS_KMAX = 20;   // coefficients per smooth held in registers
constexpr int MS_TC = 64;     // chains per CTA (32 lanes x 2)

template <int NS>
struct MsTile {
  float w[NS][MS_RB * 4];
  float y[MS_RB];
  int start[NS][MS_RB];
};

#define MS_FWD_CASE(S)                                   \
  case S:                                                \
    f = mul2(bc(wv.x), b[S]);                            \
    f = fma2(bc(wv.y), b[S + 1], f);                     \
    f = fma2(bc(wv.z), b[S + 2], f);                     \
    f = fma2(bc(wv.w), b[S + 3], f);                     \
    break;
#define MS_BWD_CASE(S)                                   \
  case S:                                                \
    g[S] = fma2(bc(wv.x), rr, g[S]);                     \
    g[S + 1] = fma2(bc(wv.y), rr, g[S + 1]);             \
    g[S + 2] = fma2(bc(wv.z), rr, g[S + 2]);             \
    g[S + 3] = fma2(bc(wv.w), rr, g[S + 3]);             \
    break;
#define MS_ALL_CASES(X)                                                                     \
  X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16)

// NS: number of smooths (compile-time); RS: row slices per smooth; W = NS * RS warps.
template <int FAM, int NS, int RS, bool GRAD>
__global__ void __launch_bounds__(NS * RS * 32, 1)
msmooth_x2_kernel(const __grid_constant__ EvalParams<float> p, int fstride) {
  constexpr int W = NS * RS;
  extern __shared__ __align__(16) unsigned char smraw[];
  MsTile<NS>* tiles = reinterpret_cast<MsTile<NS>*>(smraw);                        // [2]
  u64* F = reinterpret_cast<u64*>(smraw + 2 * sizeof(MsTile<NS>));                 // [MS_RB][NS][32]
  u64* R = F + (size_t)MS_RB * NS * 32;                                            // [MS_RB][2][32]

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int m = wid / RS, j = wid % RS;  // this warp's smooth and row slice
  const int c = blockIdx.y * MS_TC + 2 * lane;  // first of this thread's two chains (< Cpad)
  const size_t ldc = (size_t)p.Cpad;
  auto ld2 = [&](int row) -> u64 { return *reinterpret_cast<const u64*>(p.pT + (size_t)row * ldc + c); };

  // rows of this CTA
  const int f0 = blockIdx.x * fstride;
  const int f1 = min(f0 + fstride, p.nchunks);
  const long long r0 = p.chunk_start[f0], r1 = p.chunk_start[f1];

  // ---- registers: this warp's smooth ------------------------------------------------------------
  u64 b[MS_KMAX], g[MS_KMAX];
  const int so = p.sp[m].soff, km = p.sp[m].k;
#pragma unroll
  for (int t = 0; t < MS_KMAX; ++t) {
    b[t] = t < km ? ld2(so + t) : pk(0.f, 0.f);
    g[t] = pk(0.f, 0.f);
  }
  u64 o0 = pk(0.f, 0.f), o1 = pk(0.f, 0.f);
  for (int q = 0; q < p.ni; ++q) {
    const u64 v = ld2(p.ic[q].soff);
    if (p.ic[q].pred) o1 = add2(o1, v); else o0 = add2(o0, v);
  }
  unsigned pmask = 0;  // bit s set <=> smooth s feeds predictor 1
#pragma unroll
  for (int s = 0; s < NS; ++s) pmask |= (unsigned)(p.sp[s].pred & 1) << s;
  const int mypred = (pmask >> m) & 1;

  float lpa_hi = 0.f, lpa_lo = 0.f, lpb_hi = 0.f, lpb_lo = 0.f;
  u64 sr0 = pk(0.f, 0.f), sr1 = pk(0.f, 0.f);

  // ---- cp.async staging of 64-row blocks --------------------------------------------------------
  auto cp4 = [](void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(x_smem_u32(dst)), "l"(src) : "memory");
  };
  auto cp16 = [](void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(x_smem_u32(dst)), "l"(src) : "memory");
  };
  auto prefetch = [&](long long t0, int stage) {
    if (t0 < r1) {
      const int rows = (int)(r1 - t0 < MS_RB ? r1 - t0 : MS_RB);
      MsTile<NS>& tl = tiles[stage];
      for (int e = tid; e < rows; e += W * 32) cp4(&tl.y[e], p.y + t0 + e);
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        for (int e = tid; e < rows; e += W * 32) cp4(&tl.start[s][e], p.sp[s].start + t0 + e);
        for (int e = tid; e < rows; e += W * 32) cp16(&tl.w[s][4 * e], p.sp[s].w + 4 * (t0 + e));  // rows are 16-byte aligned
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  prefetch(r0, 0);

  int it = 0;
  for (long long t0 = r0; t0 < r1; t0 += MS_RB, ++it) {
    const int rows = (int)(r1 - t0 < MS_RB ? r1 - t0 : MS_RB);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();  // the tile is visible; every warp has left the previous block (its stage, F and R are free)
    prefetch(t0 + MS_RB, (it + 1) & 1);
    const MsTile<NS>& tl = tiles[it & 1];

    // ---- A: forward contribution of this warp's smooth ------------------------------------------
    for (int r = j; r < rows; r += RS) {
      const int s = tl.start[m][r];
      const float4 wv = *reinterpret_cast<const float4*>(&tl.w[m][4 * r]);
      u64 f = pk(0.f, 0.f);
      switch (s) { MS_ALL_CASES(MS_FWD_CASE) default: break; }
      F[((size_t)r * NS + m) * 32 + lane] = f;
    }
    __syncthreads();

    // ---- C: predictors, response, residuals -----------------------------------------------------
    float lta = 0.f, ltb = 0.f;
    for (int r = wid; r < rows; r += W) {
      u64 e0 = o0, e1 = o1;
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const u64 v = F[((size_t)r * NS + s) * 32 + lane];
        if ((pmask >> s) & 1) e1 = add2(e1, v); else e0 = add2(e0, v);
      }
      float e0a, e0b, e1a, e1b;
      upk(e0, e0a, e0b);
      upk(e1, e1a, e1b);
      const float yv = tl.y[r];
      float lla, llb, d0a, d0b, d1a, d1b, u0, u1;
      lik_row<float, FAM, false>(yv, e0a, e1a, p.fixed_inv_scale, lla, d0a, d1a, u0, u1);
      lik_row<float, FAM, false>(yv, e0b, e1b, p.fixed_inv_scale, llb, d0b, d1b, u0, u1);
      lta += lla;
      ltb += llb;
      if constexpr (GRAD) {
        const u64 d0 = pk(d0a, d0b);
        sr0 = add2(sr0, d0);
        R[((size_t)r * 2 + 0) * 32 + lane] = d0;
        if constexpr (FAM == FAM_NORMAL_LS) {
          const u64 d1 = pk(d1a, d1b);
          sr1 = add2(sr1, d1);
          R[((size_t)r * 2 + 1) * 32 + lane] = d1;
        }
      }
    }
    kahan_add(lpa_hi, lpa_lo, lta);
    kahan_add(lpb_hi, lpb_lo, ltb);

    // ---- E: gradient of this warp's smooth ------------------------------------------------------
    if constexpr (GRAD) {
      __syncthreads();
      for (int r = j; r < rows; r += RS) {
        const int s = tl.start[m][r];
        const float4 wv = *reinterpret_cast<const float4*>(&tl.w[m][4 * r]);
        const u64 rr = R[((size_t)r * 2 + mypred) * 32 + lane];
        switch (s) { MS_ALL_CASES(MS_BWD_CASE) default: break; }
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();  // F and R are free: reuse F as the reduction buffer

  // ---- combine the warps in a fixed order and publish this CTA's partial sums ---------------------
  float* mine = p.partial + (size_t)blockIdx.x * (p.Psmall + LSLB_NACC) * ldc + c;
  auto at2 = [&](int row) -> u64* { return reinterpret_cast<u64*>(mine + (size_t)row * ldc); };
  u64* red = F;  // [W][MS_KMAX + 4][32]
  constexpr int RW = MS_KMAX + 4;
  if constexpr (GRAD) {
#pragma unroll
    for (int t = 0; t < MS_KMAX; ++t) red[((size_t)wid * RW + t) * 32 + lane] = g[t];
    red[((size_t)wid * RW + MS_KMAX) * 32 + lane] = sr0;
    red[((size_t)wid * RW + MS_KMAX + 1) * 32 + lane] = sr1;
  }
  red[((size_t)wid * RW + MS_KMAX + 2) * 32 + lane] = pk(lpa_hi, lpb_hi);
  red[((size_t)wid * RW + MS_KMAX + 3) * 32 + lane] = pk(lpa_lo, lpb_lo);
  __syncthreads();
  if constexpr (GRAD) {
    if (j == 0) {  // slice 0 of every smooth sums its RS slices, slice order
      for (int t = 0; t < km; ++t) {
        u64 acc = red[((size_t)(m * RS) * RW + t) * 32 + lane];
        for (int q = 1; q < RS; ++q) acc = add2(acc, red[((size_t)(m * RS + q) * RW + t) * 32 + lane]);
        *at2(so + t) = acc;
      }
    }
  }
  if (wid == 0) {
    if constexpr (GRAD) {
      u64 a0 = pk(0.f, 0.f), a1 = pk(0.f, 0.f);
      for (int q = 0; q < W; ++q) {
        a0 = add2(a0, red[((size_t)q * RW + MS_KMAX) * 32 + lane]);
        a1 = add2(a1, red[((size_t)q * RW + MS_KMAX + 1) * 32 + lane]);
      }
      for (int q = 0; q < p.ni; ++q) *at2(p.ic[q].soff) = p.ic[q].pred ? a1 : a0;
    }
    float ha = 0.f, la = 0.f, hb = 0.f, lb = 0.f;
    for (int q = 0; q < W; ++q) {
      float xa, xb, ya, yb;
      upk(red[((size_t)q * RW + MS_KMAX + 2) * 32 + lane], xa, xb);
      upk(red[((size_t)q * RW + MS_KMAX + 3) * 32 + lane], ya, yb);
      kahan_add(ha, la, xa);
      kahan_add(ha, la, -ya);
      kahan_add(hb, lb, xb);
      kahan_add(hb, lb, -yb);
    }
    *at2(p.Psmall) = pk(ha, hb);
    *at2(p.Psmall + 1) = pk(la, lb);
  }
}

// host ------------------------------------------------------------------------------------------
static int ms_row_slices(int ns) { return ns <= 2 ? 8 : (ns == 3 ? 5 : (ns == 4 ? 4 : (ns == 5 ? 3 : 2))); }

static size_t ms_smem_bytes(int ns) {
  const int W = ns * ms_row_slices(ns);
  size_t tile = (size_t)ns * (MS_RB * 16 + MS_RB * 4) + MS_RB * 4;
  size_t fbuf = (size_t)MS_RB * ns * 32 * 8;
  size_t red = (size_t)W * (MS_KMAX + 4) * 32 * 8;
  if (red > fbuf) fbuf = red;
  return 2 * tile + fbuf + (size_t)MS_RB * 2 * 32 * 8;
}

bool msmooth_x2_eligible(const lslb_plan* plan) {
  static const bool off = [] {
    const char* e = getenv("LSLB_NO_MSMOOTH");
    return e && atoi(e) != 0;
  }();
  if (off) return false;
  if (plan->desc.dtype != LSLB_F32 || plan->variant != VAR_BANDED) return false;
  if (plan->ns < 2 || plan->ns > 8) return false;
  if (banded_fast_eligible(plan)) return false;  // rows span-sortable: the run-length kernels win
  for (int s = 0; s < plan->ns; ++s) {
    const lslb_block_desc& d = plan->blk[plan->sp[s]];
    if (d.ncoef > MS_KMAX || d.ncoef < 4) return false;
    if ((reinterpret_cast<uintptr_t>(d.data) & 15) != 0) return false;
  }
  return plan->desc.n >= 2048;
}

LaunchCfg msmooth_x2_cfg(const lslb_plan* plan, int C) {
  LaunchCfg k;
  k.TC = MS_TC;
  k.Cpad = (C + MS_TC - 1) / MS_TC * MS_TC;
  k.ntiles = k.Cpad / MS_TC;
  long long want = plan->sm_count / k.ntiles;  // one CTA per SM over all chain tiles
  if (want < 1) want = 1;
  const int F = plan->nchunks;
  k.fstride = (int)((F + want - 1) / want);
  if (k.fstride < 1) k.fstride = 1;
  k.nchunks = (F + k.fstride - 1) / k.fstride;
  return k;
}

template <int FAM, int NS, int RS>
static int ms_launch(const EvalParams<float>& ep, const LaunchCfg& k, bool g, cudaStream_t st) {
  const size_t smem = ms_smem_bytes(NS);
  dim3 grid(k.nchunks, k.ntiles), block(NS * RS * 32);
  if (g) {
    auto kern = msmooth_x2_kernel<FAM, NS, RS, true>;
    LSLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, block, smem, st>>>(ep, k.fstride);
  } else {
    auto kern = msmooth_x2_kernel<FAM, NS, RS, false>;
    LSLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, block, smem, st>>>(ep, k.fstride);
  }
  LSLB_LAUNCH_CHECK();
  return LSLB_OK;
}

template <int FAM>
static int ms_dispatch(const lslb_plan* plan, const EvalParams<float>& ep, const LaunchCfg& k, bool g,
                       cudaStream_t st) {
  switch (plan->ns) {
    case 2: return ms_launch<FAM, 2, 8>(ep, k, g, st);
    case 3: return ms_launch<FAM, 3, 5>(ep, k, g, st);
    case 4: return ms_launch<FAM, 4, 4>(ep, k, g, st);
    case 5: return ms_launch<FAM, 5, 3>(ep, k, g, st);
    case 6: return ms_launch<FAM, 6, 2>(ep, k, g, st);
    case 7: return ms_launch<FAM, 7, 2>(ep, k, g, st);
    case 8: return ms_launch<FAM, 8, 2>(ep, k, g, st);
  }
  return LSLB_ERR_UNSUPPORTED;
}

int launch_msmooth_x2(const lslb_plan* plan, const EvalParams<float>& ep, const LaunchCfg& k, bool g,
                      cudaStream_t st) {
  switch (plan->fam) {
    case FAM_NORMAL_LS: return ms_dispatch<FAM_NORMAL_LS>(plan, ep, k, g, st);
    case FAM_NORMAL_FIXED: return ms_dispatch<FAM_NORMAL_FIXED>(plan, ep, k, g, st);
    case FAM_BERNOULLI: return ms_dispatch<FAM_BERNOULLI>(plan, ep, k, g, st);
    case FAM_POISSON: return ms_dispatch<FAM_POISSON>(plan, ep, k, g, st);
  }
  return LSLB_ERR_INVALID;
}

}  // namespace lslb
