// Dense fixed-effect block on the FP32 pipe (S4: logistic regression, p = 256, 1024 chains).
//
//   eta[rows x chains] = X[rows x p] . B[p x chains],   G[p x chains] = X^T . R,   R = dl/d eta
//
// Both contractions are fused around the per-row likelihood; eta and R never touch HBM.
// Register-sliced layout: a warp owns 8 chains (4 packed pairs per thread, FFMA2) and its 32 lanes
// own 32 column slices of X (CS = p/32 columns each), so that
//   * the coefficient slice B[slice, 8 chains] and the gradient slice G[slice, 8 chains] live in
//     REGISTERS for the whole kernel (per chain the kernel needs 2p floats of fast storage; shared
//     memory could hold that for only ~110 chains per SM, the register file holds 8 per warp);
//   * every lane reads a DIFFERENT 4*CS-byte piece of a row from shared memory (no redundant
//     broadcast phases: the row tile is read once per warp and pass);
//   * the partial dot products of 8 rows are combined by a 5-step reduce-scatter over the lanes
//     (shuffles), after which lane L holds eta for (row L/4, pair L%4) — the likelihood is evaluated
//     exactly once per (row, chain) — and the residuals are all-gathered back the same way.
// X tiles (32 rows) are staged with cp.async through a 3-stage ring; y likewise.
//
// Why not tcgen05 here: fp32 parity (rtol 1e-5) needs a 3-pass TF32 or 6-pass BF16 split of both
// operands; with 4-byte TF32 operands the resident B tile (128 chains x 256 x 2 pieces = 256 KB) and
// the X tile (needed twice: X.B and X^T.R) do not fit the 227 KB of shared memory next to the R
// tile, and the unfused alternative writes R (5 GB per GPU and call) through HBM. See DESIGN.md §4.
#include "common.cuh"

namespace lslb {

typedef unsigned long long u64;
constexpr int DRT = 32;      // rows per tile
constexpr int DSTAGES = 3;
constexpr int DWARPS = 8;    // warps per CTA -> 64 chains per CTA

__device__ __forceinline__ u64 dpk(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ u64 dbc(float v) {
  u64 r;
  asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void dupk(u64 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 dfma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 dadd2(u64 a, u64 b) {
  u64 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 shfl_xor64(u64 v, int m) {
  return __shfl_xor_sync(0xffffffffu, v, m);
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)),
               "l"(src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// CS: columns per lane (p_pad = 32*CS); CS in {1,2,4,8}
template <int CS, int FAM, bool GRAD>
__global__ void __launch_bounds__(DWARPS * 32, 1)
dense_x2_kernel(const __grid_constant__ EvalParams<float> p, int tiles_per_cta, int ntiles_total) {
  constexpr int PP = 32 * CS;
  extern __shared__ __align__(128) unsigned char smraw[];
  float* xs = reinterpret_cast<float*>(smraw);                   // [DSTAGES][DRT][PP]
  float* ys = xs + (size_t)DSTAGES * DRT * PP;                   // [DSTAGES][DRT]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cbase = (blockIdx.x * DWARPS + warp) * 8;            // this warp's 8 chains (< Cpad)
  const int ch = blockIdx.y;
  const int tile0 = ch * tiles_per_cta;
  const int my_tiles = min(tile0 + tiles_per_cta, ntiles_total) - tile0;
  const DnsP<float>& D = p.dn[0];
  const int pc = D.p;

  auto tile_rows = [&](int t) -> int {
    long long rem = p.n - (long long)t * DRT;
    return rem >= DRT ? DRT : (int)rem;
  };
  // cooperative async copy of one tile; columns >= p and rows beyond n are zero-filled
  auto issue = [&](int t, int s) {
    const long long r0 = (long long)t * DRT;
    const int rows = tile_rows(t);
    float* xd = xs + (size_t)s * DRT * PP;
    constexpr int CH = PP / 4;  // 16-byte chunks per row
    for (int q = tid; q < DRT * CH; q += DWARPS * 32) {
      const int r = q / CH, c4 = (q % CH) * 4;
      float* dst = xd + r * PP + c4;
      if (r < rows && c4 + 3 < pc) {
        cp_async16(dst, D.X + (r0 + r) * D.ld + c4);
      } else {
        float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < rows) {
          const float* src = D.X + (r0 + r) * D.ld;
          if (c4 + 0 < pc) z.x = src[c4 + 0];
          if (c4 + 1 < pc) z.y = src[c4 + 1];
          if (c4 + 2 < pc) z.z = src[c4 + 2];
        }
        *reinterpret_cast<float4*>(dst) = z;
      }
    }
    if (tid < DRT) ys[s * DRT + tid] = tid < rows ? p.y[r0 + tid] : 0.f;
  };

  for (int k = 0; k < DSTAGES - 1; ++k) {
    if (k < my_tiles) issue(tile0 + k, k);
    cp_async_commit();
  }

  // ---- register-resident slices ----------------------------------------------------------------
  const size_t ldc = (size_t)p.Cpad;
  u64 b[CS][4], g[CS][4];
#pragma unroll
  for (int u = 0; u < CS; ++u) {
    const int j = lane * CS + u;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      b[u][k] = j < pc ? *reinterpret_cast<const u64*>(p.pT + (size_t)(D.soff + j) * ldc + cbase + 2 * k)
                       : dpk(0.f, 0.f);
      g[u][k] = dpk(0.f, 0.f);
    }
  }
  // after the reduce-scatter lane L owns (row slot L>>2, pair L&3)
  const int myk = lane & 3, myr = lane >> 2;
  u64 off = dpk(0.f, 0.f);  // intercept terms of this lane's pair
  for (int q = 0; q < p.ni; ++q)
    off = dadd2(off, *reinterpret_cast<const u64*>(p.pT + (size_t)p.ic[q].soff * ldc + cbase + 2 * myk));
  float lpa = 0.f, lpb = 0.f, lpa_lo = 0.f, lpb_lo = 0.f;
  u64 sd = dpk(0.f, 0.f);

  for (int k = 0; k < my_tiles; ++k) {
    const int t = tile0 + k, s = k % DSTAGES;
    const int rows = tile_rows(t);
    {
      const int kn = k + DSTAGES - 1;
      if (kn < my_tiles) issue(tile0 + kn, kn % DSTAGES);
      cp_async_commit();
    }
    cp_async_wait<DSTAGES - 1>();
    __syncthreads();
    const float* xt = xs + (size_t)s * DRT * PP + lane * CS;
    const float* yt = ys + s * DRT;
    float tla = 0.f, tlb = 0.f;

    for (int r0 = 0; r0 < rows; r0 += 8) {
      // ---- forward: partial dot products of 8 rows x 4 pairs over this lane's columns -------
      u64 v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = dpk(0.f, 0.f);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        float xv[CS];
        const float* xr = xt + (r0 + r) * PP;
        if constexpr (CS == 8) {
          const float4 a = *reinterpret_cast<const float4*>(xr), c2 = *reinterpret_cast<const float4*>(xr + 4);
          xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w; xv[4] = c2.x; xv[5] = c2.y; xv[6] = c2.z; xv[7] = c2.w;
        } else if constexpr (CS == 4) {
          const float4 a = *reinterpret_cast<const float4*>(xr);
          xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w;
        } else if constexpr (CS == 2) {
          const float2 a = *reinterpret_cast<const float2*>(xr);
          xv[0] = a.x; xv[1] = a.y;
        } else {
          xv[0] = xr[0];
        }
#pragma unroll
        for (int u = 0; u < CS; ++u)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) v[r * 4 + kk] = dfma2(dbc(xv[u]), b[u][kk], v[r * 4 + kk]);
      }
      // ---- reduce-scatter over the 32 lanes: entry index i ends up on lane i ----------------
#pragma unroll
      for (int half = 16; half >= 1; half >>= 1) {
        const bool up = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
          const u64 lo = v[i], hi = v[i + half];
          const u64 send = up ? lo : hi;
          const u64 keep = up ? hi : lo;
          v[i] = dadd2(keep, shfl_xor64(send, half));
        }
      }
      // ---- likelihood for (row r0 + myr, pair myk) ------------------------------------------
      const int row = r0 + myr;
      u64 dres = dpk(0.f, 0.f);
      if (row < rows) {
        float ea, eb;
        dupk(dadd2(v[0], off), ea, eb);
        const float yv = yt[row];
        float lla, llb, da, db, d1, w0, w1;
        lik_row<float, FAM, false>(yv, ea, 0.f, p.fixed_inv_scale, lla, da, d1, w0, w1);
        lik_row<float, FAM, false>(yv, eb, 0.f, p.fixed_inv_scale, llb, db, d1, w0, w1);
        tla += lla;
        tlb += llb;
        dres = dpk(da, db);
        sd = dadd2(sd, dres);
      }
      if constexpr (GRAD) {
        // ---- all-gather of the residuals: every lane gets all 8 rows x 4 pairs --------------
        v[0] = dres;
#pragma unroll
        for (int half = 1; half <= 16; half <<= 1) {
          const bool up = (lane & half) != 0;
#pragma unroll
          for (int i = 0; i < half; ++i) {
            const u64 mine = v[i];
            const u64 other = shfl_xor64(mine, half);
            v[i] = up ? other : mine;
            v[i + half] = up ? mine : other;
          }
        }
        // ---- backward: gradient slice += x * residual --------------------------------------
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          float xv[CS];
          const float* xr = xt + (r0 + r) * PP;
          if constexpr (CS == 8) {
            const float4 a = *reinterpret_cast<const float4*>(xr), c2 = *reinterpret_cast<const float4*>(xr + 4);
            xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w; xv[4] = c2.x; xv[5] = c2.y; xv[6] = c2.z; xv[7] = c2.w;
          } else if constexpr (CS == 4) {
            const float4 a = *reinterpret_cast<const float4*>(xr);
            xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w;
          } else if constexpr (CS == 2) {
            const float2 a = *reinterpret_cast<const float2*>(xr);
            xv[0] = a.x; xv[1] = a.y;
          } else {
            xv[0] = xr[0];
          }
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int u = 0; u < CS; ++u) g[u][kk] = dfma2(dbc(xv[u]), v[r * 4 + kk], g[u][kk]);
        }
      }
    }
    kahan_add(lpa, lpa_lo, tla);
    kahan_add(lpb, lpb_lo, tlb);
    __syncthreads();  // stage s may be refilled
  }
  cp_async_wait<0>();

  // ---- publish: gradient slices, then log-lik / intercept sums reduced over the 8 row slots -----
  float* mine = p.partial + (size_t)ch * (p.Psmall + LSLB_NACC) * ldc;
  if constexpr (GRAD) {
#pragma unroll
    for (int u = 0; u < CS; ++u) {
      const int j = lane * CS + u;
      if (j < pc) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          *reinterpret_cast<u64*>(mine + (size_t)(D.soff + j) * ldc + cbase + 2 * kk) = g[u][kk];
      }
    }
  }
  float sda, sdb;
  dupk(sd, sda, sdb);
  float lp_a = lpa - lpa_lo, lp_b = lpb - lpb_lo;
#pragma unroll
  for (int m = 4; m <= 16; m <<= 1) {
    lp_a += __shfl_xor_sync(0xffffffffu, lp_a, m);
    lp_b += __shfl_xor_sync(0xffffffffu, lp_b, m);
    sda += __shfl_xor_sync(0xffffffffu, sda, m);
    sdb += __shfl_xor_sync(0xffffffffu, sdb, m);
  }
  if (lane < 4) {
    const int c = cbase + 2 * lane;
    *reinterpret_cast<u64*>(mine + (size_t)p.Psmall * ldc + c) = dpk(lp_a, lp_b);
    *reinterpret_cast<u64*>(mine + (size_t)(p.Psmall + 1) * ldc + c) = dpk(0.f, 0.f);
    if constexpr (GRAD) {
      for (int q = 0; q < p.ni; ++q)
        *reinterpret_cast<u64*>(mine + (size_t)p.ic[q].soff * ldc + c) = dpk(sda, sdb);
    }
  }
}

// host ------------------------------------------------------------------------------------------
bool dense_x2_eligible(const lslb_plan* plan) {
  if (plan->desc.dtype != LSLB_F32) return false;
  if (plan->ns != 0 || plan->grp >= 0 || plan->nd != 1 || plan->has_scale) return false;
  const lslb_block_desc& d = plan->blk[plan->dn[0]];
  if (d.ncoef <= 16 || d.ncoef > 256) return false;
  if ((reinterpret_cast<uintptr_t>(d.data) & 15) != 0 || (d.ld & 3) != 0) return false;
  return true;
}

static int gcd_int(int a, int b) { return b == 0 ? a : gcd_int(b, a % b); }

BandedCfg dense_x2_cfg(const lslb_plan* plan, int C) {
  BandedCfg k;
  k.TC = DWARPS * 8;  // 64 chains per CTA
  k.Cpad = (C + k.TC - 1) / k.TC * k.TC;
  k.ntiles_chain = k.Cpad / k.TC;
  k.ntiles_total = (int)((plan->desc.n + DRT - 1) / DRT);
  if (k.ntiles_total < 1) k.ntiles_total = 1;
  // whole waves: (row chunks x chain tiles) a multiple of the SM count, about 4 CTAs per SM in all
  int unit = plan->sm_count / gcd_int(plan->sm_count, k.ntiles_chain);
  long long want = unit;
  while (want * k.ntiles_chain < 4LL * plan->sm_count) want += unit;
  if (want > k.ntiles_total) want = k.ntiles_total;
  k.tiles_per_cta = (int)((k.ntiles_total + want - 1) / want);
  k.nchunks = (k.ntiles_total + k.tiles_per_cta - 1) / k.tiles_per_cta;
  return k;
}

template <int CS, int FAM>
static int dx2_launch(const EvalParams<float>& ep, const BandedCfg& k, bool g, cudaStream_t st) {
  dim3 grid(k.ntiles_chain, k.nchunks), block(DWARPS * 32);
  const size_t smem = (size_t)DSTAGES * DRT * (32 * CS + 1) * sizeof(float);
  if (g) {
    auto kern = dense_x2_kernel<CS, FAM, true>;
    LSLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, block, smem, st>>>(ep, k.tiles_per_cta, k.ntiles_total);
  } else {
    auto kern = dense_x2_kernel<CS, FAM, false>;
    LSLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, block, smem, st>>>(ep, k.tiles_per_cta, k.ntiles_total);
  }
  LSLB_LAUNCH_CHECK();
  return LSLB_OK;
}

template <int FAM>
static int dx2_family(int pc, const EvalParams<float>& ep, const BandedCfg& k, bool g, cudaStream_t st) {
  if (pc <= 32) return dx2_launch<1, FAM>(ep, k, g, st);
  if (pc <= 64) return dx2_launch<2, FAM>(ep, k, g, st);
  if (pc <= 128) return dx2_launch<4, FAM>(ep, k, g, st);
  return dx2_launch<8, FAM>(ep, k, g, st);
}

int launch_dense_x2(const lslb_plan* plan, const EvalParams<float>& ep, const BandedCfg& k, bool g,
                    cudaStream_t st) {
  const int pc = plan->blk[plan->dn[0]].ncoef;
  switch (plan->fam) {
    case FAM_NORMAL_FIXED: return dx2_family<FAM_NORMAL_FIXED>(pc, ep, k, g, st);
    case FAM_BERNOULLI: return dx2_family<FAM_BERNOULLI>(pc, ep, k, g, st);
    case FAM_POISSON: return dx2_family<FAM_POISSON>(pc, ep, k, g, st);
    default: return LSLB_ERR_UNSUPPORTED;
  }
}

}  // namespace lslb
