// Small dense block (p <= 16, coefficients in registers) + optional random intercept, float32, two
// chains per thread with packed FP32 math (S5: Poisson, p = 8, G = 1e5 groups, n = 5e6; also S1-like
// regressions at large n).
//
//   * 128-row tiles of X (row-major, ld == p), y and the group index are staged by TMA bulk copies
//     through a 3-stage ring; all lanes read the same row -> shared-memory broadcasts, `.F32`
//     broadcast operands for the packed FMAs;
//   * rows are sorted by group and every CTA's row range starts and ends on a group boundary
//     (plan.cu: fine chunk boundaries), so a group's residual sum is accumulated in registers and
//     written exactly once to the [G x Cpad] gradient buffer: a deterministic segment sum without
//     atomics (the reference's `b[idx]` VJP is an XLA scatter-add);
//   * b[g] is read from the transposed [G x Cpad] buffer when the group changes (coalesced 8 B/lane).
#include "common.cuh"
#include "x2_common.cuh"

namespace lslb {

template <int PD>
struct GRaw {
  float x[XR * PD];
  float y[XR];
  int g[XR];
};

// One CTA of up to 16 warps (1024 chains) per SM; its warps share the staged tiles and are NOT
// synchronised per tile: the last warp to leave a stage refills it (see banded_x2.cu for why).
template <int PD, int FAM, bool GROUP, bool GRAD>
__global__ void __launch_bounds__(512, 1)
dgroup_x2_kernel(const __grid_constant__ EvalParams<float> p, int fstride) {
  constexpr int XSTAGES = PD > 8 ? 4 : 6;  // (shadows the 3-stage default of x2_common.cuh)
  __shared__ __align__(128) GRaw<PD> raw[XSTAGES];
  __shared__ __align__(8) u64 full[XSTAGES];
  __shared__ int done[XSTAGES];
  const int tid = threadIdx.x;
  const int c = 2 * (blockIdx.x * blockDim.x + tid);
  const int ch = blockIdx.y;
  const int f0 = ch * fstride, f1 = min(f0 + fstride, p.nchunks);
  const long long r0 = p.chunk_start[f0], r1 = p.chunk_start[f1];
  const int tile0 = (int)(r0 / XR);
  const int my_tiles = r1 > r0 ? (int)((r1 - 1) / XR) - tile0 + 1 : 0;
  const DnsP<float>& D = p.dn[0];

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < XSTAGES; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(x_smem_u32(&full[s])));
      done[s] = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto tile_rows = [&](int t) -> int {
    long long rem = p.n - (long long)t * XR;
    return rem >= XR ? XR : (int)rem;
  };
  auto bulk = [&](void* dst, const void* src, unsigned bytes, u64* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            x_smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(x_smem_u32(bar))
        : "memory");
  };
  auto issue = [&](int t, int s) {
    const long long rr = (long long)t * XR;
    constexpr unsigned bytes = XR * PD * 4 + XR * 4 + (GROUP ? XR * 4 : 0);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(x_smem_u32(&full[s])),
                 "r"(bytes)
                 : "memory");
    bulk(raw[s].x, D.X + rr * PD, XR * PD * 4, &full[s]);
    bulk(raw[s].y, p.y + rr, XR * 4, &full[s]);
    if constexpr (GROUP) bulk(raw[s].g, p.gidx + rr, XR * 4, &full[s]);
  };
  auto wait_full = [&](int s, unsigned parity) {
    unsigned ok = 0;
    while (!ok) {
      asm volatile(
          "{\n.reg .pred q;\nmbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\nselp.u32 %0, 1, 0, q;\n}\n"
          : "=r"(ok)
          : "r"(x_smem_u32(&full[s])), "r"(parity)
          : "memory");
    }
  };
  if (tid == 0) {
    for (int k = 0; k < XSTAGES && k < my_tiles; ++k)
      if (tile_rows(tile0 + k) == XR) issue(tile0 + k, k);
  }

  const size_t ldc = (size_t)p.Cpad;
  auto ld2 = [&](int j) -> u64 { return *reinterpret_cast<const u64*>(p.pT + (size_t)j * ldc + c); };
  float* mine = p.partial + (size_t)ch * (p.Psmall + LSLB_NACC) * ldc + c;
  auto at2 = [&](int j) -> u64* { return reinterpret_cast<u64*>(mine + (size_t)j * ldc); };
  const u64 ZERO = pk(0.f, 0.f);

  const u64 LOG2E = pk(1.4426950408889634f, 1.4426950408889634f);
  const u64 NEG1 = pk(-1.f, -1.f);
  const u64 INVS = pk(p.fixed_inv_scale, p.fixed_inv_scale);
  // Poisson: the predictor is carried as log2(e) * eta (coefficients, offset and group effect are
  // scaled when they are loaded), so the mean is a bare ex2 and sum(y eta) is rescaled per tile
  u64 b[PD], g[PD];
#pragma unroll
  for (int j = 0; j < PD; ++j) {
    b[j] = j < D.p ? ld2(D.soff + j) : ZERO;
    if constexpr (FAM == FAM_POISSON) b[j] = mul2(b[j], LOG2E);
    g[j] = ZERO;
  }
  u64 off = ZERO;
  for (int q = 0; q < p.ni; ++q) off = add2(off, ld2(p.ic[q].soff));
  if constexpr (FAM == FAM_POISSON) off = mul2(off, LOG2E);
  int curg = -1;
  u64 offg = off;  // offset + current group's effect
  u64 accg = ZERO, sd = ZERO;
  float lpa_hi = 0.f, lpa_lo = 0.f, lpb_hi = 0.f, lpb_lo = 0.f;

  for (int k = 0; k < my_tiles; ++k) {
    const int t = tile0 + k, s = k % XSTAGES;
    const int rows = tile_rows(t);
    GRaw<PD>& T = raw[s];
    const long long base = (long long)t * XR;
    if (rows == XR) {
      wait_full(s, (k / XSTAGES) & 1);
    } else {  // ragged last tile of the data
      __syncthreads();  // every warp has left the tile that used this stage before
      for (int r = tid; r < rows; r += blockDim.x) {
        T.y[r] = p.y[base + r];
        if constexpr (GROUP) T.g[r] = p.gidx[base + r];
#pragma unroll
        for (int j = 0; j < PD; ++j) T.x[r * PD + j] = D.X[(base + r) * PD + j];
      }
      __syncthreads();
    }
    // rows of this tile that belong to the CTA's (group-aligned) range
    const int ra = (int)max(r0 - base, 0LL);
    const int rb = (int)min((long long)rows, r1 - base);
    u64 lp2 = ZERO, lpm = ZERO;  // Poisson: lp2 = sum y eta', lpm = sum mu
    (void)lpm;
#pragma unroll 2
    for (int r = ra; r < rb; ++r) {
      float xv[PD];
#pragma unroll
      for (int j4 = 0; j4 < PD; j4 += 4) {
        const float4 v = *reinterpret_cast<const float4*>(&T.x[r * PD + j4]);
        xv[j4] = v.x; xv[j4 + 1] = v.y; xv[j4 + 2] = v.z; xv[j4 + 3] = v.w;
      }
      if constexpr (GROUP) {
        const int gi = T.g[r];
        if (gi != curg) {  // warp-uniform
          if constexpr (GRAD) {
            if (curg >= 0) {
              u64* q = reinterpret_cast<u64*>(p.gbT + (size_t)curg * ldc + c);
              *q = add2(*q, accg);
              sd = add2(sd, accg);
            }
          }
          curg = gi;
          u64 bg = *reinterpret_cast<const u64*>(p.bT + (size_t)gi * ldc + c);
          // the next group's lines: a dependent global load per group change would otherwise stall
          // the warp for a full L2/HBM round trip every ~50 rows
          if (gi + 1 < p.G) {
            asm volatile("prefetch.global.L1 [%0];" ::"l"(p.bT + (size_t)(gi + 1) * ldc + c));
            if constexpr (GRAD) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.gbT + (size_t)(gi + 1) * ldc + c));
          }
          if constexpr (FAM == FAM_POISSON) bg = mul2(bg, LOG2E);
          offg = add2(off, bg);
          accg = ZERO;
        }
      }
      u64 eta = offg;
#pragma unroll
      for (int j = 0; j < PD; ++j) eta = fma2(bc(xv[j]), b[j], eta);
      const float yv = T.y[r];
      u64 d;
      if constexpr (FAM == FAM_POISSON) {
        float ta, tb;
        upk(eta, ta, tb);  // log2(e) * eta
        const u64 mu = pk(ex2f(ta), ex2f(tb));
        lp2 = fma2(bc(yv), eta, lp2);
        lpm = add2(lpm, mu);
        d = fma2(mu, NEG1, bc(yv));  // y - mu
      } else if constexpr (FAM == FAM_NORMAL_FIXED) {
        const u64 z = mul2(fma2(eta, NEG1, bc(yv)), INVS);
        lp2 = fma2(mul2(z, pk(-0.5f, -0.5f)), z, lp2);
        d = mul2(z, INVS);
      } else {
        float ea, eb, lla, llb, da, db, d1, w0, w1;
        upk(eta, ea, eb);
        lik_row<float, FAM, false>(yv, ea, 0.f, p.fixed_inv_scale, lla, da, d1, w0, w1);
        lik_row<float, FAM, false>(yv, eb, 0.f, p.fixed_inv_scale, llb, db, d1, w0, w1);
        lp2 = add2(lp2, pk(lla, llb));
        d = pk(da, db);
      }
      if constexpr (GRAD) {
        if constexpr (GROUP) accg = add2(accg, d);  // (sd collects the groups' sums when they close)
        else sd = add2(sd, d);
#pragma unroll
        for (int j = 0; j < PD; ++j) g[j] = fma2(bc(xv[j]), d, g[j]);
      }
    }
    float la, lb;
    upk(lp2, la, lb);
    if constexpr (FAM == FAM_POISSON) {  // ln2 * sum(y eta') - sum mu
      float ma, mb;
      upk(lpm, ma, mb);
      la = fmaf(0.6931471805599453f, la, -ma);
      lb = fmaf(0.6931471805599453f, lb, -mb);
    }
    kahan_add(lpa_hi, lpa_lo, la);
    kahan_add(lpb_hi, lpb_lo, lb);
    __syncwarp();
    if ((tid & 31) == 0) {
      __threadfence_block();
      if (atomicAdd(&done[s], 1) == (int)(blockDim.x >> 5) - 1) {  // last warp out refills the stage
        done[s] = 0;
        const int kn = k + XSTAGES;
        if (kn < my_tiles && tile_rows(tile0 + kn) == XR) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          issue(tile0 + kn, s);
        }
      }
    }
  }

  if constexpr (GRAD) {
    if constexpr (GROUP) {
      if (curg >= 0) {
        u64* q = reinterpret_cast<u64*>(p.gbT + (size_t)curg * ldc + c);
        *q = add2(*q, accg);
        sd = add2(sd, accg);
      }
    }
#pragma unroll
    for (int j = 0; j < PD; ++j)
      if (j < D.p) *at2(D.soff + j) = g[j];
    for (int q = 0; q < p.ni; ++q) *at2(p.ic[q].soff) = sd;
  }
  *at2(p.Psmall) = pk(lpa_hi, lpb_hi);
  *at2(p.Psmall + 1) = pk(lpa_lo, lpb_lo);
}

// host ------------------------------------------------------------------------------------------
bool dgroup_x2_eligible(const lslb_plan* plan) {
  if (plan->desc.dtype != LSLB_F32) return false;
  if (plan->ns != 0 || plan->nd != 1 || plan->has_scale) return false;
  const lslb_block_desc& d = plan->blk[plan->dn[0]];
  if (d.ncoef > 16 || (d.ncoef & 3) != 0 || d.ld != d.ncoef) return false;
  auto aligned = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (!aligned(d.data) || !aligned(plan->desc.y)) return false;
  if (plan->grp >= 0 && !aligned(plan->blk[plan->grp].index)) return false;
  return plan->desc.n >= 4096;  // tiny problems are launch-latency bound either way
}

LaunchCfg dgroup_x2_cfg(const lslb_plan* plan, int C) {
  LaunchCfg k;
  // chains spread evenly over the chain tiles in whole warps (64 chains), at most 16 warps per CTA;
  // with few chains several CTAs share an SM so that it still holds ~16 warps
  k.ntiles = (C + 1023) / 1024;
  const int per_tile = (C + k.ntiles - 1) / k.ntiles;
  const int U = (per_tile + 63) / 64;
  k.TC = 64 * U;
  k.Cpad = k.ntiles * k.TC;
  int cps = 16 / U;
  if (cps < 1) cps = 1;
  if (cps > 5) cps = 5;  // shared memory: 30-37 KB of row stages per CTA
  long long want = ((long long)plan->sm_count * cps) / k.ntiles;
  if (want < 1) want = 1;
  int F = plan->nchunks;
  k.fstride = (int)((F + want - 1) / want);
  if (k.fstride < 1) k.fstride = 1;
  k.nchunks = (F + k.fstride - 1) / k.fstride;
  return k;
}

template <int PD, int FAM>
static int dg_go(const lslb_plan* plan, const EvalParams<float>& ep, const LaunchCfg& k, bool g,
                 cudaStream_t st) {
  dim3 grid(k.ntiles, k.nchunks), block(k.TC / 2);
  const bool grp = plan->grp >= 0;
  if (grp) {
    if (g) dgroup_x2_kernel<PD, FAM, true, true><<<grid, block, 0, st>>>(ep, k.fstride);
    else dgroup_x2_kernel<PD, FAM, true, false><<<grid, block, 0, st>>>(ep, k.fstride);
  } else {
    if (g) dgroup_x2_kernel<PD, FAM, false, true><<<grid, block, 0, st>>>(ep, k.fstride);
    else dgroup_x2_kernel<PD, FAM, false, false><<<grid, block, 0, st>>>(ep, k.fstride);
  }
  LSLB_LAUNCH_CHECK();
  return LSLB_OK;
}

template <int FAM>
static int dg_family(const lslb_plan* plan, const EvalParams<float>& ep, const LaunchCfg& k, bool g,
                     cudaStream_t st) {
  const int pc = plan->blk[plan->dn[0]].ncoef;
  if (pc == 4) return dg_go<4, FAM>(plan, ep, k, g, st);
  if (pc == 8) return dg_go<8, FAM>(plan, ep, k, g, st);
  if (pc == 12) return dg_go<12, FAM>(plan, ep, k, g, st);
  return dg_go<16, FAM>(plan, ep, k, g, st);
}

int launch_dgroup_x2(const lslb_plan* plan, const EvalParams<float>& ep, const LaunchCfg& k, bool g,
                     cudaStream_t st) {
  switch (plan->fam) {
    case FAM_NORMAL_FIXED: return dg_family<FAM_NORMAL_FIXED>(plan, ep, k, g, st);
    case FAM_BERNOULLI: return dg_family<FAM_BERNOULLI>(plan, ep, k, g, st);
    case FAM_POISSON: return dg_family<FAM_POISSON>(plan, ep, k, g, st);
    default: return LSLB_ERR_UNSUPPORTED;
  }
}

}  // namespace lslb
