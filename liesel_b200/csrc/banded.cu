// Fast path for "intercepts + cubic P-splines" models whose rows are ordered so that the knot
// span of every smooth changes rarely (the plan sorts rows lexicographically by span): S3 / H.
//
//   * grid (chain tiles, row chunks): the CTAs that share a row chunk are adjacent in launch
//     order, so a row tile is fetched from HBM once and re-read from L2 by the other chain tiles;
//   * row tiles of 128 rows are staged in shared memory by 1-D bulk async copies (TMA engine,
//     cp.async.bulk + mbarrier complete_tx) through a 3-stage ring, so HBM/L2 latency is never on
//     the critical path of the FMA pipe;
//   * thread = chain. The 4 active coefficients of each smooth and their 4 gradient accumulators
//     live in registers for a whole run of equal spans; all lanes of a warp read the same row, so
//     row data are shared-memory broadcasts; when the span changes (rare) the accumulators are
//     flushed into this CTA's private slice of the partial buffer (coalesced global RMW);
//   * no atomics: fixed-order reduction over chunks happens in finalize (logp_grad.cu).
#include <stdlib.h>

#include "common.cuh"

namespace lslb {

constexpr int BR = 128;      // rows per tile
constexpr int BSTAGES = 3;   // ring depth

__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes,
                                         unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ float exp_neg(float x) {  // exp(-x) = 2^(-x*log2(e)), one MUFU
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * -1.4426950408889634f));
  return r;
}
__device__ __forceinline__ double exp_neg(double x) { return exp(-x); }

template <typename T, int NS>
struct BandStage {
  T w[NS][BR * 4];
  T y[BR];
  int start[NS][BR];
};

// Normal location-scale gets a hand-scheduled form (the banded configs of BASELINE.json: S3, H); the
// other families use lik_row.
template <typename T, int FAM>
__device__ __forceinline__ void lik_fast(T y, T eta0, T eta1, T fixed_inv_scale, T& ll, T& d0, T& d1) {
  if constexpr (FAM == FAM_NORMAL_LS) {
    const T inv_s = exp_neg(eta1);
    const T z = (y - eta0) * inv_s;
    const T z2 = z * z;
    ll = fma(T(-0.5), z2, -eta1);
    d0 = z * inv_s;
    d1 = z2 - T(1);
  } else {
    T u0, u1;
    lik_row<T, FAM, false>(y, eta0, eta1, fixed_inv_scale, ll, d0, d1, u0, u1);
  }
}

template <typename T, int FAM, int NS, bool GRAD>
__global__ void __launch_bounds__(128, 6)
banded_kernel(const __grid_constant__ EvalParams<T> p, int tiles_per_cta, int ntiles_total, int tiles_extra) {
  __shared__ __align__(128) BandStage<T, NS> st[BSTAGES];
  __shared__ __align__(8) unsigned long long full[BSTAGES];

  const int tid = threadIdx.x;
  const int c = blockIdx.x * blockDim.x + tid;  // < Cpad
  const int ch = blockIdx.y;
  // balanced partition of the tiles over the row chunks (chunk sizes differ by at most one tile)
  const int tile0 = ch * tiles_per_cta + min(ch, tiles_extra);
  const int tile1 = min(tile0 + tiles_per_cta + (ch < tiles_extra ? 1 : 0), ntiles_total);
  const int my_tiles = tile1 - tile0;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < BSTAGES; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto tile_rows = [&](int t) -> int {
    long long r0 = (long long)t * BR;
    long long rem = p.n - r0;
    return rem >= BR ? BR : (int)rem;
  };
  // producer: one thread issues the bulk copies of a full tile; a ragged last tile is copied by
  // all threads with ordinary loads (bulk copies need 16-byte multiples)
  auto issue = [&](int t, int s) {
    const long long r0 = (long long)t * BR;
    constexpr unsigned bytes = BR * sizeof(T) + NS * (BR * 4 * sizeof(T) + BR * sizeof(int));
    mbar_expect_tx(&full[s], bytes);
    bulk_g2s(st[s].y, p.y + r0, BR * sizeof(T), &full[s]);
#pragma unroll
    for (int m = 0; m < NS; ++m) {
      bulk_g2s(st[s].w[m], p.sp[m].w + 4 * r0, BR * 4 * sizeof(T), &full[s]);
      bulk_g2s(st[s].start[m], p.sp[m].start + r0, BR * sizeof(int), &full[s]);
    }
  };

  if (tid == 0) {
    for (int k = 0; k < BSTAGES - 1 && k < my_tiles; ++k)
      if (tile_rows(tile0 + k) == BR) issue(tile0 + k, k);
  }

  // per-chain constants and state
  T o0 = T(0), o1 = T(0);
  for (int q = 0; q < p.ni; ++q) {
    const T v = p.pT[(size_t)p.ic[q].soff * p.Cpad + c];
    if (p.ic[q].pred) o1 += v; else o0 += v;
  }
  T* mine = p.partial + (size_t)ch * (p.Psmall + LSLB_NACC) * p.Cpad + c;
  if constexpr (GRAD) {
    for (int j = 0; j < p.Psmall; ++j) mine[(size_t)j * p.Cpad] = T(0);
  }
  int cur[NS];
  T wb[NS][4], wg[NS][4];
  bool pred1[NS];
#pragma unroll
  for (int m = 0; m < NS; ++m) {
    cur[m] = -1;
    pred1[m] = p.sp[m].pred != 0;
#pragma unroll
    for (int t = 0; t < 4; ++t) { wb[m][t] = T(0); wg[m][t] = T(0); }
  }
  T lp_hi = T(0), lp_lo = T(0), sr0 = T(0), sr1 = T(0);

  auto change = [&](int m, int s) {
    const int so = p.sp[m].soff;
    if constexpr (GRAD) {
      if (cur[m] >= 0) {
#pragma unroll
        for (int t = 0; t < 4; ++t) mine[(size_t)(so + cur[m] + t) * p.Cpad] += wg[m][t];
      }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      wb[m][t] = p.pT[(size_t)(so + s + t) * p.Cpad + c];
      wg[m][t] = T(0);
    }
    cur[m] = s;
  };

  for (int k = 0; k < my_tiles; ++k) {
    const int t = tile0 + k;
    const int s = k % BSTAGES;
    const int rows = tile_rows(t);
    // refill the stage consumed in the previous iteration (all threads passed its barrier)
    const int kn = k + BSTAGES - 1;
    if (tid == 0 && kn < my_tiles && tile_rows(tile0 + kn) == BR) issue(tile0 + kn, kn % BSTAGES);
    if (rows == BR) {
      mbar_wait(&full[s], (k / BSTAGES) & 1);
    } else {
      const long long r0 = (long long)t * BR;
      for (int r = tid; r < rows; r += blockDim.x) {
        st[s].y[r] = p.y[r0 + r];
#pragma unroll
        for (int m = 0; m < NS; ++m) {
          st[s].start[m][r] = p.sp[m].start[r0 + r];
#pragma unroll
          for (int u = 0; u < 4; ++u) st[s].w[m][4 * r + u] = p.sp[m].w[4 * (r0 + r) + u];
        }
      }
      __syncthreads();
    }
    const BandStage<T, NS>& S = st[s];

    // is every smooth's span constant over this tile? (each warp evaluates it redundantly)
    bool uni = true;
#pragma unroll
    for (int m = 0; m < NS; ++m) {
      const int first = S.start[m][0];
      bool eq = true;
      for (int r = (tid & 31); r < rows; r += 32) eq = eq && (S.start[m][r] == first);
      uni = uni && __all_sync(0xffffffffu, eq);
    }

    T lp_tile = T(0);
    if (uni) {
#pragma unroll
      for (int m = 0; m < NS; ++m) {
        const int first = S.start[m][0];
        if (first != cur[m]) change(m, first);
      }
#pragma unroll 4
      for (int r = 0; r < rows; ++r) {
        T eta0 = o0, eta1 = o1;
        Vec4<T> w[NS];
#pragma unroll
        for (int m = 0; m < NS; ++m) {
          w[m] = *reinterpret_cast<const Vec4<T>*>(&S.w[m][4 * r]);
          T v = w[m].a * wb[m][0];
          v = fma(w[m].b, wb[m][1], v);
          v = fma(w[m].c, wb[m][2], v);
          v = fma(w[m].d, wb[m][3], v);
          if (pred1[m]) eta1 += v; else eta0 += v;
        }
        T ll, d0, d1;
        lik_fast<T, FAM>(S.y[r], eta0, eta1, p.fixed_inv_scale, ll, d0, d1);
        lp_tile += ll;
        if constexpr (GRAD) {
          sr0 += d0;
          sr1 += d1;
#pragma unroll
          for (int m = 0; m < NS; ++m) {
            const T rr = pred1[m] ? d1 : d0;
            wg[m][0] = fma(w[m].a, rr, wg[m][0]);
            wg[m][1] = fma(w[m].b, rr, wg[m][1]);
            wg[m][2] = fma(w[m].c, rr, wg[m][2]);
            wg[m][3] = fma(w[m].d, rr, wg[m][3]);
          }
        }
      }
    } else {
      for (int r = 0; r < rows; ++r) {
        T eta0 = o0, eta1 = o1;
        Vec4<T> w[NS];
#pragma unroll
        for (int m = 0; m < NS; ++m) {
          const int sm = S.start[m][r];
          if (sm != cur[m]) change(m, sm);
          w[m] = *reinterpret_cast<const Vec4<T>*>(&S.w[m][4 * r]);
          T v = w[m].a * wb[m][0];
          v = fma(w[m].b, wb[m][1], v);
          v = fma(w[m].c, wb[m][2], v);
          v = fma(w[m].d, wb[m][3], v);
          if (pred1[m]) eta1 += v; else eta0 += v;
        }
        T ll, d0, d1;
        lik_fast<T, FAM>(S.y[r], eta0, eta1, p.fixed_inv_scale, ll, d0, d1);
        lp_tile += ll;
        if constexpr (GRAD) {
          sr0 += d0;
          sr1 += d1;
#pragma unroll
          for (int m = 0; m < NS; ++m) {
            const T rr = pred1[m] ? d1 : d0;
            wg[m][0] = fma(w[m].a, rr, wg[m][0]);
            wg[m][1] = fma(w[m].b, rr, wg[m][1]);
            wg[m][2] = fma(w[m].c, rr, wg[m][2]);
            wg[m][3] = fma(w[m].d, rr, wg[m][3]);
          }
        }
      }
    }
    kahan_add(lp_hi, lp_lo, lp_tile);
    __syncthreads();  // everyone is done with stage s: it may be refilled
  }

  if constexpr (GRAD) {
#pragma unroll
    for (int m = 0; m < NS; ++m) {
      if (cur[m] >= 0) {
#pragma unroll
        for (int t = 0; t < 4; ++t) mine[(size_t)(p.sp[m].soff + cur[m] + t) * p.Cpad] += wg[m][t];
      }
    }
    for (int q = 0; q < p.ni; ++q) mine[(size_t)p.ic[q].soff * p.Cpad] = p.ic[q].pred ? sr1 : sr0;
  }
  mine[(size_t)p.Psmall * p.Cpad] = lp_hi;
  mine[(size_t)(p.Psmall + 1) * p.Cpad] = lp_lo;
}

// host ------------------------------------------------------------------------------------------
bool banded_fast_eligible(const lslb_plan* plan) {
  if (plan->variant != VAR_BANDED || !plan->rows_sorted) return false;
  if (plan->ns > 2) return false;
  auto aligned = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (!aligned(plan->desc.y)) return false;
  for (int m = 0; m < plan->ns; ++m) {
    const lslb_block_desc& d = plan->blk[plan->sp[m]];
    if (!aligned(d.data) || !aligned(d.index)) return false;
  }
  return true;
}

// float32 with more than 32 chains: two chains per thread, packed f32x2 math (banded_x2.cu). Below
// that a warp of chain pairs would be at most half full and the scalar kernel (thread = chain) runs.
// (Until round 2 the threshold was 128 chains: strong-scaling the headline's 1024 chains over 8 GPUs,
// or a NUTS batch compacted to its last running chains, fell back to the scalar kernel at 2.5x the
// pro-rata time; LSLB_X2_MIN_CHAINS restores any threshold for A/B runs.)
static bool use_x2(const lslb_plan* plan, int C) {
  static const int min_chains = [] {
    const char* e = getenv("LSLB_X2_MIN_CHAINS");
    return e ? atoi(e) : 33;
  }();
  return plan->desc.dtype == LSLB_F32 && C >= min_chains;
}

BandedCfg banded_cfg(const lslb_plan* plan, int C) {
  BandedCfg k;
  const bool x2 = use_x2(plan, C);
  k.TC = x2 ? 256 : 128;  // chains per CTA
  k.Cpad = (C + k.TC - 1) / k.TC * k.TC;
  k.ntiles_chain = k.Cpad / k.TC;
  if (x2) {
    // banded_x2: ONE CTA of up to 16 warps (1024 chains) per SM whose warps share the staged row
    // tiles and drift at most LSLB_X2_STAGES tiles apart; chains are spread evenly over the chain
    // tiles in whole warps (64 chains), and with few chains several CTAs share an SM so that it
    // still holds ~16 warps. (A sampler that drops finished chains calls with any C.)
    k.ntiles_chain = (C + 1023) / 1024;
    const int per_tile = (C + k.ntiles_chain - 1) / k.ntiles_chain;
    k.tc_eff = (per_tile + 63) / 64 * 64;
    k.Cpad = k.ntiles_chain * k.tc_eff;
  }
  k.ntiles_total = (int)((plan->desc.n + BR - 1) / BR);
  if (k.ntiles_total < 1) k.ntiles_total = 1;
  int ctas_per_sm = x2 ? 16 / (k.tc_eff / 64) : 6;  // x2: at most 16 resident warps per SM (registers)
  if (x2) {
    // 2..8 chain warps per chain tile (65..512 chains) on a plan whose two smooths share their band: ONE phased
    // CTA per SM - the 16 warps split into 2 / 4 / 8 row phases that take the tiles of one TMA ring in turn
    // (banded_x2.cu). A/B on one box, H model, ms per call against the several-small-CTAs launch of round 2:
    // 128 chains 0.134 -> 0.130, 256: 0.224 -> 0.211, 512: 0.378 -> 0.370; at 64 chains (16 one-warp phases)
    // it loses 2 %, and the two-band instantiation spills (150 B) and loses 2-7 %, so those keep the small
    // CTAs (ring of 6 / 3 / 2 stages for 34 / 17 / 11 KB CTAs). LSLB_X2_NO_PHASES=1 disables it for A/B runs.
    static const bool no_phases = [] {
      const char* e = getenv("LSLB_X2_NO_PHASES");
      return e && atoi(e) != 0;
    }();
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    k.stages = ctas_per_sm >= 16 ? 2 : (ctas_per_sm >= 8 ? 3 : 6);
    if (!no_phases && plan->shared_band && plan->ns == 2 && ctas_per_sm >= 2 && ctas_per_sm <= 8) {
      int ph = 1;
      while (ph * 2 <= ctas_per_sm) ph *= 2;  // power of two: the ring depth (16) must be a multiple of it
      k.phases = ph;
      k.stages = 16;
      ctas_per_sm = 1;
    }
  } else if (ctas_per_sm > 5) {
    ctas_per_sm = 5;  // shared memory of the scalar kernel's row stages
  }
  static const int knob_ctas = [] {  // tuning knob, read once
    const char* e = getenv("LSLB_BANDED_CTAS_PER_SM");
    return e ? atoi(e) : 0;
  }();
  if (knob_ctas > 0) ctas_per_sm = knob_ctas;
  long long want = ((long long)plan->sm_count * ctas_per_sm + k.ntiles_chain - 1) / k.ntiles_chain;
  // x2 CTAs hold the SM's whole register file: one CTA too many would run as a second wave
  if (x2) want = ((long long)plan->sm_count * ctas_per_sm) / k.ntiles_chain;
  if (want < 1) want = 1;
  if (want > k.ntiles_total) want = k.ntiles_total;
  k.tiles_per_cta = (int)((k.ntiles_total + want - 1) / want);
  // balanced partition: chunk sizes differ by at most one tile (with ceil(T / CTAs) tiles per CTA the
  // last SMs of a 4-CTAs-per-SM launch held 3 CTAs and the others carried 56 tiles instead of 52.8:
  // 2-4 % at 64..512 chains, A/B on one box)
  k.nchunks = (int)want;
  k.tiles_per_cta = k.ntiles_total / k.nchunks;
  k.tiles_extra = k.ntiles_total % k.nchunks;
  k.nchunks *= k.phases;  // partial slots: one per CTA and row phase
  return k;
}

template <typename T, int FAM, int NS>
static int banded_launch(const EvalParams<T>& ep, const BandedCfg& k, bool g, cudaStream_t st) {
  dim3 grid(k.ntiles_chain, k.nchunks), block(k.TC);
  if (g) banded_kernel<T, FAM, NS, true><<<grid, block, 0, st>>>(ep, k.tiles_per_cta, k.ntiles_total, k.tiles_extra);
  else banded_kernel<T, FAM, NS, false><<<grid, block, 0, st>>>(ep, k.tiles_per_cta, k.ntiles_total, k.tiles_extra);
  LSLB_LAUNCH_CHECK();
  return LSLB_OK;
}

int launch_banded_x2(const lslb_plan* plan, const EvalParams<float>& ep, const BandedCfg& k, bool g,
                     cudaStream_t st);  // banded_x2.cu

template <typename T>
int launch_banded(const lslb_plan* plan, const EvalParams<T>& ep, const BandedCfg& k, bool g,
                  cudaStream_t st) {
  if constexpr (sizeof(T) == 4) {
    if (k.TC == 256) return launch_banded_x2(plan, ep, k, g, st);
  }
  if (plan->fam == FAM_NORMAL_LS) {
    if (plan->ns == 1) return banded_launch<T, FAM_NORMAL_LS, 1>(ep, k, g, st);
    return banded_launch<T, FAM_NORMAL_LS, 2>(ep, k, g, st);
  }
  if (plan->fam == FAM_POISSON) {
    if (plan->ns == 1) return banded_launch<T, FAM_POISSON, 1>(ep, k, g, st);
    return banded_launch<T, FAM_POISSON, 2>(ep, k, g, st);
  }
  if (plan->fam == FAM_BERNOULLI) {
    if (plan->ns == 1) return banded_launch<T, FAM_BERNOULLI, 1>(ep, k, g, st);
    return banded_launch<T, FAM_BERNOULLI, 2>(ep, k, g, st);
  }
  if (plan->ns == 1) return banded_launch<T, FAM_NORMAL_FIXED, 1>(ep, k, g, st);
  return banded_launch<T, FAM_NORMAL_FIXED, 2>(ep, k, g, st);
}
template int launch_banded<float>(const lslb_plan*, const EvalParams<float>&, const BandedCfg&, bool,
                                  cudaStream_t);
template int launch_banded<double>(const lslb_plan*, const EvalParams<double>&, const BandedCfg&,
                                   bool, cudaStream_t);

}  // namespace lslb
