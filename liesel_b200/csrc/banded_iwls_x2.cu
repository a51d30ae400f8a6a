// IWLS score + observed information of one block of a banded Gaussian model (S3: "IWLSKernel blocks,
// 256 chains"), two chains per thread with packed FP32 math; same TMA-staged run-length scheme as
// banded_x2.cu. One pass over the rows yields what the reference obtains from grad(logp) and
// -jacfwd(grad(logp)) (src/liesel/goose/iwls.py:315-321):
//   spline target : score[4] and the upper triangle of the 4x4 block  w w^T W  accumulate in
//                   registers for a run of equal spans, then flush into the band of the CTA's
//                   partial slice (row p + 4a + d holds entry (a, a+d));
//   intercept     : sum r and sum W.
// W is the OBSERVED information weight: 1/sigma^2 for the location predictor, 2 z^2 for the
// log-scale predictor (reference uses the AD Hessian, not the expected information).
#include <type_traits>

#include "common.cuh"
#include "x2_common.cuh"

namespace lslb {

// What the TMA engine writes per tile for this kernel: the evaluation tile plus, for a spline target,
// the target smooth's ten weight products w_a w_b (a <= b; row-constant, shared by every chain).
template <int NS, bool PAIRS>
struct IwRaw {
  XRaw<NS> x;
  float wp[PAIRS ? XR * 12 : 4];  // [row][12]: (0,0) (0,1) (0,2) (0,3) (1,1) (1,2) (1,3) (2,2) (2,3) (3,3) - -
};

// TK: 0 = spline target (index TM among the smooths), 1 = intercept target; TP = predictor of the target
// PAIRS: the plan holds the target's weight products (lslb_plan::d_wp): the Gram update is ten FMAs
//        on a broadcast scalar instead of four multiplies and ten FMAs.
// UNITY: every row's weights sum to one (plan->unit_rows): on tiles of constant span the predictors
//        take the three-term form of banded_x2.cu (b_u - b_1 against w_u, u != 1).
// Packed ops per row and chain pair (Normal location-scale, location target): forward 6 (8 without
// UNITY), d / z / r / W 4, score 4, Gram 10 (14 without PAIRS) = 24 (31 before round 2).
// NSL: bands staged (1 = both smooths share spans and weights: same covariate, same knots; see banded_x2.cu)
template <int FAM, int NS, int PM, int TK, int TM, int TP, bool PAIRS, bool UNITY, int NSL = NS>
__global__ void __launch_bounds__(128, 4)
banded_iwls_x2_kernel(const __grid_constant__ EvalParams<float> p, int tp_rows /*p of target*/,
                      int nq, float* __restrict__ partial, int tiles_per_cta, int ntiles_total, int tiles_extra) {
  constexpr bool SHW = NSL < NS;
  __shared__ __align__(128) IwRaw<NSL, PAIRS> raw[XSTAGES];
  __shared__ __align__(8) u64 full[XSTAGES];
  const int tid = threadIdx.x;
  const int c = 2 * (blockIdx.x * blockDim.x + tid);
  const int ch = blockIdx.y;
  // balanced partition of the tiles over the row chunks (chunk sizes differ by at most one tile; with
  // ceil(T / CTAs) tiles per CTA the last SMs of a 4-CTAs-per-SM launch held 3 CTAs and the others
  // carried 56 tiles instead of 52.8)
  // tiles_per_cta = floor(T / chunks); the first `tiles_extra` chunks take one tile more. (Plain int
  // arithmetic on purpose: the hot instantiation sits at 128 registers and a 64-bit division here made
  // it spill - 5 % on H.)
  // (written as ONE multiply of the block index by a selected constant: the hot instantiation sits at
  //  128 registers and spilled - 5 % on H - with min(ch, extra) or a 64-bit division here)
  const int tpa = ch < tiles_extra ? tiles_per_cta + 1 : tiles_per_cta;
  const int tile0 = ch * tpa + (ch < tiles_extra ? 0 : tiles_extra);
  const int my_tiles = tpa;
  (void)ntiles_total;

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
    const long long r0 = (long long)t * XR;
    constexpr unsigned bytes = XR * 4 + NSL * (XR * 16 + XR * 4) + (PAIRS ? XR * 48 : 0);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(x_smem_u32(&full[s])),
                 "r"(bytes)
                 : "memory");
    bulk(raw[s].x.y, p.y + r0, XR * 4, &full[s]);
#pragma unroll
    for (int m = 0; m < NSL; ++m) {
      bulk(raw[s].x.w[m], p.sp[m].w + 4 * r0, XR * 16, &full[s]);
      bulk(raw[s].x.start[m], p.sp[m].start + r0, XR * 4, &full[s]);
    }
    if constexpr (PAIRS) bulk(raw[s].wp, p.sp[TM].wp + 12 * r0, XR * 48, &full[s]);
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
#pragma unroll
    for (int s = 0; s < XSTAGES; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(x_smem_u32(&full[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    for (int k = 0; k < XSTAGES - 1 && k < my_tiles; ++k)
      if (tile_rows(tile0 + k) == XR) issue(tile0 + k, k);
  }

  const size_t ldc = (size_t)p.Cpad;
  auto ld2 = [&](int j) -> u64 { return *reinterpret_cast<const u64*>(p.pT + (size_t)j * ldc + c); };
  float* mine = partial + (size_t)ch * nq * ldc + c;
  auto at2 = [&](int j) -> u64* { return reinterpret_cast<u64*>(mine + (size_t)j * ldc); };
  const u64 ZERO = pk(0.f, 0.f);
  const u64 NEG1 = pk(-1.f, -1.f);
  const u64 TWO = pk(2.f, 2.f);
  const u64 NLOG2E = pk(-1.4426950408889634f, -1.4426950408889634f);
  const u64 INVS_FIXED = pk(p.fixed_inv_scale, p.fixed_inv_scale);

  u64 o0 = ZERO, o1 = ZERO;
  for (int q = 0; q < p.ni; ++q) {
    const u64 v = ld2(p.ic[q].soff);
    if (p.ic[q].pred) o1 = add2(o1, v); else o0 = add2(o0, v);
  }
  // Normal location-scale: predictor 1 is carried as -log2(e) * eta1 (coefficients scaled when a
  // run starts), so 1/sigma = exp(-eta1) is a bare ex2
  if constexpr (FAM == FAM_NORMAL_LS) o1 = mul2(o1, NLOG2E);
  for (int j = 0; j < nq; ++j) *at2(j) = ZERO;
  int cur[NS];
  u64 wb[NS][4];
#pragma unroll
  for (int m = 0; m < NS; ++m) {
    cur[m] = -1;
#pragma unroll
    for (int t = 0; t < 4; ++t) wb[m][t] = ZERO;
  }
  u64 ts[4] = {ZERO, ZERO, ZERO, ZERO};
  u64 tg[10];
#pragma unroll
  for (int t = 0; t < 10; ++t) tg[t] = ZERO;
  u64 is = ZERO, ig = ZERO;
  bool fastmode = false;   // wb currently holds the three-term (UNITY) form
  u64 e0i = o0, e1i = o1;  // predictor start values in that form (offsets + sum of the b_1)

  auto flush_target = [&](int s) {  // s = span the accumulators belong to
    if (s < 0) return;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      u64* q = at2(s + a);
      *q = add2(*q, ts[a]);
      ts[a] = ZERO;
    }
    int u = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = a; b < 4; ++b) {
        u64* q = at2(tp_rows + 4 * (s + a) + (b - a));
        *q = add2(*q, tg[u]);
        tg[u] = ZERO;
        ++u;
      }
  };
  auto change = [&](int m, int s, bool fast) {
    if (TK == 0 && m == TM) flush_target(cur[m]);
    const int so = p.sp[m].soff;
#pragma unroll
    for (int t = 0; t < 4; ++t) wb[m][t] = ld2(so + s + t);
    if constexpr (FAM == FAM_NORMAL_LS) {
      if ((PM >> m) & 1) {
#pragma unroll
        for (int t = 0; t < 4; ++t) wb[m][t] = mul2(wb[m][t], NLOG2E);
      }
    }
    if (fast) {
      const u64 nb1 = mul2(wb[m][1], NEG1);
      wb[m][0] = add2(wb[m][0], nb1);
      wb[m][2] = add2(wb[m][2], nb1);
      wb[m][3] = add2(wb[m][3], nb1);
    }
    cur[m] = s;
  };

  for (int k = 0; k < my_tiles; ++k) {
    const int t = tile0 + k, s = k % XSTAGES;
    const int rows = tile_rows(t);
    {
      const int kn = k + XSTAGES - 1;
      if (tid == 0 && kn < my_tiles && tile_rows(tile0 + kn) == XR) issue(tile0 + kn, kn % XSTAGES);
    }
    XRaw<NSL>& D = raw[s].x;
    const float* WP = raw[s].wp;
    if (rows == XR) {
      wait_full(s, (k / XSTAGES) & 1);
    } else {  // ragged last tile: ordinary loads (bulk copies need 16-byte multiples)
      const long long r0 = (long long)t * XR;
      for (int r = tid; r < rows; r += blockDim.x) {
        D.y[r] = p.y[r0 + r];
#pragma unroll
        for (int m = 0; m < NSL; ++m) {
          D.start[m][r] = p.sp[m].start[r0 + r];
#pragma unroll
          for (int u = 0; u < 4; ++u) D.w[m][4 * r + u] = p.sp[m].w[4 * (r0 + r) + u];
        }
        if constexpr (PAIRS) {
#pragma unroll
          for (int u = 0; u < 12; ++u) raw[s].wp[12 * r + u] = p.sp[TM].wp[12 * (r0 + r) + u];
        }
      }
      __syncthreads();
    }
    bool uni = true;
#pragma unroll
    for (int m = 0; m < NSL; ++m) {
      const int first = D.start[m][0];
      bool eq = true;
      for (int r = (tid & 31); r < rows; r += 32) eq = eq && (D.start[m][r] == first);
      uni = uni && __all_sync(0xffffffffu, eq);
    }

    auto row = [&](int r, auto FASTtag) {
      constexpr bool FAST = decltype(FASTtag)::value;
      float w[NSL][4];
      u64 eta0 = FAST ? e0i : o0, eta1 = FAST ? e1i : o1;
#pragma unroll
      for (int m = 0; m < NSL; ++m) {
        const float4 v = *reinterpret_cast<const float4*>(&D.w[m][4 * r]);
        w[m][0] = v.x; w[m][1] = v.y; w[m][2] = v.z; w[m][3] = v.w;
      }
#pragma unroll
      for (int m = 0; m < NS; ++m) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (FAST && u == 1) continue;
          if ((PM >> m) & 1) eta1 = fma2(bc(w[SHW ? 0 : m][u]), wb[m][u], eta1);
          else eta0 = fma2(bc(w[SHW ? 0 : m][u]), wb[m][u], eta0);
        }
      }
      const u64 d = fma2(eta0, NEG1, bc(D.y[r]));
      u64 inv_s;
      if constexpr (FAM == FAM_NORMAL_LS) {
        float ta, tb;
        upk(eta1, ta, tb);  // already -log2(e) * eta1
        inv_s = pk(ex2f(ta), ex2f(tb));
      } else {
        inv_s = INVS_FIXED;
      }
      const u64 z = mul2(d, inv_s);
      u64 rr, W;
      if constexpr (TP == 0) {
        rr = mul2(z, inv_s);      // dl/d eta0 = z / sigma
        W = mul2(inv_s, inv_s);   // 1 / sigma^2
      } else {
        const u64 z2 = mul2(z, z);
        rr = add2(z2, NEG1);      // z^2 - 1
        W = mul2(z2, TWO);        // observed information 2 z^2
      }
      if constexpr (TK == 0) {
#pragma unroll
        for (int a = 0; a < 4; ++a) ts[a] = fma2(bc(w[SHW ? 0 : TM][a]), rr, ts[a]);
        if constexpr (PAIRS) {
          const float4 p0 = *reinterpret_cast<const float4*>(&WP[12 * r]);
          const float4 p1 = *reinterpret_cast<const float4*>(&WP[12 * r + 4]);
          const float2 p2 = *reinterpret_cast<const float2*>(&WP[12 * r + 8]);
          const float pp[10] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w, p2.x, p2.y};
#pragma unroll
          for (int u = 0; u < 10; ++u) tg[u] = fma2(bc(pp[u]), W, tg[u]);
        } else {
          u64 wa[4];
#pragma unroll
          for (int a = 0; a < 4; ++a) wa[a] = mul2(bc(w[SHW ? 0 : TM][a]), W);
          int u = 0;
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = a; b < 4; ++b) {
              tg[u] = fma2(bc(w[SHW ? 0 : TM][b]), wa[a], tg[u]);
              ++u;
            }
        }
      } else {
        is = add2(is, rr);
        ig = add2(ig, W);
      }
    };
    using TFast = std::integral_constant<bool, true>;
    using TFull = std::integral_constant<bool, false>;

    const bool want_fast = UNITY && uni && rows == XR;
    if (want_fast != fastmode) {  // the coefficient registers change form: reload them on next use
#pragma unroll
      for (int m = 0; m < NS; ++m) {
        if (TK == 0 && m == TM) flush_target(cur[m]);
        cur[m] = -1;
      }
      fastmode = want_fast;
    }
    if (uni) {
      bool reloaded = false;
#pragma unroll
      for (int m = 0; m < NS; ++m) {
        const int first = D.start[SHW ? 0 : m][0];
        if (first != cur[m]) {
          change(m, first, fastmode);
          reloaded = true;
        }
      }
      if (fastmode) {
        if (reloaded) {
          e0i = o0;
          e1i = o1;
#pragma unroll
          for (int m = 0; m < NS; ++m) {
            if ((PM >> m) & 1) e1i = add2(e1i, wb[m][1]);
            else e0i = add2(e0i, wb[m][1]);
          }
        }
#pragma unroll 4
        for (int r = 0; r < XR; ++r) row(r, TFast{});
      } else {
#pragma unroll 4
        for (int r = 0; r < rows; ++r) row(r, TFull{});
      }
    } else {
      for (int r = 0; r < rows; ++r) {
#pragma unroll
        for (int m = 0; m < NS; ++m) {
          const int sm = D.start[SHW ? 0 : m][r];
          if (sm != cur[m]) change(m, sm, false);
        }
        row(r, TFull{});
      }
    }
    __syncthreads();
  }
  if constexpr (TK == 0) {
    flush_target(cur[TM]);
  } else {
    *at2(0) = is;
    *at2(1) = ig;
  }
}

// host ------------------------------------------------------------------------------------------
template <int FAM, int NS, int PM, int TK, int TM, int TP>
static int iwx2_go(const EvalParams<float>& ep, int p, int nq, float* partial, const BandedCfg& k,
                   cudaStream_t st) {
  dim3 grid(k.ntiles_chain, k.nchunks), block(128);
  const bool pairs = TK == 0 && ep.sp[TM].wp != nullptr;
  const bool unity = (k.stages & 1) != 0;  // banded_iwls_cfg: bit 0 plan->unit_rows, bit 1 plan->shared_band
  const bool shw = (k.stages & 2) != 0;
#define LSLB_IW_GO2(PR, UN, NL)                                                                        \
  banded_iwls_x2_kernel<FAM, NS, PM, TK, TM, TP, PR, UN, NL><<<grid, block, 0, st>>>(ep, p, nq, partial,  \
                                                                                     k.tiles_per_cta, k.ntiles_total, k.tiles_extra)
#define LSLB_IW_GO(PR, UN)                      \
  do {                                          \
    if constexpr (NS == 2) {                    \
      if (shw) LSLB_IW_GO2(PR, UN, 1);          \
      else LSLB_IW_GO2(PR, UN, NS);             \
    } else {                                    \
      LSLB_IW_GO2(PR, UN, NS);                  \
    }                                           \
  } while (0)
  if constexpr (TK == 0) {
    if (pairs) { if (unity) LSLB_IW_GO(true, true); else LSLB_IW_GO(true, false); }
    else { if (unity) LSLB_IW_GO(false, true); else LSLB_IW_GO(false, false); }
  } else {
    if (unity) LSLB_IW_GO(false, true); else LSLB_IW_GO(false, false);
  }
#undef LSLB_IW_GO2
#undef LSLB_IW_GO
  LSLB_LAUNCH_CHECK();
  return LSLB_OK;
}

template <int FAM, int NS, int PM>
static int iwx2_target(int tk, int tm, int tp, const EvalParams<float>& ep, int p, int nq, float* partial,
                       const BandedCfg& k, cudaStream_t st) {
  if (tk == 1) {
    if (tp == 0) return iwx2_go<FAM, NS, PM, 1, 0, 0>(ep, p, nq, partial, k, st);
    if constexpr (FAM == FAM_NORMAL_LS) return iwx2_go<FAM, NS, PM, 1, 0, 1>(ep, p, nq, partial, k, st);
    return LSLB_ERR_INVALID;
  }
  if (tm == 0) {
    if constexpr ((PM & 1) == 0) return iwx2_go<FAM, NS, PM, 0, 0, 0>(ep, p, nq, partial, k, st);
    else return iwx2_go<FAM, NS, PM, 0, 0, 1>(ep, p, nq, partial, k, st);
  }
  if constexpr (NS >= 2) {
    if constexpr (((PM >> 1) & 1) == 0) return iwx2_go<FAM, NS, PM, 0, 1, 0>(ep, p, nq, partial, k, st);
    else return iwx2_go<FAM, NS, PM, 0, 1, 1>(ep, p, nq, partial, k, st);
  }
  return LSLB_ERR_INVALID;
}

// tk: 0 spline / 1 intercept; tm: smooth index; tp: predictor of the target block
int launch_banded_iwls_x2(const lslb_plan* plan, const EvalParams<float>& ep, int tk, int tm, int tp,
                          int p, int nq, float* partial, const BandedCfg& k, cudaStream_t st) {
  int pm = 0;
  for (int m = 0; m < plan->ns; ++m) pm |= (plan->blk[plan->sp[m]].predictor & 1) << m;
  if (plan->fam == FAM_NORMAL_LS) {
    if (plan->ns == 1)
      return pm ? iwx2_target<FAM_NORMAL_LS, 1, 1>(tk, tm, tp, ep, p, nq, partial, k, st)
                : iwx2_target<FAM_NORMAL_LS, 1, 0>(tk, tm, tp, ep, p, nq, partial, k, st);
    switch (pm) {
      case 0: return iwx2_target<FAM_NORMAL_LS, 2, 0>(tk, tm, tp, ep, p, nq, partial, k, st);
      case 1: return iwx2_target<FAM_NORMAL_LS, 2, 1>(tk, tm, tp, ep, p, nq, partial, k, st);
      case 2: return iwx2_target<FAM_NORMAL_LS, 2, 2>(tk, tm, tp, ep, p, nq, partial, k, st);
      default: return iwx2_target<FAM_NORMAL_LS, 2, 3>(tk, tm, tp, ep, p, nq, partial, k, st);
    }
  }
  if (plan->ns == 1) return iwx2_target<FAM_NORMAL_FIXED, 1, 0>(tk, tm, tp, ep, p, nq, partial, k, st);
  return iwx2_target<FAM_NORMAL_FIXED, 2, 0>(tk, tm, tp, ep, p, nq, partial, k, st);
}

BandedCfg banded_iwls_cfg(const lslb_plan* plan, int C) {
  BandedCfg k;
  k.TC = 256;
  k.Cpad = (C + k.TC - 1) / k.TC * k.TC;
  k.ntiles_chain = k.Cpad / k.TC;
  k.ntiles_total = (int)((plan->desc.n + XR - 1) / XR);
  if (k.ntiles_total < 1) k.ntiles_total = 1;
  // this kernel's ring depth is fixed (XSTAGES); the field carries the two plan flags
  k.stages = (plan->unit_rows ? 1 : 0) | (plan->shared_band ? 2 : 0);
  long long want = ((long long)plan->sm_count * 4 + k.ntiles_chain - 1) / k.ntiles_chain;
  if (want > k.ntiles_total) want = k.ntiles_total;
  if (want < 1) want = 1;
  k.tiles_per_cta = (int)((k.ntiles_total + want - 1) / want);
  // balanced partition: chunk sizes differ by at most one tile (with ceil(T / CTAs) tiles per CTA the
  // last SMs of a 4-CTAs-per-SM launch held 3 CTAs and the others carried 56 tiles instead of 52.8:
  // 2-4 % at 64..512 chains, A/B on one box)
  k.nchunks = (int)want;
  k.tiles_per_cta = k.ntiles_total / k.nchunks;
  k.tiles_extra = k.ntiles_total % k.nchunks;
  return k;
}

}  // namespace lslb
