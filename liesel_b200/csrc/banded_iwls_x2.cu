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
#include "common.cuh"
#include "x2_common.cuh"

namespace lslb {

// TK: 0 = spline target (index TM among the smooths), 1 = intercept target; TP = predictor of the target
template <int FAM, int NS, int PM, int TK, int TM, int TP>
__global__ void __launch_bounds__(128, 4)
banded_iwls_x2_kernel(const __grid_constant__ EvalParams<float> p, int tp_rows /*p of target*/,
                      int nq, float* __restrict__ partial, int tiles_per_cta, int ntiles_total) {
  __shared__ __align__(128) XRaw<NS> raw[XSTAGES];
  __shared__ __align__(8) u64 full[XSTAGES];
  const int tid = threadIdx.x;
  const int c = 2 * (blockIdx.x * blockDim.x + tid);
  const int ch = blockIdx.y;
  const int tile0 = ch * tiles_per_cta;
  const int my_tiles = min(tile0 + tiles_per_cta, ntiles_total) - tile0;
  XPipe<NS> pipe{raw, full, &p};
  pipe.init(tid);
  if (tid == 0) {
    for (int k = 0; k < XSTAGES - 1 && k < my_tiles; ++k)
      if (pipe.tile_rows(tile0 + k) == XR) pipe.issue(tile0 + k, k);
  }

  const size_t ldc = (size_t)p.Cpad;
  auto ld2 = [&](int j) -> u64 { return *reinterpret_cast<const u64*>(p.pT + (size_t)j * ldc + c); };
  float* mine = partial + (size_t)ch * nq * ldc + c;
  auto at2 = [&](int j) -> u64* { return reinterpret_cast<u64*>(mine + (size_t)j * ldc); };
  const u64 ZERO = pk(0.f, 0.f);

  u64 o0 = ZERO, o1 = ZERO;
  for (int q = 0; q < p.ni; ++q) {
    const u64 v = ld2(p.ic[q].soff);
    if (p.ic[q].pred) o1 = add2(o1, v); else o0 = add2(o0, v);
  }
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
  const u64 NEG1 = pk(-1.f, -1.f);
  const u64 TWO = pk(2.f, 2.f);
  const u64 NLOG2E = pk(-1.4426950408889634f, -1.4426950408889634f);
  const u64 INVS_FIXED = pk(p.fixed_inv_scale, p.fixed_inv_scale);

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
  auto change = [&](int m, int s) {
    if (TK == 0 && m == TM) flush_target(cur[m]);
    const int so = p.sp[m].soff;
#pragma unroll
    for (int t = 0; t < 4; ++t) wb[m][t] = ld2(so + s + t);
    cur[m] = s;
  };

  for (int k = 0; k < my_tiles; ++k) {
    const int t = tile0 + k, s = k % XSTAGES;
    const int rows = pipe.tile_rows(t);
    {
      const int kn = k + XSTAGES - 1;
      if (tid == 0 && kn < my_tiles && pipe.tile_rows(tile0 + kn) == XR) pipe.issue(tile0 + kn, kn % XSTAGES);
    }
    XRaw<NS>& D = raw[s];
    if (rows == XR) {
      pipe.wait_full(s, (k / XSTAGES) & 1);
    } else {
      pipe.load_ragged(t, s, rows, tid, blockDim.x);
      __syncthreads();
    }
    bool uni = true;
#pragma unroll
    for (int m = 0; m < NS; ++m) {
      const int first = D.start[m][0];
      bool eq = true;
      for (int r = (tid & 31); r < rows; r += 32) eq = eq && (D.start[m][r] == first);
      uni = uni && __all_sync(0xffffffffu, eq);
    }

    auto row = [&](int r) {
      float w[NS][4];
      u64 eta0 = o0, eta1 = o1;
#pragma unroll
      for (int m = 0; m < NS; ++m) {
        const float4 v = *reinterpret_cast<const float4*>(&D.w[m][4 * r]);
        w[m][0] = v.x; w[m][1] = v.y; w[m][2] = v.z; w[m][3] = v.w;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if ((PM >> m) & 1) eta1 = fma2(bc(w[m][u]), wb[m][u], eta1);
          else eta0 = fma2(bc(w[m][u]), wb[m][u], eta0);
        }
      }
      const u64 d = fma2(eta0, NEG1, bc(D.y[r]));
      u64 inv_s;
      if constexpr (FAM == FAM_NORMAL_LS) {
        const u64 tt = mul2(eta1, NLOG2E);
        float ta, tb;
        upk(tt, ta, tb);
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
        u64 wa[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          ts[a] = fma2(bc(w[TM][a]), rr, ts[a]);
          wa[a] = mul2(bc(w[TM][a]), W);
        }
        int u = 0;
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = a; b < 4; ++b) {
            tg[u] = fma2(bc(w[TM][b]), wa[a], tg[u]);
            ++u;
          }
      } else {
        is = add2(is, rr);
        ig = add2(ig, W);
      }
    };

    if (uni) {
#pragma unroll
      for (int m = 0; m < NS; ++m) {
        const int first = D.start[m][0];
        if (first != cur[m]) change(m, first);
      }
#pragma unroll 4
      for (int r = 0; r < rows; ++r) row(r);
    } else {
      for (int r = 0; r < rows; ++r) {
#pragma unroll
        for (int m = 0; m < NS; ++m) {
          const int sm = D.start[m][r];
          if (sm != cur[m]) change(m, sm);
        }
        row(r);
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
  banded_iwls_x2_kernel<FAM, NS, PM, TK, TM, TP><<<grid, block, 0, st>>>(ep, p, nq, partial,
                                                                         k.tiles_per_cta, k.ntiles_total);
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
  long long want = ((long long)plan->sm_count * 4 + k.ntiles_chain - 1) / k.ntiles_chain;
  if (want > k.ntiles_total) want = k.ntiles_total;
  if (want < 1) want = 1;
  k.tiles_per_cta = (int)((k.ntiles_total + want - 1) / want);
  k.nchunks = (k.ntiles_total + k.tiles_per_cta - 1) / k.tiles_per_cta;
  return k;
}

}  // namespace lslb
