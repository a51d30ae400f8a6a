// Headline kernel (H / S3, float32): the run-length banded evaluation of banded.cu with TWO chains
// per thread, computed with Blackwell's packed FP32 instructions (fma/mul/add .f32x2 -> FFMA2 etc.).
//
// Why packed: a packed instruction does two chains' worth of FP32 work per issue slot, which frees
// issue bandwidth for the shared-memory broadcasts and MUFU ops that accompany the 16 FMAs of a row.
//
// Data flow per CTA (up to 512 threads = 16 warps = 1024 chains; ONE CTA per SM, or several smaller
// ones when the chain count is low; each CTA owns one chunk of rows):
//   TMA bulk copies (the caller's SoA arrays, 128-row tiles, LSLB_X2_STAGES = 6-stage ring, mbarrier
//   complete_tx) -> all warps run the row loop on the tile with shared-memory broadcasts. The warps
//   are NOT synchronised per tile: a shared counter per stage lets the last warp that leaves a stage
//   refill it, so warps drift up to 6 tiles apart and nobody waits at a barrier. A row's weights are
//   scalars shared by both chains of a thread: FFMA2 takes them as a 32-bit operand broadcast to
//   both halves (SASS `Rn.F32`), so nothing is duplicated in shared memory and the register file
//   reads one register instead of a pair.
//
// Register-file bandwidth, not the FMA pipe, is the first limit of this loop: an FFMA2 with three
// distinct register-PAIR operands needs 3 RF cycles, one with a broadcast scalar or a reused
// operand 2 (tools/microbench.py: 3-register scalar FFMA reaches 45 of 72 TFLOP/s). Rows are
// therefore processed four at a time, ordered so that the coefficient operand (forward) and the
// residual operand (backward) repeat in consecutive instructions and the operand-reuse cache
// serves them.
#include <type_traits>

#include "common.cuh"
#include "x2_common.cuh"

namespace lslb {

#ifdef LSLB_X2_TRACE
// Investigation build only (NVCC_EXTRA=-DLSLB_X2_TRACE): per-CTA start/end times and SM ids.
__device__ unsigned long long g_x2_trace[8192][6];
__device__ __forceinline__ unsigned long long x2_gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned x2_smid() {
  unsigned s;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
  return s;
}
#endif

// PM: bit m set <=> smooth m feeds predictor 1 (log scale)
// UNITY: every row's four weights sum to one (all x inside the knot range; checked at plan creation).
// Then, over a run of equal spans, eta = b1 + sum_{u != 1} w_u (b_u - b1) and the gradient of b1 is
// (sum of residuals) - g0 - g2 - g3: three FMAs per smooth and direction instead of four.
// X2S: depth of the TMA ring (tiles the fastest warp may run ahead of the slowest). 6 for the 4..16-warp
// CTAs; the 2- and 1-warp CTAs of small chain counts take 3 and 2 stages so that 8 / 16 of them fit an
// SM's shared memory (with 6 stages = 34 KB each only 6 did: 6 resident warps at 64 chains).
// SHW: both smooths are splines of the SAME covariate on the same knots (H / S3: mean and log-sd smooths of
// one x; checked bit for bit at plan creation): one copy of the weights and spans is staged and held in
// registers for both (32 registers and half of the shared-memory loads of the hot loop).
// PHASED (at most 8 chain warps per chain tile, i.e. up to 512 chains): the SM holds ONE CTA of ~16 warps, split
// into nph = 2 / 4 / 8 / 16 row PHASES of cw chain warps each. Tile k of the CTA's row range is evaluated by
// the warps of phase k mod nph only, and every phase writes its own partial slot. X2S is a multiple of nph,
// so a ring stage always serves the same phase (the parity wait on its mbarrier cannot alias a use by
// another phase) and each phase refills its own stages: in effect nph rings of X2S / nph stages, launched
// and retired together. Replaces the several small CTAs per SM of round 2 (A/B in DESIGN section 4).
template <int NSL, int X2S, bool PHASED>
struct X2Smem {
  XRaw<NSL> raw[X2S + (PHASED ? 1 : 0)];  // PHASED: one more tile buffer for the ragged last tile of the data
  u64 full[X2S];
  int done[X2S];  // warps that have finished with the tile in stage s
  alignas(16) float tws[X2S][NSL][4];  // the tile's weight-column sums (plan constant)
};
extern __shared__ __align__(128) unsigned char x2_dyn_smem[];

template <int FAM, int NS, int PM, bool GRAD, bool UNITY, int X2S, int NSL = NS, bool PHASED = false>  // NSL: bands loaded (1 = shared)
#ifndef LSLB_X2_NRB
#define LSLB_X2_NRB 4   // rows per register block (4 or 2)
#endif
#ifndef LSLB_X2_STAGES
#define LSLB_X2_STAGES 6  // ring depth: how many tiles the fastest warp of a CTA may run ahead of the slowest
#endif
__global__ void __launch_bounds__(512, 1)  // 16 warps per SM in ONE CTA
banded_x2_kernel(const __grid_constant__ EvalParams<float> p, int tiles_per_cta, int ntiles_total, int tiles_extra,
                 int cw /* PHASED: chain warps per phase */, int nph /* PHASED: row phases */) {
  constexpr bool SHW = NSL < NS;
  X2Smem<NSL, X2S, PHASED>* smp;
  if constexpr (PHASED) {  // 16- and 32-stage rings: beyond the 48 KB of static shared memory
    smp = reinterpret_cast<X2Smem<NSL, X2S, PHASED>*>(x2_dyn_smem);
  } else {
    __shared__ __align__(128) X2Smem<NSL, X2S, false> sm_static;
    smp = &sm_static;
  }
  XRaw<NSL>* const raw = smp->raw;
  u64* const full = smp->full;
  int* const done = smp->done;
  float (*const tws)[NSL][4] = smp->tws;

  const int tid = threadIdx.x;
  // c: first of this thread's two chains (< Cpad); ch: partial slot; cy: the CTA's row chunk; k0: first tile
  // (cw and nph are read from the parameter bank where they are used and the phase index is recomputed in the
  //  rare ragged-tile branch: the hot instantiations sit at the 128-register cap)
  const int cy = blockIdx.y;
  int c, ch, k0 = 0;
  if constexpr (PHASED) {
    const int warp = tid >> 5;
    const int ph = warp / cw;
    c = 2 * ((blockIdx.x * cw + (warp - ph * cw)) * 32 + (tid & 31));
    ch = cy * nph + ph;
    k0 = ph;
  } else {
    c = 2 * (blockIdx.x * blockDim.x + tid);
    ch = cy;
  }
  (void)k0; (void)nph; (void)cw;
  // balanced partition of the tiles over the row chunks (chunk sizes differ by at most one tile; with
  // ceil(T / CTAs) tiles per CTA the last SMs of a 4-CTAs-per-SM launch held 3 CTAs and the others
  // carried 56 tiles instead of 52.8)
  // tiles_per_cta = floor(T / chunks); the first `tiles_extra` chunks take one tile more. (Plain int
  // arithmetic on purpose: the hot instantiation sits at 128 registers and a 64-bit division here made
  // it spill - 5 % on H.)
  // (written as ONE multiply of the block index by a selected constant: the hot instantiation sits at
  //  128 registers and spilled - 5 % on H - with min(ch, extra) or a 64-bit division here)
  const int tpa = cy < tiles_extra ? tiles_per_cta + 1 : tiles_per_cta;
  const int tile0 = cy * tpa + (cy < tiles_extra ? 0 : tiles_extra);
  const int my_tiles = tpa;
  (void)ntiles_total;
#ifdef LSLB_X2_TRACE
  unsigned long long tr_t0 = x2_gtime();
  int tr_nonuni = 0;
#endif

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < X2S; ++s) {
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
    const long long r0 = (long long)t * XR;
    constexpr unsigned bytes = XR * 4 + NSL * (XR * 16 + XR * 4 + 16);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(x_smem_u32(&full[s])),
                 "r"(bytes)
                 : "memory");
    bulk(raw[s].y, p.y + r0, XR * 4, &full[s]);
#pragma unroll
    for (int m = 0; m < NSL; ++m) {
      bulk(raw[s].w[m], p.sp[m].w + 4 * r0, XR * 16, &full[s]);
      bulk(raw[s].start[m], p.sp[m].start + r0, XR * 4, &full[s]);
      bulk(tws[s][m], p.sp[m].tw + 4 * (long long)t, 16, &full[s]);
    }
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
    for (int k = 0; k < X2S && k < my_tiles; ++k)
      if (tile_rows(tile0 + k) == XR) issue(tile0 + k, k);
  }

  // ---- per-thread state: two chains ------------------------------------------------------------
  const size_t ldc = (size_t)p.Cpad;
  auto ld2 = [&](int j) -> u64 { return *reinterpret_cast<const u64*>(p.pT + (size_t)j * ldc + c); };
  float* mine = p.partial + (size_t)ch * (p.Psmall + LSLB_NACC) * ldc + c;
  auto at2 = [&](int j) -> u64* { return reinterpret_cast<u64*>(mine + (size_t)j * ldc); };

  u64 o0 = pk(0.f, 0.f), o1 = pk(0.f, 0.f);
  for (int q = 0; q < p.ni; ++q) {
    const u64 v = ld2(p.ic[q].soff);
    if (p.ic[q].pred) o1 = add2(o1, v); else o0 = add2(o0, v);
  }
  // Normal location-scale: predictor 1 is carried as -log2(e) * eta1 (its coefficients are scaled
  // once when a run starts), so exp(-eta1) is a bare ex2 with no per-row multiply
  if constexpr (FAM == FAM_NORMAL_LS) o1 = mul2(o1, pk(-1.4426950408889634f, -1.4426950408889634f));
  // Poisson: predictor 0 is carried as log2(e) * eta0, so the mean exp(eta0) is a bare ex2
  if constexpr (FAM == FAM_POISSON) o0 = mul2(o0, pk(1.4426950408889634f, 1.4426950408889634f));
  if constexpr (GRAD) {
    for (int j = 0; j < p.Psmall; ++j) *at2(j) = pk(0.f, 0.f);
  }
  int cur[NS];
  u64 wb[NS][4], wg[NS][4];
#pragma unroll
  for (int m = 0; m < NS; ++m) {
    cur[m] = -1;
#pragma unroll
    for (int t = 0; t < 4; ++t) { wb[m][t] = pk(0.f, 0.f); wg[m][t] = pk(0.f, 0.f); }
  }
  const u64 NEG1 = pk(-1.f, -1.f);
  const u64 NLOG2E = pk(-1.4426950408889634f, -1.4426950408889634f);
  const u64 ONE = pk(1.f, 1.f);
  (void)ONE;
  const u64 INVS_FIXED = pk(p.fixed_inv_scale, p.fixed_inv_scale);
  float lpa_hi = 0.f, lpa_lo = 0.f, lpb_hi = 0.f, lpb_lo = 0.f;  // Kahan pairs, chains a and b
  u64 sr0 = pk(0.f, 0.f), sr1 = pk(0.f, 0.f);

  bool fastmode = false;            // accumulators currently hold the 3-term (UNITY) form
  u64 srs[NS];                      // residual-sum snapshot at the start of the current run
  u64 e0i = o0, e1i = o1;           // predictor start values in 3-term form (offset + sum of b1)
#pragma unroll
  for (int m = 0; m < NS; ++m) srs[m] = pk(0.f, 0.f);

  auto flush = [&](int m) {
    if constexpr (GRAD) {
      if (cur[m] >= 0) {
        const int so = p.sp[m].soff + cur[m];
        if (fastmode) {
          const u64 S = add2(((PM >> m) & 1) ? sr1 : sr0, mul2(srs[m], NEG1));
          const u64 g1 = add2(add2(S, mul2(wg[m][0], NEG1)), mul2(add2(wg[m][2], wg[m][3]), NEG1));
          wg[m][1] = g1;
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          u64* q = at2(so + t);
          *q = add2(*q, wg[m][t]);
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) wg[m][t] = pk(0.f, 0.f);
  };
  auto load = [&](int m, int s, bool fast) {
    const int so = p.sp[m].soff;
#pragma unroll
    for (int t = 0; t < 4; ++t) wb[m][t] = ld2(so + s + t);
    if constexpr (FAM == FAM_NORMAL_LS) {
      if ((PM >> m) & 1) {
#pragma unroll
        for (int t = 0; t < 4; ++t) wb[m][t] = mul2(wb[m][t], NLOG2E);
      }
    }
    if constexpr (FAM == FAM_POISSON) {
#pragma unroll
      for (int t = 0; t < 4; ++t) wb[m][t] = mul2(wb[m][t], mul2(NLOG2E, NEG1));
    }
    if (fast) {
      const u64 nb1 = mul2(wb[m][1], NEG1);
      wb[m][0] = add2(wb[m][0], nb1);
      wb[m][2] = add2(wb[m][2], nb1);
      wb[m][3] = add2(wb[m][3], nb1);
    }
    srs[m] = ((PM >> m) & 1) ? sr1 : sr0;
    cur[m] = s;
  };
  auto change = [&](int m, int s) {  // 4-term form (rows of a tile with mixed spans)
    flush(m);
    load(m, s, false);
  };

  for (int k = k0; k < my_tiles; k += PHASED ? nph : 1) {
    const int t = tile0 + k;
    const int s = k % X2S;
    const int rows = tile_rows(t);
    // PHASED: the ragged tile (the last tile of the data, one per launch and chain tile) has its own buffer
    XRaw<NSL>& D = raw[(PHASED && rows != XR) ? X2S : s];
    if (rows == XR) {
      wait_full(s, (k / X2S) & 1);
    } else {  // ragged last tile of the data: ordinary loads (bulk copies need 16-byte multiples)
      // not PHASED: every warp has left the tile that used this stage before
      if constexpr (!PHASED) __syncthreads();
      const long long r0 = (long long)t * XR;
      const int ph = PHASED ? (tid >> 5) / cw : 0;
      const int tq = PHASED ? tid - ph * cw * 32 : tid, nq = PHASED ? cw * 32 : (int)blockDim.x;
      for (int r = tq; r < rows; r += nq) {
        D.y[r] = p.y[r0 + r];
#pragma unroll
        for (int m = 0; m < NSL; ++m) {
          D.start[m][r] = p.sp[m].start[r0 + r];
#pragma unroll
          for (int u = 0; u < 4; ++u) D.w[m][4 * r + u] = p.sp[m].w[4 * (r0 + r) + u];
        }
      }
      if constexpr (PHASED) {  // the warps of this phase only (named barrier 1 + ph; one warp: __syncwarp)
        if (cw > 1) asm volatile("bar.sync %0, %1;" ::"r"(1 + ph), "r"(cw * 32) : "memory");
        else __syncwarp();
      } else {
        __syncthreads();
      }
    }

    // is every smooth's span constant over this tile? (each warp evaluates it redundantly)
    // full tiles: the plan stored the answer (the tile's span, or -1) with the tile's weight sums
    bool uni = true;
    int span0[NSL];
    if (rows == XR) {
#pragma unroll
      for (int m = 0; m < NSL; ++m) {
        span0[m] = __float_as_int(tws[s][m][1]);
        uni = uni && span0[m] >= 0;
      }
    } else {
#pragma unroll
      for (int m = 0; m < NSL; ++m) {
        span0[m] = D.start[m][0];
        bool eq = true;
        for (int r = (tid & 31); r < rows; r += 32) eq = eq && (D.start[m][r] == span0[m]);
        uni = uni && __all_sync(0xffffffffu, eq);
      }
    }

    u64 sz2 = pk(0.f, 0.f), se1 = pk(0.f, 0.f);  // per-tile sums: z^2 and eta1 (Normal); see the other families
    u64 sl2 = pk(0.f, 0.f);
    (void)sl2;

    // NR rows whose weights/responses are already in registers: forward with wb fixed across rows
    // (u-major), likelihood, backward with the residual fixed across the four weights of a row
    auto compute_block = [&](const float (&w)[NSL][4][4], const float (&yv)[4], auto NRtag, auto FASTtag) {
      constexpr int NR = decltype(NRtag)::value;
      constexpr bool FAST = decltype(FASTtag)::value;
      u64 eta0[NR], eta1[NR];
#pragma unroll
      for (int i = 0; i < NR; ++i) { eta0[i] = FAST ? e0i : o0; eta1[i] = FAST ? e1i : o1; }
#pragma unroll
      for (int m = 0; m < NS; ++m) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (FAST && u == 1) continue;
#pragma unroll
          for (int i = 0; i < NR; ++i) {
            if ((PM >> m) & 1) eta1[i] = fma2(bc(w[(SHW ? 0 : m)][i][u]), wb[m][u], eta1[i]);
            else eta0[i] = fma2(bc(w[(SHW ? 0 : m)][i][u]), wb[m][u], eta0[i]);
          }
        }
      }
      u64 d0[NR], d1[NR];
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        if constexpr (FAM == FAM_POISSON) {
          // eta0 holds log2(e) * eta: mu = ex2(.), ll = y eta - mu (sz2 = sum y eta', se1 = sum mu)
          float ta, tb;
          upk(eta0[i], ta, tb);
          const u64 mu = pk(ex2f(ta), ex2f(tb));
          const u64 yy = bc(yv[i]);
          sz2 = fma2(yy, eta0[i], sz2);
          se1 = add2(se1, mu);
          if constexpr (GRAD) {
            d0[i] = fma2(mu, NEG1, yy);
            sr0 = add2(sr0, d0[i]);
            d1[i] = d0[i];
          }
        } else if constexpr (FAM == FAM_BERNOULLI) {
          // a = |eta|, e = exp(-a), s = 1 + e:  pi - 1/2 = copysign(1/s - 1/2, eta),
          // ll = (y - 1/2) eta - a/2 - log(s)   (sz2 = sum (y - 1/2) eta, se1 = sum a, sl2 = sum log2 s)
          const u64 a = eta0[i] & 0x7FFFFFFF7FFFFFFFull;
          float ta, tb;
          upk(mul2(a, NLOG2E), ta, tb);
          const u64 s1 = add2(pk(ex2f(ta), ex2f(tb)), ONE);
          float sa, sb;
          upk(s1, sa, sb);
          const u64 h = add2(pk(rcpf(sa), rcpf(sb)), pk(-0.5f, -0.5f));
          sl2 = add2(sl2, pk(lg2f(sa), lg2f(sb)));
          const u64 ym = bc(yv[i] - 0.5f);
          sz2 = fma2(ym, eta0[i], sz2);
          se1 = add2(se1, a);
          if constexpr (GRAD) {
            const u64 hs = h ^ (eta0[i] & 0x8000000080000000ull);
            d0[i] = fma2(hs, NEG1, ym);  // y - pi
            sr0 = add2(sr0, d0[i]);
            d1[i] = d0[i];
          }
        } else {
        const u64 d = fma2(eta0[i], NEG1, bc(yv[i]));  // y - eta0
        u64 inv_s;
        if constexpr (FAM == FAM_NORMAL_LS) {
          float ta, tb;
          upk(eta1[i], ta, tb);  // already -log2(e) * eta1
          inv_s = pk(ex2f(ta), ex2f(tb));  // exp(-eta1)
          if constexpr (!FAST) se1 = add2(se1, eta1[i]);  // FAST: closed form per tile, below
        } else {
          inv_s = INVS_FIXED;
        }
        const u64 z = mul2(d, inv_s);
        sz2 = fma2(z, z, sz2);
        if constexpr (GRAD) {
          d0[i] = mul2(z, inv_s);
          sr0 = add2(sr0, d0[i]);
          d1[i] = d0[i];
          if constexpr (PM != 0) d1[i] = fma2(z, z, NEG1);
        }
      }  // Normal families
      }
      if constexpr (GRAD) {
#pragma unroll
        for (int m = 0; m < NS; ++m) {
#pragma unroll
          for (int i = 0; i < NR; ++i) {
            const u64 rr = ((PM >> m) & 1) ? d1[i] : d0[i];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (FAST && u == 1) continue;
              wg[m][u] = fma2(bc(w[(SHW ? 0 : m)][i][u]), rr, wg[m][u]);
            }
          }
        }
      }
    };
    using TFast = std::integral_constant<bool, true>;
    using TFull = std::integral_constant<bool, false>;
    auto load_block = [&](float (&w)[NSL][4][4], float (&yv)[4], int r, auto NRtag) {
      constexpr int NR = decltype(NRtag)::value;
#pragma unroll
      for (int m = 0; m < NSL; ++m)
#pragma unroll
        for (int i = 0; i < NR; ++i) {
          const float4 v = *reinterpret_cast<const float4*>(&D.w[m][4 * (r + i)]);
          w[m][i][0] = v.x; w[m][i][1] = v.y; w[m][i][2] = v.z; w[m][i][3] = v.w;
        }
      if constexpr (NR == 4) {
        const float4 v = *reinterpret_cast<const float4*>(&D.y[r]);
        yv[0] = v.x; yv[1] = v.y; yv[2] = v.z; yv[3] = v.w;
      } else if constexpr (NR == 2) {
        const float2 v = *reinterpret_cast<const float2*>(&D.y[r]);
        yv[0] = v.x; yv[1] = v.y;
      } else {
        yv[0] = D.y[r];
      }
    };
    using N1 = std::integral_constant<int, 1>;
    using N4 = std::integral_constant<int, LSLB_X2_NRB>;
    constexpr int NRB = LSLB_X2_NRB;

#ifdef LSLB_X2_TRACE
    tr_nonuni += uni ? 0 : 1;
#endif
    const bool want_fast = UNITY && uni && rows == XR;
    if (want_fast != fastmode) {  // the accumulators change form: write them out in the old form
#pragma unroll
      for (int m = 0; m < NS; ++m) {
        flush(m);
        cur[m] = -1;
      }
      fastmode = want_fast;
    }
    if (uni) {
      bool reloaded = false;
#pragma unroll
      for (int m = 0; m < NS; ++m) {
        const int first = span0[(SHW ? 0 : m)];
        if (first != cur[m]) {
          flush(m);
          load(m, first, fastmode);
          reloaded = true;
        }
      }
      if (fastmode && reloaded) {
        e0i = o0;
        e1i = o1;
#pragma unroll
        for (int m = 0; m < NS; ++m) {
          if ((PM >> m) & 1) e1i = add2(e1i, wb[m][1]);
          else e0i = add2(e0i, wb[m][1]);
        }
      }
      if (rows == XR) {
        // software pipeline over 4-row blocks: the next block's weights are loaded into a second
        // register set while the current block computes
        float wa[NSL][4][4], wbuf[NSL][4][4], ya[4], yb[4];
        load_block(wa, ya, 0, N4{});
        if (fastmode) {
#pragma unroll 1
          for (int r = 0; r < XR; r += 2 * NRB) {
            load_block(wbuf, yb, r + NRB, N4{});
            compute_block(wa, ya, N4{}, TFast{});
            if (r + 2 * NRB < XR) load_block(wa, ya, r + 2 * NRB, N4{});
            compute_block(wbuf, yb, N4{}, TFast{});
          }
          if constexpr (FAM == FAM_NORMAL_LS) {
            // sum over the tile's rows of predictor 1 in closed form: every row is
            // e1i + sum_{u != 1} w_u (b_u - b_1), and the column sums of w are a plan constant
            se1 = mul2(e1i, pk((float)XR, (float)XR));
#pragma unroll
            for (int m = 0; m < NS; ++m) {
              if ((PM >> m) & 1) {
                const float4 tw = *reinterpret_cast<const float4*>(tws[s][(SHW ? 0 : m)]);
                se1 = fma2(bc(tw.x), wb[m][0], se1);
                se1 = fma2(bc(tw.z), wb[m][2], se1);
                se1 = fma2(bc(tw.w), wb[m][3], se1);
              }
            }
          }
        } else {
#pragma unroll 1
          for (int r = 0; r < XR; r += 2 * NRB) {
            load_block(wbuf, yb, r + NRB, N4{});
            compute_block(wa, ya, N4{}, TFull{});
            if (r + 2 * NRB < XR) load_block(wa, ya, r + 2 * NRB, N4{});
            compute_block(wbuf, yb, N4{}, TFull{});
          }
        }
      } else {
        float w1[NSL][4][4], y1[4];
        for (int r = 0; r < rows; ++r) {
          load_block(w1, y1, r, N1{});
          compute_block(w1, y1, N1{}, TFull{});
        }
      }
    } else {
      float w1[NSL][4][4], y1[4];
      for (int r = 0; r < rows; ++r) {
#pragma unroll
        for (int m = 0; m < NS; ++m) {
          const int sm = D.start[(SHW ? 0 : m)][r];
          if (sm != cur[m]) change(m, sm);
        }
        load_block(w1, y1, r, N1{});
        compute_block(w1, y1, N1{}, TFull{});
      }
    }
    // tile sums -> log-likelihood (row constants are added in finalize) and d/d eta1 sums
    float z2a, z2b, e1a, e1b;
    upk(sz2, z2a, z2b);
    upk(se1, e1a, e1b);
    if constexpr (FAM == FAM_NORMAL_LS) {  // se1 holds -log2(e) * sum eta1
      e1a *= -0.6931471805599453f;
      e1b *= -0.6931471805599453f;
    }
    if constexpr (FAM == FAM_POISSON) {  // ln2 * sum(y eta') - sum mu
      kahan_add(lpa_hi, lpa_lo, fmaf(0.6931471805599453f, z2a, -e1a));
      kahan_add(lpb_hi, lpb_lo, fmaf(0.6931471805599453f, z2b, -e1b));
    } else if constexpr (FAM == FAM_BERNOULLI) {  // sum (y - 1/2) eta - sum a / 2 - ln2 * sum log2 s
      float l2a, l2b;
      upk(sl2, l2a, l2b);
      kahan_add(lpa_hi, lpa_lo, fmaf(-0.6931471805599453f, l2a, fmaf(-0.5f, e1a, z2a)));
      kahan_add(lpb_hi, lpb_lo, fmaf(-0.6931471805599453f, l2b, fmaf(-0.5f, e1b, z2b)));
    } else {
      kahan_add(lpa_hi, lpa_lo, fmaf(-0.5f, z2a, -e1a));
      kahan_add(lpb_hi, lpb_lo, fmaf(-0.5f, z2b, -e1b));
    }
    if constexpr (GRAD && FAM == FAM_NORMAL_LS) {
      const float nr = -(float)rows;
      sr1 = add2(sr1, add2(sz2, pk(nr, nr)));  // sum (z^2 - 1)
    }
    // The warps of a CTA are NOT synchronised per tile: whichever warp is the last to finish with
    // stage s refills it. (With a barrier here — or with several small CTAs per SM — the hardware's
    // oldest-first warp scheduling lets some warps finish the whole chunk at 45 % of the kernel's
    // duration while the rest of the SM runs under-occupied: measured CTA durations 326..755 us.)
    __syncwarp();
    if ((tid & 31) == 0) {
      __threadfence_block();
      if (atomicAdd(&done[s], 1) == (PHASED ? cw : (int)(blockDim.x >> 5)) - 1) {
        done[s] = 0;
        const int kn = k + X2S;
        if (kn < my_tiles && tile_rows(tile0 + kn) == XR) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          issue(tile0 + kn, s);
        }
      }
    }
  }

#ifdef LSLB_X2_TRACE
  if (tid == 0) {
    const int id = blockIdx.y * gridDim.x + blockIdx.x;
    if (id < 8192) {
      g_x2_trace[id][0] = x2_smid();
      g_x2_trace[id][1] = tr_t0;
      g_x2_trace[id][2] = x2_gtime();
      g_x2_trace[id][3] = blockIdx.x;
      g_x2_trace[id][4] = blockIdx.y;
      g_x2_trace[id][5] = ((unsigned long long)my_tiles << 32) | (unsigned)tr_nonuni;
    }
  }
#endif
  if constexpr (GRAD) {
#pragma unroll
    for (int m = 0; m < NS; ++m) flush(m);
    for (int q = 0; q < p.ni; ++q) *at2(p.ic[q].soff) = p.ic[q].pred ? sr1 : sr0;
  }
  *at2(p.Psmall) = pk(lpa_hi, lpb_hi);
  *at2(p.Psmall + 1) = pk(lpa_lo, lpb_lo);
}

// host ------------------------------------------------------------------------------------------
template <int FAM, int NS, int PM>
static int x2_launch(const EvalParams<float>& ep, const BandedCfg& k, bool g, cudaStream_t st, bool unity, bool shw) {
  // Not phased: a CTA owns tc_eff chains = tc_eff / 2 threads (whole warps). Phased: cw = tc_eff / 64 chain
  // warps x k.phases row phases, one partial slot per phase (k.nchunks = CTAs x phases).
  const int cw = k.tc_eff / 64;
  dim3 grid(k.ntiles_chain, k.nchunks / k.phases), block(k.tc_eff / 2 * k.phases);
  // the shared-band variant pays off for CTAs of up to 8 chain warps (A/B on one box, H model: 64 chains 0.095 ->
  // 0.089 ms, 256: 0.230 -> 0.224, 512: 0.380 -> 0.378) and LOSES 1.2 % with the 16-warp CTA of 1024 chains
  // (0.689 -> 0.697: ptxas schedules the 120-register variant differently), so that one keeps two bands
  shw = shw && k.tc_eff <= 512;
#define LSLB_X2_ARGS ep, k.tiles_per_cta, k.ntiles_total, k.tiles_extra, cw, k.phases
#define LSLB_X2_GO3(GR, UN, ST, SH, PHS)                                                   \
  do {                                                                                     \
    auto kern = banded_x2_kernel<FAM, NS, PM, GR, UN, ST, SH, PHS>;                        \
    size_t dyn = 0;                                                                        \
    if (PHS) {                                                                             \
      dyn = sizeof(X2Smem<SH, ST, PHS>);                                                   \
      LSLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn)); \
    }                                                                                      \
    kern<<<grid, block, dyn, st>>>(LSLB_X2_ARGS);                                          \
  } while (0)
#define LSLB_X2_GO2(ST, SH, PHS)                                    \
  do {                                                              \
    if (unity) {                                                    \
      if (g) LSLB_X2_GO3(true, true, ST, SH, PHS);                  \
      else LSLB_X2_GO3(false, true, ST, SH, PHS);                   \
    } else {                                                        \
      if (g) LSLB_X2_GO3(true, false, ST, SH, PHS);                 \
      else LSLB_X2_GO3(false, false, ST, SH, PHS);                  \
    }                                                               \
  } while (0)
#define LSLB_X2_GO(ST)                                  \
  do {                                                  \
    if constexpr (NS == 2) {                            \
      if (shw) LSLB_X2_GO2(ST, 1, false);               \
      else LSLB_X2_GO2(ST, NS, false);                  \
    } else {                                            \
      LSLB_X2_GO2(ST, NS, false);                       \
    }                                                   \
  } while (0)
  if (k.phases > 1) {
    // phased CTAs: shared-band plans only (banded_cfg); ring of 16 stages = a multiple of 2, 4 and 8 phases
    if constexpr (NS == 2) LSLB_X2_GO2(16, 1, true);
    else return LSLB_ERR_INVALID;
  } else if (k.stages >= LSLB_X2_STAGES) LSLB_X2_GO(LSLB_X2_STAGES);
  else if (k.stages == 3) LSLB_X2_GO(3);
  else LSLB_X2_GO(2);
#undef LSLB_X2_GO
#undef LSLB_X2_GO2
#undef LSLB_X2_GO3
#undef LSLB_X2_ARGS
  LSLB_LAUNCH_CHECK();
  return LSLB_OK;
}

int launch_banded_x2(const lslb_plan* plan, const EvalParams<float>& ep, const BandedCfg& k, bool g,
                     cudaStream_t st) {
  int pm = 0;
  for (int m = 0; m < plan->ns; ++m) pm |= (plan->blk[plan->sp[m]].predictor & 1) << m;
#ifdef LSLB_X2_DEV  // development builds: the H / S3 instantiations only (the full file takes minutes to compile)
  return x2_launch<FAM_NORMAL_LS, 2, 2>(ep, k, g, st, plan->unit_rows, plan->shared_band);
#else
  if (plan->fam == FAM_NORMAL_LS) {
    if (plan->ns == 1) return pm ? x2_launch<FAM_NORMAL_LS, 1, 1>(ep, k, g, st, plan->unit_rows, plan->shared_band)
                                 : x2_launch<FAM_NORMAL_LS, 1, 0>(ep, k, g, st, plan->unit_rows, plan->shared_band);
    switch (pm) {
      case 0: return x2_launch<FAM_NORMAL_LS, 2, 0>(ep, k, g, st, plan->unit_rows, plan->shared_band);
      case 1: return x2_launch<FAM_NORMAL_LS, 2, 1>(ep, k, g, st, plan->unit_rows, plan->shared_band);
      case 2: return x2_launch<FAM_NORMAL_LS, 2, 2>(ep, k, g, st, plan->unit_rows, plan->shared_band);
      default: return x2_launch<FAM_NORMAL_LS, 2, 3>(ep, k, g, st, plan->unit_rows, plan->shared_band);
    }
  }
  if (plan->fam == FAM_POISSON) {
    if (plan->ns == 1) return x2_launch<FAM_POISSON, 1, 0>(ep, k, g, st, plan->unit_rows, plan->shared_band);
    return x2_launch<FAM_POISSON, 2, 0>(ep, k, g, st, plan->unit_rows, plan->shared_band);
  }
  if (plan->fam == FAM_BERNOULLI) {
    if (plan->ns == 1) return x2_launch<FAM_BERNOULLI, 1, 0>(ep, k, g, st, plan->unit_rows, plan->shared_band);
    return x2_launch<FAM_BERNOULLI, 2, 0>(ep, k, g, st, plan->unit_rows, plan->shared_band);
  }
  if (plan->ns == 1) return x2_launch<FAM_NORMAL_FIXED, 1, 0>(ep, k, g, st, plan->unit_rows, plan->shared_band);
  return x2_launch<FAM_NORMAL_FIXED, 2, 0>(ep, k, g, st, plan->unit_rows, plan->shared_band);
#endif
}

}  // namespace lslb

#ifdef LSLB_X2_TRACE
extern "C" int lslb_debug_x2_trace(unsigned long long* out_host, int rows) {
  if (rows > 8192) rows = 8192;
  return cudaMemcpyFromSymbol(out_host, lslb::g_x2_trace, (size_t)rows * 6 * sizeof(unsigned long long)) == cudaSuccess ? 0 : 3;
}
#endif
