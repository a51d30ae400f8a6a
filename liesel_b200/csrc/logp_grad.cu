// Log-posterior + gradient for C chains against ONE copy of the data.
//
// Replaces, per chain, one `Model.update_state` walk + `_model_log_prob` read + reverse-mode
// gradient of the reference (src/liesel/model/model.py:2786-2865, src/liesel/goose/kernel.py:149-158,
// src/liesel/goose/nuts.py:193-194). Layout of the computation:
//
//   gather   params [C x P]  ->  pT [Psmall x Cpad]   (+ bT [G x Cpad] for a random intercept)
//   eval     grid (row chunks, chain tiles); thread = chain; rows streamed once per chain tile;
//            per-chunk partial sums -> partial [nchunks x (Psmall+2) x Cpad]
//   finalize fixed-order sum over chunks (double), priors added, outputs in [C x P] / [C]
//
// No atomics anywhere: results are bitwise reproducible run to run.
#include "common.cuh"
#include "prologue.cuh"

namespace lslb {

// gbT [G x Cpad] -> grad[:, goff:goff+G]
template <typename T>
__global__ void group_scatter_kernel(const T* __restrict__ gbT, int P, int goff, int G, int C,
                                     int Cpad, T* __restrict__ grad) {
  __shared__ T tile[32][33];
  int g0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int gl = ty + 8 * k;
    int g = g0 + gl, c = c0 + tx;
    tile[gl][tx] = (g < G && c < Cpad) ? gbT[(long long)g * Cpad + c] : T(0);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    int cl = ty + 8 * k;
    int g = g0 + tx, c = c0 + cl;
    if (g < G && c < C) grad[(long long)c * P + goff + g] = tile[tx][cl];
  }
}

// =======================================================================================
// the evaluation kernel
// =======================================================================================
// NS  : compile-time bound on the number of spline blocks (p.ns <= NS)
// PD  : 0 = no dense block; >0 = ONE dense block with <= PD columns held in registers;
//       -1 = any number of dense blocks, coefficients and gradients in shared memory
template <typename T, int FAM, int NS, int PD, bool GROUP, bool GRAD>
__global__ void __launch_bounds__(128)
eval_kernel(const __grid_constant__ EvalParams<T> p, int fstride) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int TC = blockDim.x, tid = threadIdx.x;
  const int c = blockIdx.y * TC + tid;
  // Spline-only models (STAGED) keep only the gradient columns in shared memory and read the
  // coefficients from the transposed copy in global memory (L1-resident): half the shared memory
  // per chain, twice the resident warps for these latency-bound shapes.
  constexpr bool STAGED = NS > 0 && PD == 0 && !GROUP;
  T* bs = reinterpret_cast<T*>(smraw);
  T* gs = STAGED ? bs : bs + (size_t)p.Psmall * TC;
  auto beta = [&](int j) -> T {
    if constexpr (STAGED) return __ldg(p.pT + (size_t)j * p.Cpad + c);
    else return bs[j * TC + tid];
  };

  for (int j = 0; j < p.Psmall; ++j) {
    if constexpr (!STAGED) bs[j * TC + tid] = p.pT[(size_t)j * p.Cpad + c];
    if constexpr (GRAD) gs[j * TC + tid] = T(0);
  }
  // every thread reads and writes only its own column of bs/gs: no barrier needed

  int f0 = blockIdx.x * fstride;
  int f1 = min(f0 + fstride, p.nchunks);
  const long long r0 = p.chunk_start[f0], r1 = p.chunk_start[f1];

  T o0 = T(0), o1 = T(0);
  for (int q = 0; q < p.ni; ++q) {
    T v = beta(p.ic[q].soff);
    if (p.ic[q].pred) o1 += v; else o0 += v;
  }

  int cur[NS > 0 ? NS : 1];
  T wb[NS > 0 ? NS : 1][4], wg[NS > 0 ? NS : 1][4];
#pragma unroll
  for (int m = 0; m < NS; ++m) {
    cur[m] = -1;
#pragma unroll
    for (int t = 0; t < 4; ++t) { wb[m][t] = T(0); wg[m][t] = T(0); }
  }

  constexpr int PDR = PD > 0 ? PD : 1;
  T bd[PDR], gd[PDR];
  if constexpr (PD > 0) {
#pragma unroll
    for (int j = 0; j < PD; ++j) {
      bd[j] = j < p.dn[0].p ? beta(p.dn[0].soff + j) : T(0);
      gd[j] = T(0);
    }
  }

  int curg = -1;
  T bg = T(0), accg = T(0);

  T lp_hi = T(0), lp_lo = T(0), sr0 = T(0), sr1 = T(0);

  // Spline-only models: the 32-row blocks of (y, start, weights) are staged in shared memory with
  // cp.async (two stages), so the row loop never waits on global memory. Works for any alignment.
  constexpr int NSR = NS > 0 ? NS : 1;
  struct RowTile {
    T w[NSR][32 * 4];
    T y[32];
    int start[NSR][32];
  };
  RowTile* tiles = reinterpret_cast<RowTile*>(
      smraw + (((size_t)(STAGED ? 1 : 2) * p.Psmall * TC * sizeof(T)) + 15) / 16 * 16);
  auto cp_elem = [](void* dst, const void* src, int bytes) {
    if (bytes == 4)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
    else
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
  };
  auto prefetch = [&](long long t0, int stage) {
    if constexpr (STAGED) {
      if (t0 < r1) {
        const int rows = (int)(r1 - t0 < 32 ? r1 - t0 : 32);
        RowTile& tl = tiles[stage];
        for (int e = tid; e < rows; e += TC) cp_elem(&tl.y[e], p.y + t0 + e, sizeof(T));
        for (int m = 0; m < p.ns; ++m) {
          for (int e = tid; e < rows; e += TC) cp_elem(&tl.start[m][e], p.sp[m].start + t0 + e, 4);
          for (int e = tid; e < rows * 4; e += TC) cp_elem(&tl.w[m][e], p.sp[m].w + 4 * t0 + e, sizeof(T));
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
  };
  prefetch(r0, 0);

  int it = 0;
  for (long long t0 = r0; t0 < r1; t0 += 32, ++it) {
    const long long t1 = t0 + 32 < r1 ? t0 + 32 : r1;
    if constexpr (STAGED) {
      prefetch(t0 + 32, (it + 1) & 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
      __syncthreads();
    }
    const RowTile& tl = tiles[it & 1];
    T lp_tile = T(0);
    for (long long i = t0; i < t1; ++i) {
      T eta0 = o0, eta1 = o1;
      Vec4<T> w[NS > 0 ? NS : 1];
#pragma unroll
      for (int m = 0; m < NS; ++m) {
        if (m < p.ns) {
          int s;
          if constexpr (STAGED) {
            s = tl.start[m][i - t0];
            w[m] = *reinterpret_cast<const Vec4<T>*>(&tl.w[m][4 * (i - t0)]);
          } else {
            s = __ldg(p.sp[m].start + i);
            w[m] = ld4(p.sp[m].w + 4 * i);
          }
          if (s != cur[m]) {  // warp-uniform: every lane looks at the same row
            const int so = p.sp[m].soff;
            if constexpr (GRAD) {
              if (cur[m] >= 0) {
#pragma unroll
                for (int t = 0; t < 4; ++t) gs[(so + cur[m] + t) * TC + tid] += wg[m][t];
              }
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              wb[m][t] = beta(so + s + t);
              wg[m][t] = T(0);
            }
            cur[m] = s;
          }
          T v = w[m].a * wb[m][0];
          v = fma(w[m].b, wb[m][1], v);
          v = fma(w[m].c, wb[m][2], v);
          v = fma(w[m].d, wb[m][3], v);
          if (p.sp[m].pred) eta1 += v; else eta0 += v;
        }
      }
      T xr[PDR];
      if constexpr (PD > 0) {
        const T* xrow = p.dn[0].X + i * p.dn[0].ld;
        T v = T(0);
#pragma unroll
        for (int j = 0; j < PD; ++j) {
          xr[j] = j < p.dn[0].p ? __ldg(xrow + j) : T(0);
          v = fma(xr[j], bd[j], v);
        }
        if (p.dn[0].pred) eta1 += v; else eta0 += v;
      } else if constexpr (PD < 0) {
        for (int d = 0; d < p.nd; ++d) {
          const T* xrow = p.dn[d].X + i * p.dn[d].ld;
          const T* bcol = bs + (size_t)p.dn[d].soff * TC + tid;
          T v = T(0);
          for (int j = 0; j < p.dn[d].p; ++j) v = fma(__ldg(xrow + j), bcol[j * TC], v);
          if (p.dn[d].pred) eta1 += v; else eta0 += v;
        }
      }
      if constexpr (GROUP) {
        const int g = __ldg(p.gidx + i);
        if (g != curg) {
          if constexpr (GRAD) {
            if (curg >= 0) p.gbT[(size_t)curg * p.Cpad + c] += accg;
          }
          curg = g;
          bg = p.bT[(size_t)g * p.Cpad + c];
          accg = T(0);
        }
        if (p.gpred) eta1 += bg; else eta0 += bg;
      }

      T ll, d0, d1, u0, u1;
      T yi;
      if constexpr (STAGED) yi = tl.y[i - t0]; else yi = __ldg(p.y + i);
      lik_row<T, FAM, false>(yi, eta0, eta1, p.fixed_inv_scale, ll, d0, d1, u0, u1);
      lp_tile += ll;

      if constexpr (GRAD) {
        sr0 += d0;
        sr1 += d1;
#pragma unroll
        for (int m = 0; m < NS; ++m) {
          if (m < p.ns) {
            const T r = p.sp[m].pred ? d1 : d0;
            wg[m][0] = fma(w[m].a, r, wg[m][0]);
            wg[m][1] = fma(w[m].b, r, wg[m][1]);
            wg[m][2] = fma(w[m].c, r, wg[m][2]);
            wg[m][3] = fma(w[m].d, r, wg[m][3]);
          }
        }
        if constexpr (PD > 0) {
          const T r = p.dn[0].pred ? d1 : d0;
#pragma unroll
          for (int j = 0; j < PD; ++j) gd[j] = fma(xr[j], r, gd[j]);
        } else if constexpr (PD < 0) {
          for (int d = 0; d < p.nd; ++d) {
            const T r = p.dn[d].pred ? d1 : d0;
            const T* xrow = p.dn[d].X + i * p.dn[d].ld;
            T* gcol = gs + (size_t)p.dn[d].soff * TC + tid;
            for (int j = 0; j < p.dn[d].p; ++j) gcol[j * TC] = fma(__ldg(xrow + j), r, gcol[j * TC]);
          }
        }
        if constexpr (GROUP) accg += p.gpred ? d1 : d0;
      }
    }
    kahan_add(lp_hi, lp_lo, lp_tile);
    if constexpr (STAGED) __syncthreads();  // the stage is refilled by the next iteration's prefetch
  }

  // ---- flush register state and publish this chunk's partial sums ----------------------
  const int ch = blockIdx.x;
  T* out = p.partial + (size_t)ch * (p.Psmall + LSLB_NACC) * p.Cpad + c;
  if constexpr (GRAD) {
#pragma unroll
    for (int m = 0; m < NS; ++m) {
      if (m < p.ns && cur[m] >= 0) {
#pragma unroll
        for (int t = 0; t < 4; ++t) gs[(p.sp[m].soff + cur[m] + t) * TC + tid] += wg[m][t];
      }
    }
    if constexpr (PD > 0) {
#pragma unroll
      for (int j = 0; j < PD; ++j)
        if (j < p.dn[0].p) gs[(p.dn[0].soff + j) * TC + tid] = gd[j];
    }
    for (int q = 0; q < p.ni; ++q) gs[p.ic[q].soff * TC + tid] = p.ic[q].pred ? sr1 : sr0;
    if constexpr (GROUP) {
      if (curg >= 0) p.gbT[(size_t)curg * p.Cpad + c] += accg;
    }
    for (int j = 0; j < p.Psmall; ++j) out[(size_t)j * p.Cpad] = gs[j * TC + tid];
  }
  out[(size_t)p.Psmall * p.Cpad] = lp_hi;
  out[(size_t)(p.Psmall + 1) * p.Cpad] = lp_lo;
}

// =======================================================================================
// finalize
// =======================================================================================
// Both finalize kernels use blocks of (32 chains) x (FIN_G helpers): helper g sums the chunk
// partials ch = g, g + FIN_G, ... in double; the FIN_G partial sums are then combined in a fixed
// order through shared memory, so the result does not depend on scheduling.
// FIN_G is 8 for wide launches and 32 when the grid would otherwise leave most SMs idle (few chains).
template <typename T, int FIN_G>
__device__ __forceinline__ void finalize_grad_body(const FinalizeParams& fp, const T* __restrict__ partial,
                                                   const T* __restrict__ pT, const T* __restrict__ aux,
                                                   T* __restrict__ grad, T* __restrict__ vsum,
                                                   double* qf) {
  // One CTA per (32 chains, coefficient row j). It reduces the row's chunk partials (gradient of the
  // likelihood), adds the prior's gradient and ALSO emits the row's share of the prior's quadratic
  // form, qf[j][c] = -1/2 x_j (K x)_j / tau2 (or -1/2 (x_j - m)^2 / s^2): the log-posterior kernel
  // then only sums Psmall numbers per chain. (Until round 2 the log-posterior slice recomputed K x
  // for every block inside ONE CTA per 32 chains: 25-30 us on the critical path of every call.)
  // grad == NULL (value-only call): the prior part alone.
  __shared__ double red[FIN_G][32];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int j = blockIdx.y;
  const int g = threadIdx.y;
  double acc = 0.0;
  if (grad != nullptr) {
    const size_t stride = (size_t)(fp.Psmall + LSLB_NACC) * fp.Cpad;
    const T* src = partial + (size_t)j * fp.Cpad + c;
    // eight loads in flight per thread, summed in the order of the plain loop
    int ch = g;
    for (; ch + 7 * FIN_G < fp.nchunks; ch += 8 * FIN_G) {
      T v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = src[(size_t)(ch + u * FIN_G) * stride];
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += (double)v[u];
    }
    for (; ch < fp.nchunks; ch += FIN_G) acc += (double)src[(size_t)ch * stride];
    red[g][threadIdx.x] = acc;
    __syncthreads();
  }
  if (g != 0) return;
  acc = 0.0;
  if (grad != nullptr) {
#pragma unroll
    for (int k = 0; k < FIN_G; ++k) acc += red[k][threadIdx.x];
  }
  double q = 0.0;
  if (c < fp.C) {
    if (j >= fp.Preal) {  // residual sum of a predictor with centred splines: used by center_fix_kernel
      if (grad != nullptr) vsum[(size_t)(j - fp.Preal) * fp.Cpad + c] = (T)acc;
    } else {
      for (int b = 0; b < fp.nblocks; ++b) {
        const PriorP& pr = fp.pr[b];
        if (pr.soff < 0 || j < pr.soff || j >= pr.soff + pr.ncoef) continue;
        int jj = j - pr.soff;
        if (!(fp.flags & LSLB_SKIP_PRIOR)) {
          const double xj = (double)pT[(size_t)j * fp.Cpad + c];
          if (pr.prior == LSLB_PRIOR_NORMAL) {
            acc -= (xj - pr.m) * pr.inv_s2;
            q = -0.5 * pr.inv_s2 * (xj - pr.m) * (xj - pr.m);
          } else if (pr.prior == LSLB_PRIOR_MVND) {
            const T* K = reinterpret_cast<const T*>(pr.K) + (size_t)jj * pr.ncoef;
            double kb = 0.0;
            for (int i = 0; i < pr.ncoef; ++i)
              kb += (double)K[i] * (double)pT[(size_t)(pr.soff + i) * fp.Cpad + c];
            double tau2 = pr.tau2_slot >= 0 ? (double)aux[(size_t)c * fp.A + pr.tau2_slot] : pr.tau2_fixed;
            acc -= kb / tau2;
            q = -0.5 * (kb * xj) / tau2;  // row j of x' K x (mvn_degen.py:442)
          }
        }
        if (grad != nullptr) grad[(size_t)c * fp.P + pr.off + jj] = (T)acc;
        break;
      }
    }
  }
  qf[(size_t)j * fp.Cpad + c] = q;
}

// grad[c, block] -= cbar * (residual sum of the block's predictor): the rank-one part of
// (B - 1 cbar^T)^T r for centred spline blocks. Thread per (coefficient of a centred block, chain).
template <typename T>
__global__ void center_fix_kernel(const __grid_constant__ FinalizeParams fp,
                                  const T* __restrict__ vsum, T* __restrict__ grad) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  const PriorP& pr = fp.pr[b];
  if (c >= fp.C || pr.vrow < 0) return;
  const T R = vsum[(size_t)(pr.vrow - fp.Preal) * fp.Cpad + c];
  const T* cb = reinterpret_cast<const T*>(pr.center);
  T* g = grad + (size_t)c * fp.P + pr.off;
  for (int t = 0; t < pr.ncoef; ++t) g[t] -= cb[t] * R;
}

template <typename T, int FIN_G>
__device__ __forceinline__ void finalize_logp_body(const FinalizeParams& fp, const T* __restrict__ partial,
                                                   const T* __restrict__ pT, const T* __restrict__ aux,
                                                   const T* __restrict__ gpl, int gtiles, int G,
                                                   const double* qf, T* __restrict__ logp) {
  // runs in the LAST CTA of a chain group inside finalize_rows_kernel: qf was written by other CTAs of
  // the same launch, so it is read with ld.global.cg (L2), never through the non-coherent path
  __shared__ double red[FIN_G][32];
  const int c = blockIdx.x * 32 + threadIdx.x;  // < Cpad (Cpad is a multiple of 32)
  const int g = threadIdx.y;
  const bool with_prior = !(fp.flags & LSLB_SKIP_PRIOR);
  (void)G;
  // log-likelihood: summed over the row chunks by the extra slice of finalize_rows_kernel (row Psmall of qf)
  double acc = g == 0 ? __ldcg(qf + (size_t)fp.Psmall * fp.Cpad + c) : 0.0;
  // quadratic forms of the priors: one number per coefficient row from finalize_grad_body, plus the
  // group block's per-tile partials
  if (with_prior) {
    for (int j = g; j < fp.Preal; j += FIN_G) acc += __ldcg(qf + (size_t)j * fp.Cpad + c);
    for (int b = 0; b < fp.nblocks; ++b) {
      const PriorP& pr = fp.pr[b];
      if (pr.prior == LSLB_PRIOR_NORMAL && pr.soff < 0) {  // group block: reduced per 32-group tile
        double q = 0.0;
        for (int t = g; t < gtiles; t += FIN_G) q += (double)gpl[(size_t)t * fp.Cpad + c];
        acc += q;
      }
    }
  }
  red[g][threadIdx.x] = acc;
  __syncthreads();
  if (g != 0 || c >= fp.C) return;
  const double half_log_2pi = 0.91893853320467274178;
  acc = 0.0;
#pragma unroll
  for (int k = 0; k < FIN_G; ++k) acc += red[k][threadIdx.x];
  acc += fp.lik_const;
  if (with_prior) {
    acc += fp.const_term;
    for (int b = 0; b < fp.nblocks; ++b) {
      const PriorP& pr = fp.pr[b];
      if (pr.prior == LSLB_PRIOR_NORMAL) {
        acc -= (double)pr.ncoef * (pr.log_s + half_log_2pi);
      } else if (pr.prior == LSLB_PRIOR_MVND) {
        double tau2 = pr.tau2_slot >= 0 ? (double)aux[(size_t)c * fp.A + pr.tau2_slot] : pr.tau2_fixed;
        double log_tau2 = log(tau2);
        // 0.5 * (-x'Kx/tau2 - (rank*log(2 pi) - (log_pdet - rank*log tau2))): constants here
        acc += -0.5 * (pr.rank * 2.0 * half_log_2pi - (pr.log_pdet - pr.rank * log_tau2));
        if (pr.has_ig && pr.tau2_slot >= 0)
          acc += pr.ig_const - (pr.ig_a + 1.0) * log_tau2 - pr.ig_b / tau2;
      }
    }
  }
  logp[c] = (T)acc;
}

// Extra slice of finalize_rows_kernel (blockIdx.y == Psmall): the chains' log-likelihood, summed over
// the row chunks (Kahan pairs hi - lo) in double, into row Psmall of qf.
template <typename T, int FIN_G>
__device__ __forceinline__ void finalize_loglik_body(const FinalizeParams& fp, const T* __restrict__ partial,
                                                     double* qf) {
  __shared__ double red[FIN_G][32];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int g = threadIdx.y;
  const size_t stride = (size_t)(fp.Psmall + LSLB_NACC) * fp.Cpad;
  const T* hi = partial + (size_t)fp.Psmall * fp.Cpad + c;
  const T* lo = hi + fp.Cpad;
  double acc = 0.0;
  int ch = g;
  for (; ch + 3 * FIN_G < fp.nchunks_logp; ch += 4 * FIN_G) {
    T vh[4], vl[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      vh[u] = hi[(size_t)(ch + u * FIN_G) * stride];
      vl[u] = lo[(size_t)(ch + u * FIN_G) * stride];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) acc += (double)vh[u] - (double)vl[u];
  }
  for (; ch < fp.nchunks_logp; ch += FIN_G) acc += (double)hi[(size_t)ch * stride] - (double)lo[(size_t)ch * stride];
  red[g][threadIdx.x] = acc;
  __syncthreads();
  if (g != 0) return;
  acc = 0.0;
#pragma unroll
  for (int k = 0; k < FIN_G; ++k) acc += red[k][threadIdx.x];
  qf[(size_t)fp.Psmall * fp.Cpad + c] = acc;
}

// ONE launch per call. grid.y = Psmall + 1 slices per 32 chains: gradient rows + the rows' shares of the
// priors' quadratic forms (value-only calls: grad == NULL) and the log-likelihood slice. The LAST slice
// of a chain group to finish (a counter per chain group, zeroed by gather_small_kernel earlier in the
// same call; __threadfence before the increment, coherent loads after it) adds the Psmall + 1 numbers
// and the constants into the log-posterior. The sum is over fixed rows in a fixed order: which CTA
// happens to be last does not change a bit of the result.
template <typename T, int FIN_G>
__global__ void __launch_bounds__(32 * FIN_G)
finalize_rows_kernel(const __grid_constant__ FinalizeParams fp, const T* __restrict__ partial,
                     const T* __restrict__ pT, const T* __restrict__ aux, T* __restrict__ grad,
                     T* __restrict__ vsum, double* qf, const T* __restrict__ gpl, int gtiles, int G,
                     int* __restrict__ counters, T* __restrict__ logp) {
  __shared__ int is_last;
  if ((int)blockIdx.y < fp.Psmall) finalize_grad_body<T, FIN_G>(fp, partial, pT, aux, grad, vsum, qf);
  else finalize_loglik_body<T, FIN_G>(fp, partial, qf);
  __threadfence();  // this CTA's qf row is visible device-wide before the counter moves
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0)
    is_last = atomicAdd(&counters[blockIdx.x], 1) == (int)gridDim.y - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  finalize_logp_body<T, FIN_G>(fp, partial, pT, aux, gpl, gtiles, G, qf, logp);
}

// =======================================================================================
// host side
// =======================================================================================
LaunchCfg choose_cfg(const lslb_plan* plan, int C, size_t smem_per_chain, size_t smem_fixed) {
  LaunchCfg k;
  k.TC = C > 64 ? 128 : (C > 32 ? 64 : 32);
  while (smem_per_chain * k.TC > 200 * 1024 && k.TC > 32) k.TC /= 2;  // many coefficients
  k.Cpad = (C + k.TC - 1) / k.TC * k.TC;
  k.ntiles = k.Cpad / k.TC;
  int ctas_per_sm = 512 / k.TC;
  {  // resident CTAs are bounded by shared memory: size the grid to whole waves of what fits
    size_t per_cta = smem_per_chain * k.TC + smem_fixed + 1024;
    int fit = (int)((size_t)227 * 1024 / per_cta);
    if (fit < 1) fit = 1;
    if (fit < ctas_per_sm) ctas_per_sm = fit;
  }
  long long want = ((long long)plan->sm_count * ctas_per_sm + k.ntiles - 1) / k.ntiles;
  if (want < 1) want = 1;
  int F = plan->nchunks;
  k.fstride = (int)((F + want - 1) / want);
  if (k.fstride < 1) k.fstride = 1;
  k.nchunks = (F + k.fstride - 1) / k.fstride;
  return k;
}

// spline-only plans run the staged form of eval_kernel: gradient columns only + two row stages
static bool eval_staged(const lslb_plan* plan) {
  return plan->variant == VAR_BANDED;
}
static size_t eval_smem_per_chain(const lslb_plan* plan) {
  return (size_t)(eval_staged(plan) ? 1 : 2) * plan->Psmall * (plan->desc.dtype == LSLB_F32 ? 4 : 8);
}
static size_t eval_smem_fixed(const lslb_plan* plan) {
  if (!eval_staged(plan)) return 0;
  const size_t ts = plan->desc.dtype == LSLB_F32 ? 4 : 8;
  const int NS = plan->ns <= 1 ? 1 : (plan->ns <= 3 ? plan->ns : (plan->ns <= 5 ? 5 : 8));  // dispatch_structure
  return 16 + 2 * ((size_t)NS * (128 * ts + 128) + 32 * ts);
}

// launch geometry of the logp/grad pass: the TMA run-length kernel when the plan qualifies
static LaunchCfg eval_cfg(const lslb_plan* plan, int C) {
  if (dense_tc_eligible(plan, C)) {
    TcCfg t = dense_tc_cfg(plan, C);
    LaunchCfg k;
    k.TC = 128;
    k.Cpad = t.Cpad;
    k.ntiles = t.ntiles_chain;
    k.fstride = 0;
    k.nchunks = t.nchunks;
    return k;
  }
  const bool dense = dense_x2_eligible(plan);
  if (!dense && dgroup_x2_eligible(plan)) return dgroup_x2_cfg(plan, C);
  if (!dense && msort_x2_eligible(plan)) return msort_x2_cfg(plan, C);
  if (!dense && msmooth_x2_eligible(plan)) return msmooth_x2_cfg(plan, C);
  if (dense || banded_fast_eligible(plan)) {
    BandedCfg b = dense ? dense_x2_cfg(plan, C) : banded_cfg(plan, C);
    LaunchCfg k;
    k.TC = b.TC;
    k.Cpad = b.Cpad;
    k.ntiles = b.ntiles_chain;
    k.fstride = 0;
    k.nchunks = b.nchunks;
    return k;
  }
  return choose_cfg(plan, C, eval_smem_per_chain(plan), eval_smem_fixed(plan));
}

template <typename T>
struct WsLayout {
  T *pT, *partial, *bT, *gbT, *gpl, *rt, *vsum;
  int* counters;  // [Cpad / 32] finished slices per chain group (last-CTA-done reduction of the log-posterior)
  double* qf;  // [(Psmall + 1) x Cpad] per-row shares of the priors' quadratic forms; last row: log-likelihood
  int gtiles;
  size_t bytes;
};

template <typename T>
static WsLayout<T> layout_ws(const lslb_plan* plan, const LaunchCfg& k, void* ws) {
  WsLayout<T> L;
  size_t off = 0;
  auto take = [&](size_t n) {
    T* ptr = reinterpret_cast<T*>(reinterpret_cast<char*>(ws) + off);
    off += align_up(n * sizeof(T), 256);
    return ptr;
  };
  L.pT = take((size_t)plan->Psmall * k.Cpad);
  L.partial = take((size_t)k.nchunks * (plan->Psmall + LSLB_NACC) * k.Cpad);
  L.vsum = take((size_t)2 * k.Cpad);  // residual sums of the (at most two) centring rows
  L.qf = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + off);
  off += align_up((size_t)(plan->Psmall + 1) * k.Cpad * sizeof(double), 256);
  L.counters = reinterpret_cast<int*>(reinterpret_cast<char*>(ws) + off);  // one per 32 chains
  off += align_up((size_t)(k.Cpad / 32) * sizeof(int), 256);
  L.gtiles = 0;
  L.bT = L.gbT = L.gpl = nullptr;
  if (plan->grp >= 0) {
    int G = plan->blk[plan->grp].ncoef;
    L.gtiles = (G + 31) / 32;
    L.bT = take((size_t)G * k.Cpad);
    L.gbT = take((size_t)G * k.Cpad);
    L.gpl = take((size_t)L.gtiles * k.Cpad);
  }
  L.rt = nullptr;
  if (plan->d_XT && k.TC == 128 && k.fstride == 0 && dense_tc_eligible(plan, k.Cpad))
    L.rt = take(dense_tc_rt_floats(plan, k.Cpad));
  L.bytes = off;
  return L;
}

size_t iwls_workspace_bytes(const lslb_plan* plan, int C);  // iwls.cu

size_t workspace_bytes(const lslb_plan* plan, int C) {
  if (C < 1) return 0;
  LaunchCfg k = eval_cfg(plan, C);
  size_t a = plan->desc.dtype == LSLB_F32 ? layout_ws<float>(plan, k, nullptr).bytes
                                          : layout_ws<double>(plan, k, nullptr).bytes;
  size_t b = iwls_workspace_bytes(plan, C);
  return a > b ? a : b;
}

void fill_finalize_params(const lslb_plan* plan, FinalizeParams& fp, int C, int Cpad, int nchunks,
                          unsigned flags) {
  fp.nblocks = plan->desc.nblocks;
  for (int b = 0; b < fp.nblocks; ++b) {
    const lslb_block_desc& d = plan->blk[b];
    PriorP& pr = fp.pr[b];
    pr.soff = plan->soff[b];
    pr.off = plan->off[b];
    pr.ncoef = d.ncoef;
    pr.prior = d.prior;
    pr.m = d.prior_m;
    pr.inv_s2 = d.prior == LSLB_PRIOR_NORMAL ? 1.0 / (d.prior_s * d.prior_s) : 0.0;
    pr.log_s = d.prior == LSLB_PRIOR_NORMAL ? log(d.prior_s) : 0.0;
    pr.K = d.K;
    pr.rank = d.rank;
    pr.log_pdet = d.log_pdet;
    pr.tau2_fixed = d.tau2_fixed;
    pr.ig_a = d.ig_a;
    pr.ig_b = d.ig_b;
    pr.ig_const = d.has_ig ? d.ig_a * log(d.ig_b) - lgamma(d.ig_a) : 0.0;
    pr.tau2_slot = d.tau2_slot;
    pr.has_ig = d.has_ig;
    pr.vrow = d.center ? plan->vrow[d.predictor] : -1;
    pr.center = d.center;
  }
  fp.Preal = plan->Preal;
  fp.P = plan->P;
  fp.A = plan->A;
  fp.Psmall = plan->Psmall;
  fp.C = C;
  fp.Cpad = Cpad;
  fp.nchunks = nchunks;
  fp.nchunks_logp = nchunks;
  fp.lik_const = plan->lik_const;
  fp.const_term = plan->desc.const_term;
  fp.flags = flags;
}

template <typename T>
void fill_eval_params(const lslb_plan* plan, EvalParams<T>& ep, int C, int Cpad) {
  ep.n = plan->desc.n;
  ep.y = reinterpret_cast<const T*>(plan->desc.y);
  ep.fixed_inv_scale = plan->fam == FAM_NORMAL_FIXED ? (T)(1.0 / plan->desc.fixed_scale) : T(1);
  ep.ns = plan->ns;
  ep.nd = plan->nd;
  ep.ni = plan->ni;
  for (int m = 0; m < plan->ns; ++m) {
    const lslb_block_desc& d = plan->blk[plan->sp[m]];
    ep.sp[m] = {reinterpret_cast<const T*>(d.data), d.index, plan->soff[plan->sp[m]], d.ncoef,
                d.predictor, 0, plan->d_tw[m], plan->d_wp[m]};
  }
  for (int m = 0; m < plan->nd; ++m) {
    const lslb_block_desc& d = plan->blk[plan->dn[m]];
    ep.dn[m] = {reinterpret_cast<const T*>(d.data), (long long)d.ld, plan->soff[plan->dn[m]],
                d.ncoef, d.predictor, 0};
  }
  for (int m = 0; m < plan->ni; ++m)
    ep.ic[m] = {plan->soff[plan->ic[m]], plan->blk[plan->ic[m]].predictor};
  for (int k = 0; k < 2; ++k)  // centring terms enter as one more intercept per predictor
    if (plan->vrow[k] >= 0) ep.ic[ep.ni++] = {plan->vrow[k], k};
  ep.gidx = nullptr;
  ep.gpred = 0;
  ep.G = 0;
  ep.bT = nullptr;
  ep.gbT = nullptr;
  if (plan->grp >= 0) {
    ep.gidx = plan->blk[plan->grp].index;
    ep.gpred = plan->blk[plan->grp].predictor;
    ep.G = plan->blk[plan->grp].ncoef;
  }
  ep.chunk_start = plan->d_chunk_start;
  ep.nchunks = plan->nchunks;
  ep.C = C;
  ep.Cpad = Cpad;
  ep.Psmall = plan->Psmall;
}
template void fill_eval_params<float>(const lslb_plan*, EvalParams<float>&, int, int);
template void fill_eval_params<double>(const lslb_plan*, EvalParams<double>&, int, int);

template <typename T, int FAM, int NS, int PD, bool GROUP>
static int launch_eval(const EvalParams<T>& ep, const LaunchCfg& k, bool want_grad, size_t smem,
                       cudaStream_t stream) {
  dim3 grid(k.nchunks, k.ntiles), block(k.TC);
  if (want_grad) {
    auto kern = eval_kernel<T, FAM, NS, PD, GROUP, true>;
    if (smem > 48 * 1024)
      LSLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, block, smem, stream>>>(ep, k.fstride);
  } else {
    auto kern = eval_kernel<T, FAM, NS, PD, GROUP, false>;
    if (smem > 48 * 1024)
      LSLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, block, smem, stream>>>(ep, k.fstride);
  }
  LSLB_LAUNCH_CHECK();
  return LSLB_OK;
}

template <typename T, int FAM>
static int dispatch_structure(const lslb_plan* plan, const EvalParams<T>& ep, const LaunchCfg& k,
                              bool g, size_t smem, cudaStream_t st) {
  const bool grp = plan->grp >= 0;
  if (plan->variant == VAR_BANDED) {
    if (plan->ns <= 1) return launch_eval<T, FAM, 1, 0, false>(ep, k, g, smem, st);
    if (plan->ns <= 2) return launch_eval<T, FAM, 2, 0, false>(ep, k, g, smem, st);
    if (plan->ns <= 3) return launch_eval<T, FAM, 3, 0, false>(ep, k, g, smem, st);
    if (plan->ns <= 5) return launch_eval<T, FAM, 5, 0, false>(ep, k, g, smem, st);
    return launch_eval<T, FAM, 8, 0, false>(ep, k, g, smem, st);
  }
  if (plan->variant == VAR_DENSE_REG) {
    if (grp) return launch_eval<T, FAM, 0, 16, true>(ep, k, g, smem, st);
    return launch_eval<T, FAM, 0, 16, false>(ep, k, g, smem, st);
  }
  if (grp) return launch_eval<T, FAM, 8, -1, true>(ep, k, g, smem, st);
  return launch_eval<T, FAM, 8, -1, false>(ep, k, g, smem, st);
}

template <typename T>
int launch_logp_grad(const lslb_plan* plan, int C, const T* params, const T* aux, T* logp, T* grad,
                     unsigned flags, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (!plan || C < 1 || !params || !logp || !ws) return LSLB_ERR_INVALID;
  if (plan->A > 0 && !aux) return LSLB_ERR_INVALID;
  if ((plan->desc.dtype == LSLB_F32) != (sizeof(T) == 4)) return LSLB_ERR_INVALID;
  const bool tcp = dense_tc_eligible(plan, C);
  const bool dense = !tcp && dense_x2_eligible(plan);
  const bool dgrp = !tcp && !dense && dgroup_x2_eligible(plan);
  const bool msr = !tcp && !dense && !dgrp && msort_x2_eligible(plan);
  const bool msm = !tcp && !dense && !dgrp && !msr && msmooth_x2_eligible(plan);
  const bool fast = !tcp && !dense && !dgrp && !msm && !msr && banded_fast_eligible(plan);
  LaunchCfg k = eval_cfg(plan, C);
  size_t smem = (fast || dense || dgrp || tcp || msm || msr) ? 0 : eval_smem_per_chain(plan) * k.TC + eval_smem_fixed(plan);
  if (smem > 200 * 1024) return LSLB_ERR_UNSUPPORTED;
  WsLayout<T> L = layout_ws<T>(plan, k, ws);
  if (L.bytes > ws_bytes) return LSLB_ERR_WORKSPACE;  // (the ABI's contract, also where the workspace stays unused)
  if (tiny_eligible(plan, C)) {  // small regression: everything in one launch, workspace untouched (tiny.cu)
    FinalizeParams fpt;
    fill_finalize_params(plan, fpt, C, C, 1, flags);
    EvalParams<T> ept;
    fill_eval_params<T>(plan, ept, C, C);
    return launch_tiny<T>(plan, ept, fpt, C, params, logp, grad, stream);
  }

  FinalizeParams fp;
  fill_finalize_params(plan, fp, C, k.Cpad, k.nchunks, flags);
  if (tcp) fp.nchunks_logp = dense_tc_cfg(plan, C).nchunks1;
  EvalParams<T> ep;
  fill_eval_params<T>(plan, ep, C, k.Cpad);
  ep.pT = L.pT;
  ep.partial = L.partial;
  ep.bT = L.bT;
  ep.gbT = L.gbT;

  if (int gst = launch_gather_small<T>(fp, params, L.pT, stream, L.counters, k.Cpad / 32)) return gst;
  int G = 0;
  if (plan->grp >= 0) {
    const lslb_block_desc& d = plan->blk[plan->grp];
    G = d.ncoef;
    bool with_prior = d.prior == LSLB_PRIOR_NORMAL && !(flags & LSLB_SKIP_PRIOR);
    T inv_s2 = d.prior == LSLB_PRIOR_NORMAL ? (T)(1.0 / (d.prior_s * d.prior_s)) : T(0);
    dim3 grid((G + 31) / 32, (k.Cpad + 31) / 32), block(32, 8);
    group_init_kernel<T><<<grid, block, 0, stream>>>(params, plan->P, plan->off[plan->grp], G, C,
                                                     k.Cpad, (T)d.prior_m, inv_s2, with_prior,
                                                     L.bT, L.gbT, L.gpl);
    LSLB_LAUNCH_CHECK();
  }

  const bool g = grad != nullptr;
  int st = LSLB_ERR_INVALID;
  debug_record_start(stream);
  if (tcp) {
    if constexpr (sizeof(T) == 4)
      st = launch_dense_tc(plan, dense_tc_cfg(plan, C), C, params, L.rt, L.partial, g, stream);
  } else if (dense) {
    if constexpr (sizeof(T) == 4) st = launch_dense_x2(plan, ep, dense_x2_cfg(plan, C), g, stream);
  } else if (dgrp) {
    if constexpr (sizeof(T) == 4) st = launch_dgroup_x2(plan, ep, k, g, stream);
  } else if (msr) {
    if constexpr (sizeof(T) == 4) st = launch_msort_x2(plan, ep, k, g, stream);
  } else if (msm) {
    if constexpr (sizeof(T) == 4) st = launch_msmooth_x2(plan, ep, k, g, stream);
  } else if (fast) {
    st = launch_banded<T>(plan, ep, banded_cfg(plan, C), g, stream);
  } else
  switch (plan->fam) {
    case FAM_NORMAL_LS: st = dispatch_structure<T, FAM_NORMAL_LS>(plan, ep, k, g, smem, stream); break;
    case FAM_NORMAL_FIXED: st = dispatch_structure<T, FAM_NORMAL_FIXED>(plan, ep, k, g, smem, stream); break;
    case FAM_BERNOULLI: st = dispatch_structure<T, FAM_BERNOULLI>(plan, ep, k, g, smem, stream); break;
    case FAM_POISSON: st = dispatch_structure<T, FAM_POISSON>(plan, ep, k, g, smem, stream); break;
  }
  debug_record_stop(stream);
  if (st != LSLB_OK) return st;

  {
    // rows: 8 helpers per chain for wide launches, 32 when the grid would leave most SMs idle
    dim3 fgrid(k.Cpad / 32, plan->Psmall + 1);  // last slice: the log-likelihood
    T* gout = g ? grad : nullptr;
    if ((long long)fgrid.x * fgrid.y >= 4LL * plan->sm_count)
      finalize_rows_kernel<T, 8><<<fgrid, dim3(32, 8), 0, stream>>>(fp, L.partial, L.pT, aux, gout, L.vsum, L.qf,
                                                                    L.gpl, L.gtiles, G, L.counters, logp);
    else
      finalize_rows_kernel<T, 32><<<fgrid, dim3(32, 32), 0, stream>>>(fp, L.partial, L.pT, aux, gout, L.vsum, L.qf,
                                                                      L.gpl, L.gtiles, G, L.counters, logp);
    LSLB_LAUNCH_CHECK();
  }
  if (g) {
    if (plan->Psmall > plan->Preal) {
      dim3 cgrid((C + 127) / 128, plan->desc.nblocks);
      center_fix_kernel<T><<<cgrid, 128, 0, stream>>>(fp, L.vsum, grad);
      LSLB_LAUNCH_CHECK();
    }
    if (plan->grp >= 0) {
      dim3 grid((G + 31) / 32, (k.Cpad + 31) / 32), block(32, 8);
      group_scatter_kernel<T><<<grid, block, 0, stream>>>(L.gbT, plan->P, plan->off[plan->grp], G,
                                                          C, k.Cpad, grad);
      LSLB_LAUNCH_CHECK();
    }
  }
  return LSLB_OK;
}

template int launch_logp_grad<float>(const lslb_plan*, int, const float*, const float*, float*,
                                     float*, unsigned, void*, size_t, cudaStream_t);
template int launch_logp_grad<double>(const lslb_plan*, int, const double*, const double*, double*,
                                      double*, unsigned, void*, size_t, cudaStream_t);

}  // namespace lslb

extern "C" const char* lslb_plan_kernel_name(const lslb_plan* plan, int C) {
  using namespace lslb;
  if (!plan || C < 1) return "";
  if (tiny_eligible(plan, C)) return "lslb::tiny_kernel";
  if (dense_tc_eligible(plan, C)) return "lslb::dense_tc_k1 + lslb::dense_tc_k2";
  if (dense_x2_eligible(plan)) return "lslb::dense_x2_kernel";
  if (dgroup_x2_eligible(plan)) return "lslb::dgroup_x2_kernel";
  if (msort_x2_eligible(plan)) return "lslb::msort_x2_kernel";
  if (msmooth_x2_eligible(plan)) return "lslb::msmooth_x2_kernel";
  if (banded_fast_eligible(plan))
    return banded_cfg(plan, C).TC == 256 ? "lslb::banded_x2_kernel" : "lslb::banded_kernel";
  return "lslb::eval_kernel";
}

extern "C" int lslb_logp_grad_f32(const lslb_plan* plan, int C, const float* params,
                                  const float* aux, float* logp, float* grad, uint32_t flags,
                                  void* ws, size_t ws_bytes, lslb_stream_t stream) {
  return lslb::launch_logp_grad<float>(plan, C, params, aux, logp, grad, flags, ws, ws_bytes,
                                       (cudaStream_t)stream);
}
extern "C" int lslb_logp_grad_f64(const lslb_plan* plan, int C, const double* params,
                                  const double* aux, double* logp, double* grad, uint32_t flags,
                                  void* ws, size_t ws_bytes, lslb_stream_t stream) {
  return lslb::launch_logp_grad<double>(plan, C, params, aux, logp, grad, flags, ws, ws_bytes,
                                        (cudaStream_t)stream);
}
