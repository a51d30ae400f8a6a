// Small regressions in ONE launch (S1: Bayesian linear regression, n = 1000, p = 10, 4 chains).
//
// The general path is three launches per evaluation (gather -> row kernel -> finalize) through a
// workspace: 30 us per call at S1, all of it launch latency and dependent global round trips, where the
// work itself is 2e5 flop. Here one CTA per chain reads the chain's coefficients straight from
// params [C x P], walks the n rows once (thread = row slot), reduces the log-likelihood and the
// p + intercept gradient sums in double through warp shuffles and shared memory in a FIXED order
// (bitwise reproducible), adds the priors and writes logp[c] and grad[c, :]. No workspace, no atomics.
//
// Reference for what is computed: one `update_state` + `log_prob` + reverse-mode gradient per chain
// (src/liesel/model/model.py:2786-2865, src/liesel/goose/kernel.py:149-158); priors
// `tfd.Normal(m, s).log_prob(beta)` (src/liesel/model/distreg.py:102).
//
// Eligible plans: one dense block with p <= 16 (coefficients in registers), any number of intercepts,
// no spline / group block, priors Normal or none, n <= TINY_MAX_ROWS and n x chains <= TINY_MAX_WORK.
#include "common.cuh"

namespace lslb {

constexpr int TINY_THREADS = 256;
constexpr long long TINY_MAX_ROWS = 4095;        // from 4096 rows the TMA kernel of dgroup_x2.cu shares the rows between chains
constexpr long long TINY_MAX_WORK = 1 << 20;   // rows x chains: every chain's CTA reads all rows (from L2)
constexpr int TINY_PD = 16;
constexpr int TINY_NV = TINY_PD + 3;  // dense gradient sums, residual sums of both predictors, log-likelihood

bool tiny_eligible(const lslb_plan* plan, int C) {
  if (plan->variant != VAR_DENSE_REG || plan->grp >= 0 || plan->ns != 0 || plan->nd != 1) return false;
  if (plan->desc.n > TINY_MAX_ROWS || plan->desc.n < 1 || plan->desc.n * (long long)C > TINY_MAX_WORK) return false;
  if (plan->blk[plan->dn[0]].ncoef > TINY_PD) return false;
  if (plan->Psmall != plan->Preal) return false;
  for (int b = 0; b < plan->desc.nblocks; ++b)
    if (plan->blk[b].prior != LSLB_PRIOR_NONE && plan->blk[b].prior != LSLB_PRIOR_NORMAL) return false;
  static const bool off = [] {  // LSLB_NO_TINY=1: the general three-launch path (A/B runs)
    const char* e = getenv("LSLB_NO_TINY");
    return e && atoi(e) != 0;
  }();
  return !off;
}

template <typename T, int FAM, bool GRAD>
__global__ void __launch_bounds__(TINY_THREADS)
tiny_kernel(const __grid_constant__ EvalParams<T> p, const __grid_constant__ FinalizeParams fp,
            const T* __restrict__ params, T* __restrict__ logp, T* __restrict__ grad) {
  __shared__ double red[TINY_THREADS / 32][TINY_NV];
  __shared__ double tot[TINY_NV];
  const int c = blockIdx.x;
  const int tid = threadIdx.x;
  const DnsP<T>& D = p.dn[0];
  const T* th = params + (size_t)c * fp.P;
  // flat offsets of the blocks: the dense block's and the intercepts' (PriorP::off by small-space offset)
  int doff = 0;
  for (int b = 0; b < fp.nblocks; ++b)
    if (fp.pr[b].soff == D.soff) doff = fp.pr[b].off;
  T beta[TINY_PD];
#pragma unroll
  for (int j = 0; j < TINY_PD; ++j) beta[j] = j < D.p ? th[doff + j] : T(0);
  T o0 = T(0), o1 = T(0);
  for (int q = 0; q < p.ni; ++q) {
    T v = T(0);
    for (int b = 0; b < fp.nblocks; ++b)
      if (fp.pr[b].soff == p.ic[q].soff) v = th[fp.pr[b].off];
    if (p.ic[q].pred) o1 += v; else o0 += v;
  }

  double acc[TINY_NV];
#pragma unroll
  for (int j = 0; j < TINY_NV; ++j) acc[j] = 0.0;
  for (long long r = tid; r < p.n; r += TINY_THREADS) {
    const T* x = D.X + r * D.ld;
    T xv[TINY_PD];
    T dot = T(0);
#pragma unroll
    for (int j = 0; j < TINY_PD; ++j) {
      xv[j] = j < D.p ? x[j] : T(0);
      dot = fma(xv[j], beta[j], dot);
    }
    const T e0 = D.pred ? o0 : o0 + dot;
    const T e1 = D.pred ? o1 + dot : o1;
    T ll, r0, r1, w0, w1;
    lik_row<T, FAM, false>(p.y[r], e0, e1, p.fixed_inv_scale, ll, r0, r1, w0, w1);
    acc[TINY_PD + 2] += (double)ll;
    if constexpr (GRAD) {
      const T rd = D.pred ? r1 : r0;
#pragma unroll
      for (int j = 0; j < TINY_PD; ++j) acc[j] += (double)(xv[j] * rd);
      acc[TINY_PD] += (double)r0;
      acc[TINY_PD + 1] += (double)r1;
    }
  }
  // fixed-order reduction: lanes (xor tree), then the warps in index order
#pragma unroll
  for (int j = 0; j < TINY_NV; ++j) {
    if (!GRAD && j != TINY_PD + 2) continue;
    double v = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) red[tid >> 5][j] = v;
  }
  __syncthreads();
  if (tid < TINY_NV) {
    double v = 0.0;
#pragma unroll
    for (int w = 0; w < TINY_THREADS / 32; ++w) v += red[w][tid];
    tot[tid] = v;
  }
  __syncthreads();

  const bool with_prior = !(fp.flags & LSLB_SKIP_PRIOR);
  // gradient: one thread per coefficient of the flat layout
  if constexpr (GRAD) {
    if (tid < fp.P) {
      for (int b = 0; b < fp.nblocks; ++b) {
        const PriorP& pr = fp.pr[b];
        if (tid < pr.off || tid >= pr.off + pr.ncoef) continue;
        double g;
        if (pr.soff == D.soff) {
          g = tot[tid - pr.off];
        } else {
          int pred = 0;
          for (int q = 0; q < p.ni; ++q)
            if (p.ic[q].soff == pr.soff) pred = p.ic[q].pred;
          g = tot[TINY_PD + pred];
        }
        if (with_prior && pr.prior == LSLB_PRIOR_NORMAL) g -= ((double)th[tid] - pr.m) * pr.inv_s2;
        grad[(size_t)c * fp.P + tid] = (T)g;
        break;
      }
    }
  }
  if (tid == 0) {
    const double half_log_2pi = 0.91893853320467274178;
    double lp = tot[TINY_PD + 2] + fp.lik_const;
    if (with_prior) {
      lp += fp.const_term;
      for (int b = 0; b < fp.nblocks; ++b) {
        const PriorP& pr = fp.pr[b];
        if (pr.prior != LSLB_PRIOR_NORMAL) continue;
        double q = 0.0;
        for (int j = 0; j < pr.ncoef; ++j) {
          const double d = (double)th[pr.off + j] - pr.m;
          q += d * d;
        }
        lp += -0.5 * pr.inv_s2 * q - (double)pr.ncoef * (pr.log_s + half_log_2pi);
      }
    }
    logp[c] = (T)lp;
  }
}

template <typename T>
int launch_tiny(const lslb_plan* plan, const EvalParams<T>& ep, const FinalizeParams& fp, int C, const T* params,
                T* logp, T* grad, cudaStream_t st) {
#define LSLB_TINY_GO(F)                                                                       \
  do {                                                                                        \
    if (grad) tiny_kernel<T, F, true><<<C, TINY_THREADS, 0, st>>>(ep, fp, params, logp, grad); \
    else tiny_kernel<T, F, false><<<C, TINY_THREADS, 0, st>>>(ep, fp, params, logp, grad);     \
  } while (0)
  switch (plan->fam) {
    case FAM_NORMAL_LS: LSLB_TINY_GO(FAM_NORMAL_LS); break;
    case FAM_NORMAL_FIXED: LSLB_TINY_GO(FAM_NORMAL_FIXED); break;
    case FAM_BERNOULLI: LSLB_TINY_GO(FAM_BERNOULLI); break;
    case FAM_POISSON: LSLB_TINY_GO(FAM_POISSON); break;
    default: return LSLB_ERR_UNSUPPORTED;
  }
#undef LSLB_TINY_GO
  LSLB_LAUNCH_CHECK();
  return LSLB_OK;
}
template int launch_tiny<float>(const lslb_plan*, const EvalParams<float>&, const FinalizeParams&, int,
                                const float*, float*, float*, cudaStream_t);
template int launch_tiny<double>(const lslb_plan*, const EvalParams<double>&, const FinalizeParams&, int,
                                 const double*, double*, double*, cudaStream_t);

}  // namespace lslb
