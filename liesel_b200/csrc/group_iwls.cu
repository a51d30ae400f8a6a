// IWLS quantities and proposal tail of the RANDOM-INTERCEPT block.
//
// The reference's IWLSKernel takes the score from grad(log_prob) and the information from
// -jacfwd(grad(log_prob)) of the block's position (src/liesel/goose/iwls.py:206-243,315-321), i.e. a dense
// G x G matrix per chain. For b[idx[i]] every row belongs to exactly one group, so that matrix is DIAGONAL:
//     score_g = sum_{i in g} d ll_i / d eta  - (b_g - m) / s^2
//     info_gg = sum_{i in g} w_i             + 1 / s^2          (w = observed information weight)
// and its ridge + Cholesky + solves (iwls.py:232-238, iwls_utils.py:14-70) are elementwise. The two entry
// points below are what lslb_iwls_* and lslb_iwls_propose_* are for the coefficient blocks, with [C x G]
// vectors in place of [C x p x p] matrices (G = 1e5 at S5: the dense form would be 40 GB per chain).
//
//   group_iwls_kernel   CTA = 32 chains x 32 groups, warp = one group at a time: the rows of a group are
//                       contiguous (rows sorted by group; plan->d_gstart), all lanes read the same row
//                       (broadcast loads), each lane evaluates its chain. b is read and the results are
//                       written through 32 x 33 shared-memory tiles, so both sides of the [C x P] / [C x G]
//                       <-> (group, chain) transposition are coalesced. Sums in double, no atomics.
//   propose_diag_kernel CTA = chain: ridge from the mean of the diagonal, NaN/identity fallback with code 2
//                       (IWLSKernel._safe_chol, iwls.py:245-290), mean, draw and log-density; the
//                       log-density is reduced in a fixed order.
#include "common.cuh"

namespace lslb {

constexpr int GI_PD = 16;

template <typename T, int FAM>
__global__ void __launch_bounds__(256)
group_iwls_kernel(const __grid_constant__ EvalParams<T> p, const int* __restrict__ gstart, const T* __restrict__ params,
                  int P, int goff, int doff, int ioff0, int ioff1, int ioff2, int ioff3, T gm, T ginv_s2,
                  T* __restrict__ score, T* __restrict__ info, T* __restrict__ llg, int vec) {
  __shared__ T tb[32][33], ts[32][33], tw[32][33], tl[32][33];
  // chain tiles vary fastest in the grid: the CTAs that walk the same groups' rows run together and share
  // them through L2 (with the group tile fastest every chain tile re-read all rows from DRAM: 6.2 GB per pass
  // at S5 with 1024 chains against 0.2 GB of rows, ncu profiles/r02_s5_iwls_kernels_full.md)
  const int g0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int G = p.G, C = p.C;
  // b tile: params[c0 + cl][goff + g0 + tx] -> tb[cl][tx]
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int cl = ty + 8 * k, g = g0 + tx, c = c0 + cl;
    tb[cl][tx] = (g < G && c < C) ? params[(size_t)c * P + goff + g] : T(0);
  }
  const int c = c0 + tx;
  const bool live = c < C;
  const DnsP<T>& D = p.dn[0];
  const T* th = params + (size_t)(live ? c : 0) * P;
  T beta[GI_PD];
#pragma unroll
  for (int j = 0; j < GI_PD; ++j) beta[j] = (p.nd > 0 && j < D.p) ? th[doff + j] : T(0);
  T o0 = T(0), o1 = T(0);
  {
    const int io[4] = {ioff0, ioff1, ioff2, ioff3};
    for (int q = 0; q < p.ni && q < 4; ++q) {
      if (p.ic[q].pred) o1 += th[io[q]]; else o0 += th[io[q]];
    }
  }
  __syncthreads();
#pragma unroll 1
  for (int k = 0; k < 4; ++k) {
    const int gl = ty + 8 * k, g = g0 + gl;
    double as = 0.0, aw = 0.0, al = 0.0;
    T bg = T(0);
    if (g < G) {
      bg = tb[tx][gl];
      const int ra = gstart[g], rb = gstart[g + 1];
      for (int r = ra; r < rb; ++r) {  // warp-uniform rows: broadcast loads
        T dot = T(0);
        if (p.nd > 0) {
          const T* x = D.X + (size_t)r * D.ld;
          if (vec) {  // p a multiple of 4, rows 16-byte aligned: one load per four columns
#pragma unroll
            for (int j4 = 0; j4 < GI_PD; j4 += 4) {
              if (j4 < D.p) {
                const Vec4<T> v = ld4(x + j4);
                dot = fma(v.a, beta[j4], dot);
                dot = fma(v.b, beta[j4 + 1], dot);
                dot = fma(v.c, beta[j4 + 2], dot);
                dot = fma(v.d, beta[j4 + 3], dot);
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < GI_PD; ++j)
              if (j < D.p) dot = fma(x[j], beta[j], dot);
          }
        }
        T e0 = o0, e1 = o1;
        if (p.nd > 0) { if (D.pred) e1 += dot; else e0 += dot; }
        if (p.gpred) e1 += bg; else e0 += bg;
        T ll, r0, r1, w0, w1;
        lik_row<T, FAM, true>(p.y[r], e0, e1, p.fixed_inv_scale, ll, r0, r1, w0, w1);
        as += (double)(p.gpred ? r1 : r0);
        aw += (double)(p.gpred ? w1 : w0);
        al += (double)ll;
      }
    }
    const double db = (double)bg - (double)gm;
    ts[gl][tx] = (T)(as - db * (double)ginv_s2);
    tw[gl][tx] = (T)(aw + (double)ginv_s2);
    tl[gl][tx] = (T)(al - 0.5 * db * db * (double)ginv_s2);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int cl = ty + 8 * k, g = g0 + tx, cc = c0 + cl;
    if (g < G && cc < C) {
      score[(size_t)cc * G + g] = ts[tx][cl];
      info[(size_t)cc * G + g] = tw[tx][cl];
      if (llg != nullptr) llg[(size_t)cc * G + g] = tl[tx][cl];
    }
  }
}

template <typename T>
int launch_iwls_group(const lslb_plan* plan, int C, const T* params, const T* aux, T* score, T* info, T* llg,
                      unsigned flags, cudaStream_t st) {
  (void)aux;
  if (!plan || C < 1 || !params || !score || !info) return LSLB_ERR_INVALID;
  if ((plan->desc.dtype == LSLB_F32) != (sizeof(T) == 4)) return LSLB_ERR_INVALID;
  // the S5 family of models: a random intercept next to at most one small dense block and intercepts
  if (plan->grp < 0 || !plan->d_gstart || plan->ns != 0 || plan->nd > 1 || plan->ni > 4) return LSLB_ERR_UNSUPPORTED;
  if (plan->nd == 1 && plan->blk[plan->dn[0]].ncoef > GI_PD) return LSLB_ERR_UNSUPPORTED;
  EvalParams<T> ep;
  fill_eval_params<T>(plan, ep, C, C);
  const lslb_block_desc& d = plan->blk[plan->grp];
  const bool with_prior = d.prior == LSLB_PRIOR_NORMAL && !(flags & LSLB_SKIP_PRIOR);
  const T inv_s2 = with_prior ? (T)(1.0 / (d.prior_s * d.prior_s)) : T(0);
  int io[4] = {0, 0, 0, 0};
  for (int q = 0; q < plan->ni; ++q) io[q] = plan->off[plan->ic[q]];
  const int doff = plan->nd ? plan->off[plan->dn[0]] : 0;
  const int G = d.ncoef;
  int vec = 0;
  if (plan->nd == 1) {
    const lslb_block_desc& dd = plan->blk[plan->dn[0]];
    vec = dd.ncoef % 4 == 0 && dd.ld % 4 == 0 && (reinterpret_cast<uintptr_t>(dd.data) & (4 * sizeof(T) - 1)) == 0;
  }
  if ((G + 31) / 32 > 65535) return LSLB_ERR_UNSUPPORTED;  // grid.y
  dim3 grid((C + 31) / 32, (G + 31) / 32), block(32, 8);
#define LSLB_GI_GO(F)                                                                                        \
  group_iwls_kernel<T, F><<<grid, block, 0, st>>>(ep, plan->d_gstart, params, plan->P, plan->off[plan->grp], doff, \
                                                  io[0], io[1], io[2], io[3], (T)d.prior_m, inv_s2, score, info, llg, vec)
  switch (plan->fam) {
    case FAM_NORMAL_LS: LSLB_GI_GO(FAM_NORMAL_LS); break;
    case FAM_NORMAL_FIXED: LSLB_GI_GO(FAM_NORMAL_FIXED); break;
    case FAM_BERNOULLI: LSLB_GI_GO(FAM_BERNOULLI); break;
    case FAM_POISSON: LSLB_GI_GO(FAM_POISSON); break;
    default: return LSLB_ERR_UNSUPPORTED;
  }
#undef LSLB_GI_GO
  LSLB_LAUNCH_CHECK();
  return LSLB_OK;
}

// ---------------------------------------------------------------------------------------------------
// proposal tail with a diagonal information (iwls.py:323-327,335-336 with chol = diag(sqrt(info + ridge)))
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
propose_diag_kernel(int p, const T* __restrict__ beta, const T* __restrict__ score, const T* __restrict__ info,
                    const T* __restrict__ step, const T* __restrict__ z, const T* __restrict__ x_eval,
                    T* __restrict__ prop, T* __restrict__ logq, int* __restrict__ code, T* __restrict__ logq_el,
                    int local_ridge) {
  __shared__ double red[8];
  __shared__ int bad_any;
  __shared__ double ridge_s;
  const int c = blockIdx.x, tid = threadIdx.x;
  const size_t base = (size_t)c * p;
  auto block_sum = [&](double v) -> double {  // fixed order: lanes (xor tree), then the warps by index
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    return t;
  };
  // ridge = 1e-6 * mean(diag(info))  (iwls.py:233-237)
  double s = 0.0;
  for (int j = tid; j < p; j += 256) s += (double)info[base + j];
  const double ridge = 1e-6 * block_sum(s) / (double)p;
  if (tid == 0) { bad_any = 0; ridge_s = ridge; }
  __syncthreads();
  // Cholesky of a diagonal matrix: NaN (all of it) unless every entry is positive and finite
  int bad = 0;
  for (int j = tid; j < p; j += 256) {
    const double v = local_ridge ? (double)info[base + j] : (double)info[base + j] + ridge;
    if (!(v > 0.0) || !isfinite(v)) bad = 1;
  }
  if (bad) atomicOr(&bad_any, 1);  // (a flag, not a sum: the result does not depend on the order)
  __syncthreads();
  const bool fallback = bad_any != 0;  // identity, error code 2 (_safe_chol, iwls.py:245-290)
  const double st = (double)step[c];
  double lq = 0.0;
  for (int j = tid; j < p; j += 256) {
    // local_ridge: the ridge of entry j is 1e-6 of ITSELF, so that a group's proposal depends on that group alone
    const double ij = (double)info[base + j];
    const double dj = fallback ? 1.0 : sqrt(local_ridge ? ij * (1.0 + 1e-6) : ij + ridge_s);
    const double mu = (double)beta[base + j] + 0.5 * st * st * (double)score[base + j] / (dj * dj);
    double x;
    if (z != nullptr) {
      x = mu + st * (double)z[base + j] / dj;
      prop[base + j] = (T)x;
      x = (double)(T)x;  // the density is evaluated at the stored proposal
    } else {
      x = (double)x_eval[base + j];
    }
    const double u = (x - mu) * dj / st;
    const double le = -0.5 * u * u - 0.91893853320467274178 + log(dj / st);
    if (logq_el != nullptr) logq_el[base + j] = (T)le;
    lq += le;
  }
  lq = block_sum(lq);
  if (tid == 0) {
    logq[c] = (T)lq;
    if (code) code[c] = fallback ? 2 : 0;
  }
}

template <typename T>
int launch_propose_diag(int C, int p, const T* beta, const T* score, const T* info, const T* step, const T* z,
                        const T* x_eval, T* prop, T* logq, int* code, T* logq_el, int local_ridge, cudaStream_t st) {
  if (C < 1 || p < 1 || !beta || !score || !info || !step || !logq) return LSLB_ERR_INVALID;
  if ((z == nullptr) == (x_eval == nullptr)) return LSLB_ERR_INVALID;  // exactly one of them
  if (z != nullptr && !prop) return LSLB_ERR_INVALID;
  propose_diag_kernel<T><<<C, 256, 0, st>>>(p, beta, score, info, step, z, x_eval, prop, logq, code, logq_el, local_ridge);
  LSLB_LAUNCH_CHECK();
  return LSLB_OK;
}

}  // namespace lslb

extern "C" int lslb_iwls_group_f32(const lslb_plan* plan, int C, const float* params, const float* aux,
                                   float* score, float* info_diag, float* loglik_group, uint32_t flags,
                                   lslb_stream_t stream) {
  return lslb::launch_iwls_group<float>(plan, C, params, aux, score, info_diag, loglik_group, flags,
                                      (cudaStream_t)stream);
}
extern "C" int lslb_iwls_group_f64(const lslb_plan* plan, int C, const double* params, const double* aux,
                                   double* score, double* info_diag, double* loglik_group, uint32_t flags,
                                   lslb_stream_t stream) {
  return lslb::launch_iwls_group<double>(plan, C, params, aux, score, info_diag, loglik_group, flags,
                                      (cudaStream_t)stream);
}
extern "C" int lslb_iwls_propose_diag_f32(int C, int p, const float* beta, const float* score,
                                          const float* info_diag, const float* step, const float* z,
                                          const float* x_eval, float* prop, float* logq, int* code,
                                          float* logq_elem, int local_ridge, lslb_stream_t stream) {
  return lslb::launch_propose_diag<float>(C, p, beta, score, info_diag, step, z, x_eval, prop, logq, code,
                                          logq_elem, local_ridge, (cudaStream_t)stream);
}
extern "C" int lslb_iwls_propose_diag_f64(int C, int p, const double* beta, const double* score,
                                          const double* info_diag, const double* step, const double* z,
                                          const double* x_eval, double* prop, double* logq, int* code,
                                          double* logq_elem, int local_ridge, lslb_stream_t stream) {
  return lslb::launch_propose_diag<double>(C, p, beta, score, info_diag, step, z, x_eval, prop, logq, code,
                                           logq_elem, local_ridge, (cudaStream_t)stream);
}
