// State machine of the asynchronous batched NUTS driver (liesel_b200/nuts_async.py): ONE kernel between
// two batched log-posterior+gradient evaluations, one warp per chain. Every chain is at its own place
// of its own tree; a step consumes the evaluation at q_n and produces the next q_n:
//   LEAF    : second half step, energy error, divergence, multinomial update of the subtree proposal,
//             momentum sums, checkpointed generalised U-turn test (as nuts.cu's leaf kernel);
//             subtree complete -> biased progressive merge into the tree, U-turn test of the whole tree;
//             tree complete    -> the draw, dual averaging (goose/da.py), mass-matrix window sums and
//                                 epoch ends (goose/mm.py, goose/warmup.py), optional Gibbs update of
//                                 smoothing variances (model/distreg.py:277-297), next transition;
//   REFRESH : the evaluation at the chain's position after a Gibbs update becomes its carried
//             log-posterior and gradient; the next transition starts.
// Semantics are those of AsyncNUTS._step_torch (masked torch ops), which the tests compare against step
// by step; the random numbers of a step (P normals, 3 uniforms, one Gamma variate per Gibbs block and
// chain) are inputs. All reductions are over the chain's own P coordinates in a fixed order.
#include "common.cuh"

namespace lslb {

enum { NA_REFRESH = 0, NA_LEAF = 1, NA_DONE = 2 };
enum { NA_DIVERGED = 1, NA_TURNED = 2, NA_TURN_S = 4, NA_DIV_S = 8 };

template <typename T>
struct NutsAsync {
  int C, P, D1, maxd, W, S, nep, nb, A;
  T thr, target, gamma, kappa, t0;
  T *q, *lp, *g, *step, *inv_mass, *h0, *qL, *rL, *gL, *qR, *rR, *gR, *qp, *lpp, *gp, *log_w, *rsum, *sum_acc,
      *n_leaf, *qe, *re, *ge, *qs, *lps, *gs, *logw_s, *rsum_s, *ck_r, *ck_s, *qn, *r_half;
  const T *lp_n, *g_n, *Z, *U;
  const double* GAM;
  T *da_err, *da_logavg, *da_mu;
  double *ws1, *ws2;
  T *draws, *aux;
  const T* gib_M;
  const double* gib_par;
  int *phase, *depth, *dir, *i_leaf, *n_leaves, *it, *ep, *t_ep, *flags;
  const int *ep_dur, *ep_kind;
  double* stats;
  int* n_done;
};

template <typename T>
__device__ __forceinline__ T na_warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <typename T>
__device__ __forceinline__ T na_logaddexp(T a, T b) {
  const T mx = fmax(a, b), mn = fmin(a, b);
  return (mx == (T)(-INFINITY)) ? (T)(-INFINITY) : mx + log1p(exp(mn - mx));
}

template <typename T>
__global__ void __launch_bounds__(128) nuts_async_step_kernel(const __grid_constant__ NutsAsync<T> a) {
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= a.C) return;
  const int P = a.P;
  const size_t base = (size_t)c * P;
  int ph = a.phase[c];
  if (ph == NA_DONE) return;  // q_n already equals q

  // ---- the chain's scalars (every lane holds them; lane 0 writes them back at the end) ----------------
  int fl = a.flags[c], dep = a.depth[c], dr = a.dir[c], il = a.i_leaf[c], nlv = a.n_leaves[c], itc = a.it[c];
  int epi = a.ep[c], tep = a.t_ep[c];
  T stp = a.step[c], lp = a.lp[c], lpp = a.lpp[c], lps = a.lps[c], logw_s = a.logw_s[c], log_w = a.log_w[c];
  T sum_acc = a.sum_acc[c], n_leaf = a.n_leaf[c], h0 = a.h0[c];
  T da_err = a.da_err[c], da_logavg = a.da_logavg[c], da_mu = a.da_mu[c];
  const T u_dir = a.U[(size_t)c * 3], u_leaf = a.U[(size_t)c * 3 + 1], u_merge = a.U[(size_t)c * 3 + 2];
  bool start_tr = false, new_sub = false;
  const T NINF = (T)(-INFINITY);

  if (ph == NA_REFRESH) {
    lp = a.lp_n[c];
    for (int j = lane; j < P; j += 32) a.g[base + j] = a.g_n[base + j];
    start_tr = true;
  } else {
    // ---- the leaf that was just evaluated -------------------------------------------------------------
    const T e = (T)dr * stp;
    T kin = T(0);
    for (int j = lane; j < P; j += 32) {
      const T rn = a.r_half[base + j] + T(0.5) * e * a.g_n[base + j];
      a.r_half[base + j] = rn;  // r_half now holds r_n
      kin += a.inv_mass[base + j] * rn * rn;
    }
    kin = T(0.5) * na_warp_sum(kin);
    const T lpn = a.lp_n[c];
    T dh = (-lpn + kin) - h0;
    if (dh != dh) dh = (T)INFINITY;
    const bool div_n = dh > a.thr;
    const T logw_n = -dh;
    const T acc_n = exp(fmin(-dh, T(0)));
    const T lw_new = na_logaddexp(logw_s, logw_n);
    const bool take = log(u_leaf) < logw_n - lw_new;  // false when both weights are -inf (NaN)
    const int odd = il & 1;
    const int hi = __popc(il >> 1);
    const int lo = hi - (__ffs(~il) - 1) + 1;
    T* ckr = a.ck_r + (size_t)c * a.D1 * P;
    T* cks = a.ck_s + (size_t)c * a.D1 * P;
    bool turning = false;
    if (odd) {
      for (int k = hi; k >= lo; --k) {
        T dl = T(0), drr = T(0);
        for (int j = lane; j < P; j += 32) {
          const T rn = a.r_half[base + j];
          const T rs = a.rsum_s[base + j] + rn;
          const T rl = ckr[(size_t)k * P + j];
          const T sub = rs - cks[(size_t)k * P + j] + rl;
          const T rho = sub - T(0.5) * (rn + rl);
          const T im = a.inv_mass[base + j];
          dl += im * rl * rho;
          drr += im * rn * rho;
        }
        dl = na_warp_sum(dl);
        drr = na_warp_sum(drr);
        turning = turning || (dl <= T(0)) || (drr <= T(0));
      }
    }
    for (int j = lane; j < P; j += 32) {
      const T rn = a.r_half[base + j];
      const T qn = a.qn[base + j], gn = a.g_n[base + j];
      const T rs = a.rsum_s[base + j] + rn;
      a.rsum_s[base + j] = rs;
      if (take) {
        a.qs[base + j] = qn;
        a.gs[base + j] = gn;
      }
      a.qe[base + j] = qn;
      a.re[base + j] = rn;
      a.ge[base + j] = gn;
      if (!odd) {
        ckr[(size_t)hi * P + j] = rn;
        cks[(size_t)hi * P + j] = rs;
      }
    }
    if (take) lps = lpn;
    logw_s = lw_new;
    sum_acc += acc_n;
    n_leaf += T(1);
    const bool div_s = (fl & NA_DIV_S) || div_n;
    const bool turn_s = (fl & NA_TURN_S) || turning;
    fl = (fl & ~(NA_DIV_S | NA_TURN_S)) | (div_s ? NA_DIV_S : 0) | (turn_s ? NA_TURN_S : 0);
    il += 1;
    if (div_s || turn_s || il == nlv) {
      // ---- the subtree is complete: merge it into the tree -----------------------------------------------
      const bool ok = !turn_s && !div_s;
      bool turn_full = false;
      if (ok) {
        const bool take2 = log(u_merge) < logw_s - log_w;
        log_w = na_logaddexp(log_w, logw_s);
        if (take2) lpp = lps;
        T dl = T(0), drr = T(0);
        for (int j = lane; j < P; j += 32) {
          if (take2) {
            a.qp[base + j] = a.qs[base + j];
            a.gp[base + j] = a.gs[base + j];
          }
          const T qe = a.qe[base + j], re = a.re[base + j], ge = a.ge[base + j];
          T rl, rr;
          if (dr > 0) {
            a.qR[base + j] = qe; a.rR[base + j] = re; a.gR[base + j] = ge;
            rr = re;
            rl = a.rL[base + j];
          } else {
            a.qL[base + j] = qe; a.rL[base + j] = re; a.gL[base + j] = ge;
            rl = re;
            rr = a.rR[base + j];
          }
          const T rsm = a.rsum[base + j] + a.rsum_s[base + j];
          a.rsum[base + j] = rsm;
          const T rho = rsm - T(0.5) * (rr + rl);
          const T im = a.inv_mass[base + j];
          dl += im * rl * rho;
          drr += im * rr * rho;
        }
        dl = na_warp_sum(dl);
        drr = na_warp_sum(drr);
        turn_full = (dl <= T(0)) || (drr <= T(0));
      }
      if (div_s) fl |= NA_DIVERGED;
      if (turn_s || (ok && turn_full)) fl |= NA_TURNED;
      dep += 1;
      const bool alive = ok && !turn_full;
      if (alive && dep < a.maxd) {
        new_sub = true;
      } else {
        // ---- the transition is complete ----------------------------------------------------------------------
        for (int j = lane; j < P; j += 32) {
          a.q[base + j] = a.qp[base + j];
          a.g[base + j] = a.gp[base + j];
        }
        lp = lpp;
        const T acc = sum_acc / fmax(n_leaf, T(1));
        if (itc < a.W) {
          // dual averaging (goose/da.py:69-127) and the window sums of the mass-matrix tuner
          const T tt = (T)(tep + 1);
          const T eta = pow(tt, -a.kappa);
          da_err += a.target - acc;
          const T log_step = da_mu - (da_err * sqrt(tt)) / (a.gamma * (a.t0 + tt));
          stp = exp(log_step);
          da_logavg = (T(1) - eta) * da_logavg + eta * log_step;
          for (int j = lane; j < P; j += 32) {
            const double qd = (double)a.q[base + j];
            a.ws1[base + j] += qd;
            a.ws2[base + j] += qd * qd;
          }
          tep += 1;
          const int dur = a.ep_dur[epi], kind = a.ep_kind[epi];
          if (tep == dur) {  // the epoch ends (goose/nuts.py:306-362)
            stp = exp(da_logavg);
            if (kind == 1 && dur > 1) {
              T so = T(0), sn = T(0);
              for (int j = lane; j < P; j += 32) {
                const double s1 = a.ws1[base + j], s2 = a.ws2[base + j];
                const double var = (s2 - s1 * s1 / (double)dur) / (double)(dur - 1);
                const T ni = (T)(var + 0.001);
                so += a.inv_mass[base + j];
                sn += ni;
                a.inv_mass[base + j] = ni;
              }
              so = na_warp_sum(so);
              sn = na_warp_sum(sn);
              stp *= sqrt(so / sn);
            }
            for (int j = lane; j < P; j += 32) {
              a.ws1[base + j] = 0.0;
              a.ws2[base + j] = 0.0;
            }
            epi += 1;
            tep = 0;
            da_err = T(0);
            da_logavg = log(stp);
            da_mu = log(T(10) * stp);
          }
        } else if (a.S > 0) {
          T* out = a.draws + ((size_t)(itc - a.W) * a.C + c) * P;
          for (int j = lane; j < P; j += 32) out[j] = a.q[base + j];
          if (lane == 0) {
            double* st = a.stats + (size_t)c * 4;
            st[0] += (double)acc;
            st[1] += (double)dep;
            st[2] += (fl & NA_DIVERGED) ? 1.0 : 0.0;
            st[3] += (double)n_leaf;
          }
        }
        itc += 1;
        if (a.nb > 0) {
          // Gibbs update of the smoothing variances: tau2 = (b0 + q' M q / 2) / Gamma(a_post, 1)
          for (int b = 0; b < a.nb; ++b) {
            const T* M = a.gib_M + (size_t)b * P * P;
            T part = T(0);
            for (int j = lane; j < P; j += 32) {
              T v = T(0);
              for (int i = 0; i < P; ++i) v += a.q[base + i] * M[(size_t)i * P + j];
              part += v * a.q[base + j];
            }
            const double quad = (double)na_warp_sum(part);
            const double tau2 = (a.gib_par[b * 3 + 1] + 0.5 * quad) / a.GAM[(size_t)c * a.nb + b];
            if (lane == 0) a.aux[(size_t)c * a.A + (int)a.gib_par[b * 3 + 2]] = (T)tau2;
          }
        }
        if (itc == a.W + a.S) {  // the stage is complete for this chain (a.S = 0 during the warm-up stage)
          ph = NA_DONE;
          if (lane == 0) atomicAdd(a.n_done, 1);
        } else if (a.nb > 0) {
          ph = NA_REFRESH;
        } else {
          start_tr = true;
        }
      }
    }
  }

  if (start_tr) {
    // ---- a transition starts: fresh momentum, the tree is the current point --------------------------------
    T kin = T(0);
    for (int j = lane; j < P; j += 32) {
      const T im = a.inv_mass[base + j];
      const T r0 = a.Z[base + j] / sqrt(im);
      const T qv = a.q[base + j], gv = a.g[base + j];
      a.qL[base + j] = qv; a.qR[base + j] = qv; a.qp[base + j] = qv;
      a.gL[base + j] = gv; a.gR[base + j] = gv; a.gp[base + j] = gv;
      a.rL[base + j] = r0; a.rR[base + j] = r0; a.rsum[base + j] = r0;
      kin += im * r0 * r0;
    }
    kin = T(0.5) * na_warp_sum(kin);
    h0 = -lp + kin;
    lpp = lp;
    log_w = T(0);
    sum_acc = T(0);
    n_leaf = T(0);
    dep = 0;
    fl &= ~(NA_DIVERGED | NA_TURNED);
    new_sub = true;
  }
  if (new_sub) {
    // ---- a subtree of 2^depth leaves starts at the left or right edge ----------------------------------------
    dr = u_dir < T(0.5) ? 1 : -1;
    for (int j = lane; j < P; j += 32) {
      const T qe = dr > 0 ? a.qR[base + j] : a.qL[base + j];
      const T re = dr > 0 ? a.rR[base + j] : a.rL[base + j];
      const T ge = dr > 0 ? a.gR[base + j] : a.gL[base + j];
      a.qe[base + j] = qe; a.re[base + j] = re; a.ge[base + j] = ge;
      a.qs[base + j] = qe; a.gs[base + j] = ge;
      a.rsum_s[base + j] = T(0);
    }
    lps = lpp;
    logw_s = NINF;
    fl &= ~(NA_TURN_S | NA_DIV_S);
    il = 0;
    nlv = 1 << dep;
    ph = NA_LEAF;
  }
  // ---- where the next evaluation happens -------------------------------------------------------------------------
  if (ph == NA_LEAF) {
    const T e = (T)dr * stp;
    for (int j = lane; j < P; j += 32) {
      const T rh = a.re[base + j] + T(0.5) * e * a.ge[base + j];
      a.r_half[base + j] = rh;
      a.qn[base + j] = a.qe[base + j] + e * (a.inv_mass[base + j] * rh);
    }
  } else {
    for (int j = lane; j < P; j += 32) a.qn[base + j] = a.q[base + j];
  }
  if (lane == 0) {
    a.phase[c] = ph; a.flags[c] = fl; a.depth[c] = dep; a.dir[c] = dr; a.i_leaf[c] = il; a.n_leaves[c] = nlv;
    a.it[c] = itc; a.ep[c] = epi; a.t_ep[c] = tep;
    a.step[c] = stp; a.lp[c] = lp; a.lpp[c] = lpp; a.lps[c] = lps; a.logw_s[c] = logw_s; a.log_w[c] = log_w;
    a.sum_acc[c] = sum_acc; a.n_leaf[c] = n_leaf; a.h0[c] = h0;
    a.da_err[c] = da_err; a.da_logavg[c] = da_logavg; a.da_mu[c] = da_mu;
  }
}

}  // namespace lslb

// ptrs: the pointer members of NutsAsync in declaration order (58 entries);
// ints = {C, P, D1, max_treedepth, warmup, samples, n_epochs, n_gibbs_blocks, A};
// dbls = {divergence threshold, target accept, gamma, kappa, t0}
template <typename T>
static int na_launch(void* const* p, const int* ints, const double* dbls, cudaStream_t s) {
  if (!p || !ints || !dbls) return LSLB_ERR_INVALID;
  lslb::NutsAsync<T> a;
  a.C = ints[0]; a.P = ints[1]; a.D1 = ints[2]; a.maxd = ints[3]; a.W = ints[4]; a.S = ints[5]; a.nep = ints[6];
  a.nb = ints[7]; a.A = ints[8];
  a.thr = (T)dbls[0]; a.target = (T)dbls[1]; a.gamma = (T)dbls[2]; a.kappa = (T)dbls[3]; a.t0 = (T)dbls[4];
  if (a.C < 1 || a.P < 1 || a.D1 < 2 || a.maxd < 1 || a.maxd >= a.D1 + 1) return LSLB_ERR_INVALID;
  int i = 0;
  T** tp[] = {&a.q, &a.lp, &a.g, &a.step, &a.inv_mass, &a.h0, &a.qL, &a.rL, &a.gL, &a.qR, &a.rR, &a.gR, &a.qp, &a.lpp,
              &a.gp, &a.log_w, &a.rsum, &a.sum_acc, &a.n_leaf, &a.qe, &a.re, &a.ge, &a.qs, &a.lps, &a.gs, &a.logw_s,
              &a.rsum_s, &a.ck_r, &a.ck_s, &a.qn, &a.r_half};
  for (T** t : tp) *t = (T*)p[i++];
  a.lp_n = (const T*)p[i++]; a.g_n = (const T*)p[i++]; a.Z = (const T*)p[i++]; a.U = (const T*)p[i++];
  a.GAM = (const double*)p[i++];
  a.da_err = (T*)p[i++]; a.da_logavg = (T*)p[i++]; a.da_mu = (T*)p[i++];
  a.ws1 = (double*)p[i++]; a.ws2 = (double*)p[i++];
  a.draws = (T*)p[i++]; a.aux = (T*)p[i++];
  a.gib_M = (const T*)p[i++]; a.gib_par = (const double*)p[i++];
  int** ip[] = {&a.phase, &a.depth, &a.dir, &a.i_leaf, &a.n_leaves, &a.it, &a.ep, &a.t_ep, &a.flags};
  for (int** t : ip) *t = (int*)p[i++];
  a.ep_dur = (const int*)p[i++]; a.ep_kind = (const int*)p[i++];
  a.stats = (double*)p[i++]; a.n_done = (int*)p[i++];
  lslb::nuts_async_step_kernel<T><<<(a.C + 3) / 4, 128, 0, s>>>(a);
  LSLB_LAUNCH_CHECK();
  return LSLB_OK;
}

extern "C" int lslb_nuts_async_step_f32(void* const* ptrs, const int* ints, const double* dbls, lslb_stream_t s) {
  return na_launch<float>(ptrs, ints, dbls, (cudaStream_t)s);
}
extern "C" int lslb_nuts_async_step_f64(void* const* ptrs, const int* ints, const double* dbls, lslb_stream_t s) {
  return na_launch<double>(ptrs, ints, dbls, (cudaStream_t)s);
}
