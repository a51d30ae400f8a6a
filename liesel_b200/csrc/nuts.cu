// Fused per-leaf bookkeeping of the batched NUTS driver (liesel_b200/nuts.py): the ~30 element-wise
// and reduction steps between two batched log-posterior+gradient evaluations, as two kernels.
//   pre : half-step momentum + position update        (q_n = q_e + eps * M^-1 (r_e + eps/2 g_e))
//   post: second half-step, energy, divergence test, multinomial update of the subtree proposal,
//         momentum sums, checkpointed generalised U-turn test, edge update, activity mask.
// One warp per chain; every reduction is over the chain's own P coefficients (fixed order).
// Semantics are those of BatchedNUTS.transition (torch path), which follows the reference's
// NUTSKernel / BlackJAX iterative NUTS (src/liesel/goose/nuts.py:240-264).
#include "common.cuh"

namespace lslb {

template <typename T>
struct NutsLeaf {
  int C, P, D1;
  T *q_e, *r_e, *g_e;           // [C,P] moving edge of the subtree
  T *q_n, *r_half;              // [C,P] proposal position (input of the evaluation), half-step momentum
  const T *g_n, *lp_n;          // [C,P], [C] evaluation results at q_n
  const T *eps, *inv_mass, *h0; // [C] signed step, [C,P], [C]
  T *q_s, *g_s, *lp_s;          // subtree proposal
  T *logw_s, *rsum_s;           // [C], [C,P]
  T *ck_r, *ck_s;               // [C,D1,P] checkpoints
  T *sum_acc, *n_leaf;          // [C]
  unsigned char *act, *turn_s, *div_s;  // [C]
  const T* u;                   // [C] uniforms for this leaf
  int odd, lo, hi;
  T div_thr;
};

template <typename T>
__global__ void nuts_pre_kernel(const __grid_constant__ NutsLeaf<T> a) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)a.C * a.P) return;
  const int c = (int)(idx / a.P);
  const T e = a.eps[c];
  const T rh = a.r_e[idx] + T(0.5) * e * a.g_e[idx];
  a.r_half[idx] = rh;
  a.q_n[idx] = a.q_e[idx] + e * (a.inv_mass[idx] * rh);
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T>
__global__ void nuts_post_kernel(const __grid_constant__ NutsLeaf<T> a) {
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= a.C) return;
  const size_t base = (size_t)c * a.P;
  const T e = a.eps[c];
  const bool act = a.act[c] != 0;
  // second half step and kinetic energy (computed for every chain; masked chains discard it)
  T kin = T(0);
  for (int j = lane; j < a.P; j += 32) {
    const T rn = a.r_half[base + j] + T(0.5) * e * a.g_n[base + j];
    a.r_half[base + j] = rn;  // r_half now holds r_n
    kin += a.inv_mass[base + j] * rn * rn;
  }
  kin = T(0.5) * warp_sum(kin);
  if (!act) return;  // warp-uniform
  const T lp = a.lp_n[c];
  T dh = (-lp + kin) - a.h0[c];
  if (dh != dh) dh = (T)INFINITY;
  const bool div = dh > a.div_thr;
  const T logw_n = -dh;
  const T acc = exp(fmin(-dh, T(0)));
  const T lw_old = a.logw_s[c];
  // logaddexp
  const T mx = fmax(lw_old, logw_n), mn = fmin(lw_old, logw_n);
  const T lw_new = (mx == (T)(-INFINITY)) ? (T)(-INFINITY) : mx + log1p(exp(mn - mx));
  const bool take = log(a.u[c]) < logw_n - lw_new;  // false when both weights are -inf (NaN)
  // turning test against the checkpoints touched by this leaf (odd leaves only)
  bool turning = false;
  T* ckr = a.ck_r + (size_t)c * a.D1 * a.P;
  T* cks = a.ck_s + (size_t)c * a.D1 * a.P;
  if (a.odd) {
    for (int k = a.hi; k >= a.lo; --k) {
      T dl = T(0), dr = T(0);
      for (int j = lane; j < a.P; j += 32) {
        const T rn = a.r_half[base + j];
        const T rs = a.rsum_s[base + j] + rn;  // subtree momentum sum including this leaf
        const T rl = ckr[(size_t)k * a.P + j];
        const T sub = rs - cks[(size_t)k * a.P + j] + rl;
        const T rho = sub - T(0.5) * (rn + rl);
        const T im = a.inv_mass[base + j];
        dl += im * rl * rho;
        dr += im * rn * rho;
      }
      dl = warp_sum(dl);
      dr = warp_sum(dr);
      turning = turning || (dl <= T(0)) || (dr <= T(0));
    }
  }
  for (int j = lane; j < a.P; j += 32) {
    const T rn = a.r_half[base + j];
    const T qn = a.q_n[base + j], gn = a.g_n[base + j];
    const T rs = a.rsum_s[base + j] + rn;
    a.rsum_s[base + j] = rs;
    if (take) {
      a.q_s[base + j] = qn;
      a.g_s[base + j] = gn;
    }
    a.q_e[base + j] = qn;
    a.r_e[base + j] = rn;
    a.g_e[base + j] = gn;
    if (!a.odd) {
      ckr[(size_t)a.hi * a.P + j] = rn;
      cks[(size_t)a.hi * a.P + j] = rs;
    }
  }
  if (lane == 0) {
    if (take) a.lp_s[c] = lp;
    a.logw_s[c] = lw_new;
    a.sum_acc[c] += acc;
    a.n_leaf[c] += T(1);
    const bool d = a.div_s[c] || div;
    const bool t = a.turn_s[c] || turning;
    a.div_s[c] = d;
    a.turn_s[c] = t;
    a.act[c] = !(d || t);
  }
}

template <typename T>
int nuts_pre(const NutsLeaf<T>& a, cudaStream_t s) {
  long long total = (long long)a.C * a.P;
  nuts_pre_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(a);
  LSLB_LAUNCH_CHECK();
  return LSLB_OK;
}
template <typename T>
int nuts_post(const NutsLeaf<T>& a, cudaStream_t s) {
  nuts_post_kernel<T><<<(a.C + 3) / 4, 128, 0, s>>>(a);
  LSLB_LAUNCH_CHECK();
  return LSLB_OK;
}

}  // namespace lslb

// The argument block is passed as an array of pointers + ints so the C ABI stays plain:
//   ptrs[0..21] in the order of NutsLeaf's pointer members; ints = {C, P, D1, odd, lo, hi}
template <typename T>
static int fill(lslb::NutsLeaf<T>& a, void* const* ptrs, const int* ints, double div_thr) {
  if (!ptrs || !ints) return LSLB_ERR_INVALID;
  a.C = ints[0]; a.P = ints[1]; a.D1 = ints[2]; a.odd = ints[3]; a.lo = ints[4]; a.hi = ints[5];
  int i = 0;
  a.q_e = (T*)ptrs[i++]; a.r_e = (T*)ptrs[i++]; a.g_e = (T*)ptrs[i++];
  a.q_n = (T*)ptrs[i++]; a.r_half = (T*)ptrs[i++];
  a.g_n = (const T*)ptrs[i++]; a.lp_n = (const T*)ptrs[i++];
  a.eps = (const T*)ptrs[i++]; a.inv_mass = (const T*)ptrs[i++]; a.h0 = (const T*)ptrs[i++];
  a.q_s = (T*)ptrs[i++]; a.g_s = (T*)ptrs[i++]; a.lp_s = (T*)ptrs[i++];
  a.logw_s = (T*)ptrs[i++]; a.rsum_s = (T*)ptrs[i++];
  a.ck_r = (T*)ptrs[i++]; a.ck_s = (T*)ptrs[i++];
  a.sum_acc = (T*)ptrs[i++]; a.n_leaf = (T*)ptrs[i++];
  a.act = (unsigned char*)ptrs[i++]; a.turn_s = (unsigned char*)ptrs[i++]; a.div_s = (unsigned char*)ptrs[i++];
  a.u = (const T*)ptrs[i++];
  a.div_thr = (T)div_thr;
  if (a.C < 1 || a.P < 1 || a.D1 < 1) return LSLB_ERR_INVALID;
  return LSLB_OK;
}

extern "C" int lslb_nuts_leaf_pre_f32(void* const* ptrs, const int* ints, lslb_stream_t s) {
  lslb::NutsLeaf<float> a;
  int st = fill(a, ptrs, ints, 0.0);
  return st ? st : lslb::nuts_pre<float>(a, (cudaStream_t)s);
}
extern "C" int lslb_nuts_leaf_pre_f64(void* const* ptrs, const int* ints, lslb_stream_t s) {
  lslb::NutsLeaf<double> a;
  int st = fill(a, ptrs, ints, 0.0);
  return st ? st : lslb::nuts_pre<double>(a, (cudaStream_t)s);
}
extern "C" int lslb_nuts_leaf_post_f32(void* const* ptrs, const int* ints, double div_thr, lslb_stream_t s) {
  lslb::NutsLeaf<float> a;
  int st = fill(a, ptrs, ints, div_thr);
  return st ? st : lslb::nuts_post<float>(a, (cudaStream_t)s);
}
extern "C" int lslb_nuts_leaf_post_f64(void* const* ptrs, const int* ints, double div_thr, lslb_stream_t s) {
  lslb::NutsLeaf<double> a;
  int st = fill(a, ptrs, ints, div_thr);
  return st ? st : lslb::nuts_post<double>(a, (cudaStream_t)s);
}
