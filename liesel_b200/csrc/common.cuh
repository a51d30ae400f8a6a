// Shared internals of libliesel_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "liesel_b200.h"

#define LSLB_MAX_SPL 8
#define LSLB_MAX_DNS 4
#define LSLB_MAX_INT 6  // 4 intercept blocks + one virtual row per predictor (centred splines)
#define LSLB_NACC 2  // extra accumulator rows per chunk: logp (hi, lo)

namespace lslb {

// ---------------------------------------------------------------------------------------
// error plumbing: compute calls never abort, they return a status code
// ---------------------------------------------------------------------------------------
void set_cuda_error(cudaError_t e, const char* where);
void set_comm_error(const char* msg);  // reported by lslb_last_cuda_error() as well
void debug_record_start(cudaStream_t s);  // debug.cu: optional events around the dominant kernel
void debug_record_stop(cudaStream_t s);
#define LSLB_CUDA(call)                                  \
  do {                                                   \
    cudaError_t _e = (call);                             \
    if (_e != cudaSuccess) {                             \
      ::lslb::set_cuda_error(_e, #call);                 \
      return LSLB_ERR_CUDA;                              \
    }                                                    \
  } while (0)
#define LSLB_LAUNCH_CHECK()                              \
  do {                                                   \
    cudaError_t _e = cudaGetLastError();                 \
    if (_e != cudaSuccess) {                             \
      ::lslb::set_cuda_error(_e, "kernel launch");       \
      return LSLB_ERR_CUDA;                              \
    }                                                    \
  } while (0)

// ---------------------------------------------------------------------------------------
// kernel-side model description (passed by value as a __grid_constant__ parameter)
// ---------------------------------------------------------------------------------------
template <typename T>
struct SplP {
  const T* w;        // [n][4]
  const int* start;  // [n]
  int soff;          // offset in the small-parameter space
  int k;
  int pred;
  int pad;
  const float* tw;   // banded_x2 only: per 128-row tile the column sums of w, [ntiles][4] (plan-owned)
  const float* wp;   // banded_iwls_x2 only: per row the ten products w_a w_b (a <= b), [n][12] (plan-owned) or NULL
};
template <typename T>
struct DnsP {
  const T* X;
  long long ld;
  int soff;
  int p;
  int pred;
  int pad;
};
struct IntP {
  int soff;
  int pred;
};
// prior bookkeeping for the finalize kernels (one entry per block, incl. the group block)
struct PriorP {
  int soff;   // small-space offset (-1 for the group block)
  int off;    // offset in the flat [P] layout
  int ncoef;
  int prior;  // lslb_prior
  double m, inv_s2, log_s;
  const void* K;
  double rank, log_pdet, tau2_fixed, ig_a, ig_b, ig_const;  // ig_const = a log b - lgamma(a)
  int tau2_slot;
  int has_ig;
  int vrow;            // small-space row of the predictor's centring term (-1: block not centred)
  const void* center;  // column means of a centred spline block
};

template <typename T>
struct EvalParams {
  long long n;
  const T* y;
  T fixed_inv_scale;
  int ns, nd, ni;
  SplP<T> sp[LSLB_MAX_SPL];
  DnsP<T> dn[LSLB_MAX_DNS];
  IntP ic[LSLB_MAX_INT];
  // random intercept
  const int* gidx;
  int gpred;
  int G;
  const T* bT;  // [G][Cpad]
  T* gbT;       // [G][Cpad]  (pre-initialised with the prior gradient)
  // chunking
  const long long* chunk_start;  // [nchunks + 1]
  int nchunks;
  int C, Cpad, Psmall;
  const T* pT;  // [Psmall][Cpad]
  T* partial;   // [nchunks][Psmall + LSLB_NACC][Cpad]
};

struct FinalizeParams {
  int nblocks;
  PriorP pr[LSLB_MAX_BLOCKS];
  int P, A, Psmall, C, Cpad, nchunks;
  int nchunks_logp;  // slices whose log-likelihood rows are valid (<= nchunks): the tensor-core path's
                     // K1 runs one CTA per SM over long row ranges, its K2 many short split-K chunks
  int Preal;  // rows [Preal, Psmall) of the small space are virtual (centring terms)
  double lik_const, const_term;
  unsigned flags;
};

// ---------------------------------------------------------------------------------------
// math helpers: float uses the SFU approximations (ex2/lg2/rcp), double the libm versions
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float fexp(float x) { return __expf(x); }
__device__ __forceinline__ double fexp(double x) { return exp(x); }
__device__ __forceinline__ float flog1p_pos(float e) { return __logf(1.0f + e); }  // e in (0,1]
__device__ __forceinline__ double flog1p_pos(double e) { return log1p(e); }
__device__ __forceinline__ float frcp(float x) { return __fdividef(1.0f, x); }  // MUFU.RCP
__device__ __forceinline__ double frcp(double x) { return 1.0 / x; }

// internal family codes: the public enum plus "Normal with a constant scale"
enum { FAM_NORMAL_LS = 0, FAM_BERNOULLI = 1, FAM_POISSON = 2, FAM_NORMAL_FIXED = 3 };

// Per-row response term. ll omits row-independent constants (added once in finalize):
//   Normal:   -0.5*log(2*pi) [- log(fixed scale)]      Poisson: -lgamma(y+1)
// r0/r1 = d ll / d eta0/eta1;  w0/w1 = -d2 ll / d eta^2 (observed information).
template <typename T, int FAM, bool WANT_W>
__device__ __forceinline__ void lik_row(T y, T eta0, T eta1, T fixed_inv_scale, T& ll, T& r0,
                                        T& r1, T& w0, T& w1) {
  if constexpr (FAM == FAM_NORMAL_LS) {
    T inv_s = fexp(-eta1);
    T z = (y - eta0) * inv_s;
    T z2 = z * z;
    ll = T(-0.5) * z2 - eta1;
    r0 = z * inv_s;
    r1 = z2 - T(1);
    if constexpr (WANT_W) {
      w0 = inv_s * inv_s;
      w1 = T(2) * z2;
    }
  } else if constexpr (FAM == FAM_NORMAL_FIXED) {
    T z = (y - eta0) * fixed_inv_scale;
    ll = T(-0.5) * z * z;
    r0 = z * fixed_inv_scale;
    r1 = T(0);
    if constexpr (WANT_W) {
      w0 = fixed_inv_scale * fixed_inv_scale;
      w1 = T(0);
    }
  } else if constexpr (FAM == FAM_BERNOULLI) {
    T a = fabs(eta0);
    T e = fexp(-a);
    T inv = frcp(T(1) + e);
    T pi = eta0 >= T(0) ? inv : e * inv;
    ll = y * eta0 - (fmax(eta0, T(0)) + flog1p_pos(e));
    r0 = y - pi;
    r1 = T(0);
    if constexpr (WANT_W) {
      w0 = pi * (T(1) - pi);
      w1 = T(0);
    }
  } else {  // FAM_POISSON
    T mu = fexp(eta0);
    ll = y * eta0 - mu;
    r0 = y - mu;
    r1 = T(0);
    if constexpr (WANT_W) {
      w0 = mu;
      w1 = T(0);
    }
  }
}

// Kahan-compensated add of a (small) tile sum into a running (hi, lo) pair.
template <typename T>
__device__ __forceinline__ void kahan_add(T& hi, T& lo, T x) {
  T yv = x - lo;
  T t = hi + yv;
  lo = (t - hi) - yv;
  hi = t;
}

template <typename T>
struct alignas(16) Vec4 {
  T a, b, c, d;
};
__device__ __forceinline__ Vec4<float> ld4(const float* p) {
  float4 v = __ldg(reinterpret_cast<const float4*>(p));
  return {v.x, v.y, v.z, v.w};
}
__device__ __forceinline__ Vec4<double> ld4(const double* p) {
  double2 u = __ldg(reinterpret_cast<const double2*>(p));
  double2 v = __ldg(reinterpret_cast<const double2*>(p) + 1);
  return {u.x, u.y, v.x, v.y};
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace lslb

// ---------------------------------------------------------------------------------------
// host-side plan
// ---------------------------------------------------------------------------------------
struct lslb_plan {
  lslb_model_desc desc;
  lslb_block_desc blk[LSLB_MAX_BLOCKS];
  int soff[LSLB_MAX_BLOCKS];  // small-space offsets (-1 for the group block)
  int off[LSLB_MAX_BLOCKS];   // flat offsets
  int P, A, Psmall;  // Psmall counts the virtual centring rows; Preal does not
  int Preal;
  int vrow[2];       // per predictor: small-space row holding -sum_b cbar_b^T beta_b, or -1
  int ns, nd, ni;
  int sp[LSLB_MAX_SPL], dn[LSLB_MAX_DNS], ic[LSLB_MAX_INT];
  int grp;  // index of the group block or -1
  bool has_scale;
  int fam;  // internal family code
  int sm_count;
  int nchunks;
  long long* d_chunk_start;  // device [nchunks+1]
  int* d_gstart;             // random intercept: device [G + 1], first row of every group (rows sorted by group) or NULL
  double lik_const;
  bool rows_sorted;  // every smooth's knot span changes rarely from row to row
  bool unit_rows;    // every spline row's four weights sum to one (x inside the knot range)
  bool shared_band;  // exactly two smooths whose spans and weights are identical bit for bit (same covariate, same knots)
  float* d_XT;       // tensor-core path (or NULL): plan-owned hi part of the transposed dense block [p x n_pad]
  float *d_XTl, *d_Xh, *d_Xl;  // its lo part; hi/lo parts of X [n x p] (hi + lo == x exactly)
  float* d_tw[LSLB_MAX_SPL];   // float32 banded plans: per 128-row tile the column sums of each smooth's weights
  float* d_wp[LSLB_MAX_SPL];   // float32 banded Normal plans: per row the products w_a w_b of each smooth ([n][12])
  // chunk-sorted multi-smooth path (msort_x2.cu): per smooth the weights and local rows of every sub-chunk
  // in span-sorted order, the run offsets per (sub-chunk, smooth), the sub-chunk boundaries
  float* d_ms_w[LSLB_MAX_SPL];
  unsigned short* d_ms_idx[LSLB_MAX_SPL];
  int* d_ms_off;
  long long* d_ms_start;
  int ms_nsub;  // 0: path not prepared
  long long n_pad;
  bool tc_two_pass;  // tensor-core path: two TF32 passes per product are enough for this data (plan-time check)
  int variant;
  const char* variant_name;
};

namespace lslb {
// kernel-variant ids (plan->variant)
enum {
  VAR_GENERIC = 0,   // any structure: coefficients and gradients in shared memory
  VAR_BANDED = 1,    // intercepts + splines only (S2/S3/H): run-length register kernel
  VAR_DENSE_REG = 2, // one dense block with p <= 16 in registers (+ optional group) (S1/S5)
};

template <typename T>
int launch_logp_grad(const lslb_plan* plan, int C, const T* params, const T* aux, T* logp,
                     T* grad, unsigned flags, void* ws, size_t ws_bytes, cudaStream_t stream);
template <typename T>
int launch_iwls(const lslb_plan* plan, int C, int block, const T* params, const T* aux, T* score,
                T* info, T* chol, unsigned flags, void* ws, size_t ws_bytes, cudaStream_t stream);
template <typename T>
int launch_chol(int C, int p, const T* info, T* chol, cudaStream_t stream);
size_t workspace_bytes(const lslb_plan* plan, int C);
struct LaunchCfg {
  int TC, Cpad, ntiles, fstride, nchunks;
};
LaunchCfg choose_cfg(const lslb_plan* plan, int C, size_t smem_per_chain, size_t smem_fixed = 0);
struct BandedCfg {
  int TC, Cpad, ntiles_chain, ntiles_total, tiles_per_cta, nchunks;
  int tc_eff = 0;  // banded_x2: chains a CTA really owns (multiple of 64, <= TC); 0 = TC
  int tiles_extra = 0;  // the first tiles_extra row chunks take tiles_per_cta + 1 tiles (balanced partition)
  int stages = 6;  // banded_x2: depth of the TMA ring (6; 3 / 2 for the 2- / 1-warp CTAs of small chain counts)
  int phases = 1;  // banded_x2 / banded_iwls_x2, phased CTAs: row phases per CTA; nchunks counts partial slots = CTAs x phases
};
struct TcCfg {
  int Cpad, ntiles_chain, ntiles_rows, nkchunks, tiles_per_cta, kchunks_per_cta, nchunks;
  int two_pass;  // bit 0 / bit 1: K1 / K2 drop the product with a lo factor of the large operand
  int tiles_per_cta1, nchunks1;  // K1's own (coarse) row chunking: its accumulations are per tile
};
int dense_tc_prepare(lslb_plan* plan);                 // dense_tc.cu (called by lslb_plan_create)
void dense_tc_release(lslb_plan* plan);
bool dense_tc_eligible(const lslb_plan* plan, int C);
TcCfg dense_tc_cfg(const lslb_plan* plan, int C);
size_t dense_tc_rt_floats(const lslb_plan* plan, int C);
int launch_dense_tc(const lslb_plan* plan, const TcCfg& k, int C, const float* params, float* RT,
                    float* partial, bool want_grad, cudaStream_t st);
bool tiny_eligible(const lslb_plan* plan, int C);  // tiny.cu: small regressions, one launch per call
template <typename T>
int launch_tiny(const lslb_plan* plan, const EvalParams<T>& ep, const FinalizeParams& fp, int C, const T* params,
                T* logp, T* grad, cudaStream_t st);
bool dgroup_x2_eligible(const lslb_plan* plan);  // dgroup_x2.cu
LaunchCfg dgroup_x2_cfg(const lslb_plan* plan, int C);
int launch_dgroup_x2(const lslb_plan* plan, const EvalParams<float>& ep, const LaunchCfg& k,
                     bool want_grad, cudaStream_t st);
bool dense_x2_eligible(const lslb_plan* plan);  // dense_x2.cu
BandedCfg dense_x2_cfg(const lslb_plan* plan, int C);
int launch_dense_x2(const lslb_plan* plan, const EvalParams<float>& ep, const BandedCfg& k,
                    bool want_grad, cudaStream_t st);
int msort_prepare(lslb_plan* plan);  // msort_x2.cu (called by lslb_plan_create)
void msort_release(lslb_plan* plan);
bool msort_x2_eligible(const lslb_plan* plan);
LaunchCfg msort_x2_cfg(const lslb_plan* plan, int C);
int launch_msort_x2(const lslb_plan* plan, const EvalParams<float>& ep, const LaunchCfg& k, bool g,
                    cudaStream_t st);
bool msmooth_x2_eligible(const lslb_plan* plan);  // msmooth_x2.cu
LaunchCfg msmooth_x2_cfg(const lslb_plan* plan, int C);
int launch_msmooth_x2(const lslb_plan* plan, const EvalParams<float>& ep, const LaunchCfg& k,
                      bool want_grad, cudaStream_t st);
bool banded_fast_eligible(const lslb_plan* plan);  // banded.cu
BandedCfg banded_cfg(const lslb_plan* plan, int C);
template <typename T>
int launch_banded(const lslb_plan* plan, const EvalParams<T>& ep, const BandedCfg& k, bool want_grad,
                  cudaStream_t st);
void fill_finalize_params(const lslb_plan* plan, FinalizeParams& fp, int C, int Cpad, int nchunks,
                          unsigned flags);
template <typename T>
void fill_eval_params(const lslb_plan* plan, EvalParams<T>& ep, int C, int Cpad);
}  // namespace lslb
