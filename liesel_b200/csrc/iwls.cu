// IWLS score + observed information of ONE coefficient block for C chains in a single pass
// over the rows, ridge + batched small Cholesky, the proposal tail and the Gibbs quadratic form.
//
// Replaces `grad(flat_log_prob_fn)` and `-jacfwd(grad(...))` of the reference
// (src/liesel/goose/iwls.py:315-321), `_chol_info` (:218-243), `solve`/`mvn_log_prob`/
// `mvn_sample` (src/liesel/goose/iwls_utils.py:14-70) and `beta @ K @ beta`
// (src/liesel/model/distreg.py:291).
#include "common.cuh"
#include "prologue.cuh"
#include "x2_common.cuh"

namespace lslb {

enum { TK_SPLINE = 0, TK_INTERCEPT = 1, TK_DENSE = 2 };

struct IwlsParams {
  int tkind;   // TK_*
  int tidx;    // index within ep.sp / ep.dn / ep.ic
  int p;       // coefficients of the target block
  int tpred;
  int nq;      // accumulator rows per chunk
  void* partial;  // [nchunks][nq][Cpad]
};

// accumulator rows: score [p] then
//   spline    : band [p][4]  entry (a, a+d) at p + 4a + d
//   dense     : packed upper triangle, (a<=b) at p + a*p - a(a-1)/2 + (b-a)
//   intercept : single entry at 1
__host__ __device__ inline int iwls_nq(int tkind, int p) {
  if (tkind == TK_SPLINE) return p + 4 * p;
  if (tkind == TK_INTERCEPT) return 2;
  return p + p * (p + 1) / 2;
}

template <typename T, int FAM, int NS, bool GROUP>
__global__ void __launch_bounds__(128)
iwls_kernel(const __grid_constant__ EvalParams<T> p, const __grid_constant__ IwlsParams q,
            int fstride) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int TC = blockDim.x, tid = threadIdx.x;
  const int c = blockIdx.y * TC + tid;
  T* bs = reinterpret_cast<T*>(smraw);
  T* as = bs + (size_t)p.Psmall * TC;  // accumulators [nq][TC]

  for (int j = 0; j < p.Psmall; ++j) bs[j * TC + tid] = p.pT[(size_t)j * p.Cpad + c];
  for (int j = 0; j < q.nq; ++j) as[j * TC + tid] = T(0);

  int f0 = blockIdx.x * fstride;
  int f1 = min(f0 + fstride, p.nchunks);
  const long long r0 = p.chunk_start[f0], r1 = p.chunk_start[f1];

  T o0 = T(0), o1 = T(0);
  for (int k = 0; k < p.ni; ++k) {
    T v = bs[p.ic[k].soff * TC + tid];
    if (p.ic[k].pred) o1 += v; else o0 += v;
  }
  int cur[NS > 0 ? NS : 1];
  T wb[NS > 0 ? NS : 1][4];
#pragma unroll
  for (int m = 0; m < NS; ++m) {
    cur[m] = -1;
#pragma unroll
    for (int t = 0; t < 4; ++t) wb[m][t] = T(0);
  }
  // target-spline working set: score[4] + upper triangle of the 4x4 outer product
  int tcur = -1;
  T ts[4] = {T(0), T(0), T(0), T(0)};
  T tg[10];
#pragma unroll
  for (int t = 0; t < 10; ++t) tg[t] = T(0);
  T is = T(0), ig = T(0);  // intercept target
  int curg = -1;
  T bg = T(0);

  auto flush_target = [&]() {
    if (tcur < 0) return;
#pragma unroll
    for (int a = 0; a < 4; ++a) as[(tcur + a) * TC + tid] += ts[a];
    int u = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = a; b < 4; ++b) {
        as[(q.p + 4 * (tcur + a) + (b - a)) * TC + tid] += tg[u];
        ++u;
      }
#pragma unroll
    for (int a = 0; a < 4; ++a) ts[a] = T(0);
#pragma unroll
    for (int t = 0; t < 10; ++t) tg[t] = T(0);
  };

  for (long long i = r0; i < r1; ++i) {
    T eta0 = o0, eta1 = o1;
    Vec4<T> wt = {T(0), T(0), T(0), T(0)};
    int st = 0;
#pragma unroll
    for (int m = 0; m < NS; ++m) {
      if (m < p.ns) {
        const int s = __ldg(p.sp[m].start + i);
        Vec4<T> w = ld4(p.sp[m].w + 4 * i);
        if (s != cur[m]) {
#pragma unroll
          for (int t = 0; t < 4; ++t) wb[m][t] = bs[(p.sp[m].soff + s + t) * TC + tid];
          cur[m] = s;
        }
        T v = w.a * wb[m][0];
        v = fma(w.b, wb[m][1], v);
        v = fma(w.c, wb[m][2], v);
        v = fma(w.d, wb[m][3], v);
        if (p.sp[m].pred) eta1 += v; else eta0 += v;
        if (q.tkind == TK_SPLINE && m == q.tidx) {
          wt = w;
          st = s;
        }
      }
    }
    for (int d = 0; d < p.nd; ++d) {
      const T* xrow = p.dn[d].X + i * p.dn[d].ld;
      const T* bcol = bs + (size_t)p.dn[d].soff * TC + tid;
      T v = T(0);
      for (int j = 0; j < p.dn[d].p; ++j) v = fma(__ldg(xrow + j), bcol[j * TC], v);
      if (p.dn[d].pred) eta1 += v; else eta0 += v;
    }
    if constexpr (GROUP) {
      const int g = __ldg(p.gidx + i);
      if (g != curg) {
        curg = g;
        bg = p.bT[(size_t)g * p.Cpad + c];
      }
      if (p.gpred) eta1 += bg; else eta0 += bg;
    }
    T ll, d0, d1, u0, u1;
    lik_row<T, FAM, true>(__ldg(p.y + i), eta0, eta1, p.fixed_inv_scale, ll, d0, d1, u0, u1);
    const T r = q.tpred ? d1 : d0;
    const T W = q.tpred ? u1 : u0;

    if (q.tkind == TK_SPLINE) {
      if (st != tcur) {
        flush_target();
        tcur = st;
      }
      const T wv[4] = {wt.a, wt.b, wt.c, wt.d};
      int u = 0;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        ts[a] = fma(wv[a], r, ts[a]);
        const T wa = wv[a] * W;
#pragma unroll
        for (int b = a; b < 4; ++b) {
          tg[u] = fma(wa, wv[b], tg[u]);
          ++u;
        }
      }
    } else if (q.tkind == TK_INTERCEPT) {
      is += r;
      ig += W;
    } else {
      const DnsP<T>& D = p.dn[q.tidx];
      const T* xrow = D.X + i * D.ld;
      T* tri = as + (size_t)q.p * TC + tid;
      for (int a = 0; a < q.p; ++a) {
        const T xa = __ldg(xrow + a);
        as[a * TC + tid] = fma(xa, r, as[a * TC + tid]);
        const T xw = xa * W;
        for (int b = a; b < q.p; ++b) {
          tri[0] = fma(xw, __ldg(xrow + b), tri[0]);
          tri += TC;
        }
      }
    }
  }
  if (q.tkind == TK_SPLINE) flush_target();
  if (q.tkind == TK_INTERCEPT) {
    as[tid] = is;
    as[TC + tid] = ig;
  }
  T* out = reinterpret_cast<T*>(q.partial) + (size_t)blockIdx.x * q.nq * p.Cpad + c;
  for (int j = 0; j < q.nq; ++j) out[(size_t)j * p.Cpad] = as[j * TC + tid];
}

// Dense target with p <= PD <= 8 columns in a "dense block in registers" plan (one dense block, intercepts,
// optional random intercept: S1 / S5): score [p] and the packed upper triangle [p (p + 1) / 2] of
// X^T W X accumulate in REGISTERS (44 per chain at p = 8), thread = chain, all lanes read the same row
// (broadcast loads). iwls_kernel keeps those accumulators in shared memory and pays ~150 shared-memory
// read-modify-writes per row and chain: 55 ms per pass at S5 with 256 chains (n = 5e6), where the IWLS sweep of
// that model spends its time. Same partial layout, so reduce / finalize / Cholesky are shared.
template <typename T, int FAM, int PD, bool GROUP>
__global__ void __launch_bounds__(128)
iwls_dense_reg_kernel(const __grid_constant__ EvalParams<T> p, const __grid_constant__ IwlsParams q, int fstride,
                      int vec /* rows are PD wide, 16-byte aligned: vector loads */) {
  const int TC = blockDim.x, tid = threadIdx.x;
  const int c = blockIdx.y * TC + tid;
  const DnsP<T>& D = p.dn[0];
  const size_t ldc = (size_t)p.Cpad;
  T beta[PD], sc[PD], tri[PD * (PD + 1) / 2];
#pragma unroll
  for (int j = 0; j < PD; ++j) {
    beta[j] = j < D.p ? p.pT[(size_t)(D.soff + j) * ldc + c] : T(0);
    sc[j] = T(0);
  }
#pragma unroll
  for (int j = 0; j < PD * (PD + 1) / 2; ++j) tri[j] = T(0);
  T o0 = T(0), o1 = T(0);
  for (int k = 0; k < p.ni; ++k) {
    const T v = p.pT[(size_t)p.ic[k].soff * ldc + c];
    if (p.ic[k].pred) o1 += v; else o0 += v;
  }
  const int f0 = blockIdx.x * fstride, f1 = min(f0 + fstride, p.nchunks);
  const long long r0 = p.chunk_start[f0], r1 = p.chunk_start[f1];
  int curg = -1;
  T bg = T(0);
  // one row of look-ahead in registers: the loads of row i + 1 are in flight while row i computes (without it
  // the warp waited out an L1/L2 round trip per row: long-scoreboard 3.2 stall cycles per issued instruction)
  auto load_row = [&](long long i, T (&xr)[PD], T& yr, int& gr) {
    const T* xrow = D.X + i * D.ld;
    if (vec) {  // (ten scalar loads per row and warp made the load unit the bottleneck: 2-3 vector loads)
#pragma unroll
      for (int j4 = 0; j4 < PD; j4 += 4) {
        const Vec4<T> v = ld4(xrow + j4);
        xr[j4] = v.a; xr[j4 + 1] = v.b; xr[j4 + 2] = v.c; xr[j4 + 3] = v.d;
      }
    } else {
#pragma unroll
      for (int j = 0; j < PD; ++j) xr[j] = j < D.p ? __ldg(xrow + j) : T(0);
    }
    yr = __ldg(p.y + i);
    gr = GROUP ? __ldg(p.gidx + i) : 0;
  };
  T xn[PD], yn = T(0);
  int gn = 0;
  if (r0 < r1) load_row(r0, xn, yn, gn);
  for (long long i = r0; i < r1; ++i) {
    T x[PD];
#pragma unroll
    for (int j = 0; j < PD; ++j) x[j] = xn[j];
    const T yv = yn;
    const int g = gn;
    if (i + 1 < r1) load_row(i + 1, xn, yn, gn);
    T dot = T(0);
#pragma unroll
    for (int j = 0; j < PD; ++j) dot = fma(x[j], beta[j], dot);
    T eta0 = o0, eta1 = o1;
    if (D.pred) eta1 += dot; else eta0 += dot;
    if constexpr (GROUP) {
      if (g != curg) {  // warp-uniform
        curg = g;
        bg = p.bT[(size_t)g * ldc + c];
        if (g + 1 < p.G) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.bT + (size_t)(g + 1) * ldc + c));
      }
      if (p.gpred) eta1 += bg; else eta0 += bg;
    }
    (void)g;
    T ll, d0, d1, u0, u1;
    lik_row<T, FAM, true>(yv, eta0, eta1, p.fixed_inv_scale, ll, d0, d1, u0, u1);
    const T r = q.tpred ? d1 : d0;
    const T W = q.tpred ? u1 : u0;
    int u = 0;
#pragma unroll
    for (int a = 0; a < PD; ++a) {
      sc[a] = fma(x[a], r, sc[a]);
      const T xw = x[a] * W;
#pragma unroll
      for (int b = a; b < PD; ++b) {
        tri[u] = fma(xw, x[b], tri[u]);
        ++u;
      }
    }
  }
  // accumulator rows as iwls_nq / iwls_kernel lay them out for p = q.p <= PD: score [p], then the packed
  // upper triangle row by row (the registers hold the PD-wide triangle; entries with b >= p are zero)
  T* out = reinterpret_cast<T*>(q.partial) + (size_t)blockIdx.x * q.nq * ldc + c;
#pragma unroll
  for (int a = 0; a < PD; ++a)
    if (a < q.p) out[(size_t)a * ldc] = sc[a];
  {
    int u = 0, row = q.p;
#pragma unroll
    for (int a = 0; a < PD; ++a) {
#pragma unroll
      for (int b = a; b < PD; ++b) {
        if (a < q.p && b < q.p) {
          out[(size_t)row * ldc] = tri[u];
          ++row;
        }
        ++u;
      }
    }
  }
}

// The same with TWO chains per thread and packed FP32 math (float32 plans whose rows are PD wide and 16-byte
// aligned): the row's values are scalars shared by both chains (broadcast operands of the packed FMAs), so a
// row costs ~75 instructions per chain PAIR instead of ~100 per chain (the scalar kernel is issue-bound).
// 64 threads = 128 chains per CTA, the scalar launch geometry otherwise; same partial layout.
template <int FAM, int PD, bool GROUP>
__global__ void __launch_bounds__(64)
iwls_dense_reg_x2_kernel(const __grid_constant__ EvalParams<float> p, const __grid_constant__ IwlsParams q, int fstride) {
  const int tid = threadIdx.x;
  const int c = 2 * (blockIdx.y * blockDim.x + tid);  // first of this thread's two chains (< Cpad, even)
  const DnsP<float>& D = p.dn[0];
  const size_t ldc = (size_t)p.Cpad;
  auto ld2 = [&](const float* base, size_t row) -> u64 { return *reinterpret_cast<const u64*>(base + row * ldc + c); };
  const u64 ZERO = pk(0.f, 0.f);
  u64 beta[PD], sc[PD], tri[PD * (PD + 1) / 2];
#pragma unroll
  for (int j = 0; j < PD; ++j) {
    beta[j] = ld2(p.pT, (size_t)(D.soff + j));
    sc[j] = ZERO;
  }
#pragma unroll
  for (int j = 0; j < PD * (PD + 1) / 2; ++j) tri[j] = ZERO;
  u64 o0 = ZERO, o1 = ZERO;
  for (int k = 0; k < p.ni; ++k) {
    const u64 v = ld2(p.pT, (size_t)p.ic[k].soff);
    if (p.ic[k].pred) o1 = add2(o1, v); else o0 = add2(o0, v);
  }
  const int f0 = blockIdx.x * fstride, f1 = min(f0 + fstride, p.nchunks);
  const long long r0 = p.chunk_start[f0], r1 = p.chunk_start[f1];
  int curg = -1;
  u64 bg = ZERO;
  auto load_row = [&](long long i, float (&xr)[PD], float& yr, int& gr) {
    const float* xrow = D.X + i * D.ld;
#pragma unroll
    for (int j4 = 0; j4 < PD; j4 += 4) {
      const Vec4<float> v = ld4(xrow + j4);
      xr[j4] = v.a; xr[j4 + 1] = v.b; xr[j4 + 2] = v.c; xr[j4 + 3] = v.d;
    }
    yr = __ldg(p.y + i);
    gr = GROUP ? __ldg(p.gidx + i) : 0;
  };
  float xn[PD], yn = 0.f;
  int gn = 0;
  if (r0 < r1) load_row(r0, xn, yn, gn);
  for (long long i = r0; i < r1; ++i) {
    float x[PD];
#pragma unroll
    for (int j = 0; j < PD; ++j) x[j] = xn[j];
    const float yv = yn;
    const int g = gn;
    if (i + 1 < r1) load_row(i + 1, xn, yn, gn);
    u64 dot = ZERO;
#pragma unroll
    for (int j = 0; j < PD; ++j) dot = fma2(bc(x[j]), beta[j], dot);
    u64 eta0 = o0, eta1 = o1;
    if (D.pred) eta1 = add2(eta1, dot); else eta0 = add2(eta0, dot);
    if constexpr (GROUP) {
      if (g != curg) {  // warp-uniform
        curg = g;
        bg = ld2(p.bT, (size_t)g);
        if (g + 1 < p.G) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.bT + (size_t)(g + 1) * ldc + c));
      }
      if (p.gpred) eta1 = add2(eta1, bg); else eta0 = add2(eta0, bg);
    }
    (void)g;
    float e0a, e0b, e1a, e1b, ll, d0, d1, u0, u1;
    upk(eta0, e0a, e0b);
    upk(eta1, e1a, e1b);
    lik_row<float, FAM, true>(yv, e0a, e1a, p.fixed_inv_scale, ll, d0, d1, u0, u1);
    const float ra = q.tpred ? d1 : d0, Wa = q.tpred ? u1 : u0;
    lik_row<float, FAM, true>(yv, e0b, e1b, p.fixed_inv_scale, ll, d0, d1, u0, u1);
    const u64 r = pk(ra, q.tpred ? d1 : d0), W = pk(Wa, q.tpred ? u1 : u0);
    int u = 0;
#pragma unroll
    for (int a = 0; a < PD; ++a) {
      const u64 xa = bc(x[a]);
      sc[a] = fma2(xa, r, sc[a]);
      const u64 xw = mul2(xa, W);
#pragma unroll
      for (int b = a; b < PD; ++b) {
        tri[u] = fma2(bc(x[b]), xw, tri[u]);
        ++u;
      }
    }
  }
  float* out = reinterpret_cast<float*>(q.partial) + (size_t)blockIdx.x * q.nq * ldc + c;
#pragma unroll
  for (int a = 0; a < PD; ++a) *reinterpret_cast<u64*>(out + (size_t)a * ldc) = sc[a];
#pragma unroll
  for (int j = 0; j < PD * (PD + 1) / 2; ++j) *reinterpret_cast<u64*>(out + (size_t)(PD + j) * ldc) = tri[j];
}

// sum over the chunk partials of one (row, chain) entry in double, chunk order; eight loads in flight
// (the plain loop is a chain of nchunks dependent L2 round trips: 592 at the S3 size)
template <typename T>
__device__ __forceinline__ double chunk_sum(const T* __restrict__ src, size_t stride, int nchunks) {
  double acc = 0.0;
  int ch = 0;
  for (; ch + 8 <= nchunks; ch += 8) {
    T v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = src[(size_t)(ch + u) * stride];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += (double)v[u];
  }
  for (; ch < nchunks; ++ch) acc += (double)src[(size_t)ch * stride];
  return acc;
}

// Sums the chunk partials into chunk 0. One CTA per (32 chains, row): helper g sums the chunks g,
// g + FIN_G, ... in double (eight loads in flight), the helpers' sums are combined in a fixed order.
// Every pass runs it before iwls_finalize_kernel, which then reads one slice: with the chunk loop
// inside the finalize kernel each of its threads walked all nchunks slices serially (71 us of a
// 364 us pass at the S3 size, 559 chunks).
template <typename T, int FIN_G>
__global__ void __launch_bounds__(32 * FIN_G)
iwls_reduce_kernel(T* partial, int nq, int Cpad, int nchunks) {
  __shared__ double red[FIN_G][32];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int row = blockIdx.y;
  const int g = threadIdx.y;
  const size_t stride = (size_t)nq * Cpad;
  const T* src = partial + (size_t)row * Cpad + c;
  double acc = 0.0;
  int ch = g;
  for (; ch + 7 * FIN_G < nchunks; ch += 8 * FIN_G) {
    T v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = src[(size_t)(ch + u * FIN_G) * stride];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += (double)v[u];
  }
  for (; ch < nchunks; ch += FIN_G) acc += (double)src[(size_t)ch * stride];
  red[g][threadIdx.x] = acc;
  __syncthreads();  // every helper has read its slices (chunk 0 included) before chunk 0 is overwritten
  if (g != 0) return;
  acc = 0.0;
#pragma unroll
  for (int k = 0; k < FIN_G; ++k) acc += red[k][threadIdx.x];
  partial[(size_t)row * Cpad + c] = (T)acc;
}

// score [C x p], info [C x p x p] from the chunk partials + prior terms.
// center != NULL (spline target with column-centred basis, rows summing to one, fp.nchunks == 1):
// with s = B^T r, G = B^T W B banded, u = G 1 = B^T W 1, Omega = 1^T u, R = 1^T s
//   score = s - cbar R,   info = G - cbar u^T - u cbar^T + Omega cbar cbar^T.
template <typename T>
__global__ void iwls_finalize_kernel(const __grid_constant__ FinalizeParams fp, int block, int tkind,
                                     int p, int nq, const T* __restrict__ partial,
                                     const T* __restrict__ pT, const T* __restrict__ aux,
                                     T* __restrict__ score, T* __restrict__ info,
                                     const T* __restrict__ center) {
  // thread per (entry e in [0, p + p*p), chain c); c fastest for coalesced partial reads
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int ne = p + p * p;
  if (idx >= (long long)ne * fp.Cpad) return;
  const int c = (int)(idx % fp.Cpad);
  const int e = (int)(idx / fp.Cpad);
  if (c >= fp.C) return;
  const PriorP& pr = fp.pr[block];
  const bool with_prior = !(fp.flags & LSLB_SKIP_PRIOR);
  double tau2 = 1.0;
  if (pr.prior == LSLB_PRIOR_MVND)
    tau2 = pr.tau2_slot >= 0 ? (double)aux[(size_t)c * fp.A + pr.tau2_slot] : pr.tau2_fixed;
  const size_t stride = (size_t)nq * fp.Cpad;
  auto red = [&](int row) { return (double)partial[(size_t)row * fp.Cpad + c]; };
  auto gram = [&](int i, int j) {  // banded entry of the reduced Gram (centred case)
    const int lo = i < j ? i : j, d = i < j ? j - i : i - j;
    return d < 4 ? red(p + 4 * lo + d) : 0.0;
  };
  auto usum = [&](int i) {
    double u = 0.0;
    for (int l = (i > 3 ? i - 3 : 0); l <= (i + 3 < p ? i + 3 : p - 1); ++l) u += gram(i, l);
    return u;
  };
  if (e < p) {
    double acc = 0.0;
    const T* src = partial + (size_t)e * fp.Cpad + c;
    acc = chunk_sum(src, stride, fp.nchunks);
    if (center) {
      double R = 0.0;
      for (int j = 0; j < p; ++j) R += red(j);
      acc -= (double)center[e] * R;
    }
    if (with_prior) {
      if (pr.prior == LSLB_PRIOR_NORMAL) {
        acc -= ((double)pT[(size_t)(pr.soff + e) * fp.Cpad + c] - pr.m) * pr.inv_s2;
      } else if (pr.prior == LSLB_PRIOR_MVND) {
        const T* K = reinterpret_cast<const T*>(pr.K) + (size_t)e * p;
        double kb = 0.0;
        for (int i = 0; i < p; ++i) kb += (double)K[i] * (double)pT[(size_t)(pr.soff + i) * fp.Cpad + c];
        acc -= kb / tau2;
      }
    }
    score[(size_t)c * p + e] = (T)acc;
    return;
  }
  const int a0 = (e - p) / p, b0 = (e - p) % p;
  const int a = a0 < b0 ? a0 : b0, b = a0 < b0 ? b0 : a0;
  double acc = 0.0;
  int row = -1;
  if (tkind == TK_SPLINE) {
    if (b - a < 4) row = p + 4 * a + (b - a);
  } else if (tkind == TK_INTERCEPT) {
    row = 1;
  } else {
    row = p + a * p - a * (a - 1) / 2 + (b - a);
  }
  if (row >= 0) {
    const T* src = partial + (size_t)row * fp.Cpad + c;
    acc = chunk_sum(src, stride, fp.nchunks);
  }
  if (center) {
    double omega = 0.0;
    for (int i = 0; i < p; ++i) omega += usum(i);
    // canonical (a <= b) operand order: entries (i, j) and (j, i) are bitwise equal
    const double ca = (double)center[a], cb = (double)center[b];
    acc += -ca * usum(b) - usum(a) * cb + ca * cb * omega;
  }
  if (with_prior) {
    if (pr.prior == LSLB_PRIOR_NORMAL) {
      if (a == b) acc += pr.inv_s2;
    } else if (pr.prior == LSLB_PRIOR_MVND) {
      acc += (double)reinterpret_cast<const T*>(pr.K)[(size_t)a0 * p + b0] / tau2;
    }
  }
  info[((size_t)c * p + a0) * p + b0] = (T)acc;
}

// ---------------------------------------------------------------------------------------
// ridge + Cholesky: one warp per chain, matrix staged in shared memory
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void chol_kernel(int C, int p, const T* __restrict__ info, T* __restrict__ chol) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * (blockDim.x >> 5) + warp;
  if (c >= C) return;
  const int ld = p + 1;
  T* a = reinterpret_cast<T*>(smraw) + (size_t)warp * p * ld;
  const T* src = info + (size_t)c * p * p;
  bool finite = true;
  T tr = T(0);
  for (int e = lane; e < p * p; e += 32) {
    T v = src[e];
    a[(e / p) * ld + (e % p)] = v;
    if (!isfinite(v)) finite = false;
  }
  __syncwarp();
  for (int j = 0; j < p; ++j) tr += a[j * ld + j];  // same order on every lane
  const T ridge = T(1e-6) * (tr / T(p));           // iwls.py:233-237
  __syncwarp();
  if (lane == 0)
    for (int j = 0; j < p; ++j) a[j * ld + j] += ridge;
  __syncwarp();
  bool ok = __all_sync(0xffffffffu, finite);
  for (int j = 0; j < p && ok; ++j) {
    T d = a[j * ld + j];
    if (!(d > T(0)) || !isfinite(d)) {
      ok = false;
      break;
    }
    d = sqrt(d);
    const T inv = T(1) / d;
    __syncwarp();
    for (int i = j + lane; i < p; i += 32) a[i * ld + j] = (i == j) ? d : a[i * ld + j] * inv;
    __syncwarp();
    for (int i = j + 1 + lane; i < p; i += 32) {
      const T lij = a[i * ld + j];
      for (int k = j + 1; k <= i; ++k) a[i * ld + k] -= lij * a[k * ld + j];
    }
    __syncwarp();
  }
  T* dst = chol + (size_t)c * p * p;
  const T nanv = (T)NAN;
  for (int e = lane; e < p * p; e += 32) {
    int i = e / p, k = e % p;
    dst[e] = ok ? (k <= i ? a[i * ld + k] : T(0)) : nanv;
  }
}

template <typename T>
int launch_chol(int C, int p, const T* info, T* chol, cudaStream_t stream) {
  if (C < 1 || p < 1 || p > LSLB_MAX_IWLS_P || !info || !chol) return LSLB_ERR_INVALID;
  int wpb = 4;
  size_t smem = (size_t)wpb * p * (p + 1) * sizeof(T);
  auto kern = chol_kernel<T>;
  if (smem > 48 * 1024)
    LSLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(C + wpb - 1) / wpb, wpb * 32, smem, stream>>>(C, p, info, chol);
  LSLB_LAUNCH_CHECK();
  return LSLB_OK;
}
template int launch_chol<float>(int, int, const float*, float*, cudaStream_t);
template int launch_chol<double>(int, int, const double*, double*, cudaStream_t);

// ---------------------------------------------------------------------------------------
// proposal tail: one warp per chain
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void propose_kernel(int C, int p, const T* __restrict__ beta, const T* __restrict__ score,
                               const T* __restrict__ chol, const T* __restrict__ step,
                               const T* __restrict__ z, const T* __restrict__ x_eval,
                               T* __restrict__ prop, T* __restrict__ logq) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * (blockDim.x >> 5) + warp;
  if (c >= C) return;
  const int ld = p + 1;
  T* L = reinterpret_cast<T*>(smraw) + (size_t)warp * (p * ld + 3 * p);
  T* v = L + p * ld;   // work vector
  T* mu = v + p;
  T* u = mu + p;
  const T* Lc = chol + (size_t)c * p * p;
  for (int e = lane; e < p * p; e += 32) L[(e / p) * ld + (e % p)] = Lc[e];
  for (int j = lane; j < p; j += 32) v[j] = score[(size_t)c * p + j];
  __syncwarp();
  // forward substitution L y = score (iwls_utils.py:27)
  for (int j = 0; j < p; ++j) {
    const T yj = v[j] / L[j * ld + j];
    __syncwarp();
    if (lane == 0) v[j] = yj;
    for (int i = j + 1 + lane; i < p; i += 32) v[i] -= L[i * ld + j] * yj;
    __syncwarp();
  }
  // backward substitution L^T x = y (iwls_utils.py:28)
  for (int j = p - 1; j >= 0; --j) {
    const T xj = v[j] / L[j * ld + j];
    __syncwarp();
    if (lane == 0) v[j] = xj;
    for (int i = lane; i < j; i += 32) v[i] -= L[j * ld + i] * xj;
    __syncwarp();
  }
  const T st = step[c];
  const T half_s2 = st * st * T(0.5);
  for (int j = lane; j < p; j += 32) mu[j] = beta[(size_t)c * p + j] + half_s2 * v[j];
  __syncwarp();
  if (z != nullptr) {
    // prop = mu + (L/step)^{-T} z = mu + step * L^{-T} z   (iwls_utils.py:68-70)
    for (int j = lane; j < p; j += 32) v[j] = z[(size_t)c * p + j];
    __syncwarp();
    for (int j = p - 1; j >= 0; --j) {
      const T xj = v[j] / L[j * ld + j];
      __syncwarp();
      if (lane == 0) v[j] = xj;
      for (int i = lane; i < j; i += 32) v[i] -= L[j * ld + i] * xj;
      __syncwarp();
    }
    for (int j = lane; j < p; j += 32) {
      const T x = mu[j] + st * v[j];
      u[j] = x;
      prop[(size_t)c * p + j] = x;
    }
  } else {
    for (int j = lane; j < p; j += 32) u[j] = x_eval[(size_t)c * p + j];
  }
  __syncwarp();
  // log q = sum_j [-0.5 s_j^2 - 0.5 log 2pi] + sum_j log(L_jj / step),  s = (x - mu) @ (L/step)
  const T half_log_2pi = T(0.91893853320467274178);
  T acc = T(0);
  for (int j = lane; j < p; j += 32) {
    T s = T(0);
    for (int i = j; i < p; ++i) s += (u[i] - mu[i]) * L[i * ld + j];
    s /= st;
    acc += T(-0.5) * s * s - half_log_2pi + log(L[j * ld + j] / st);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) logq[c] = acc;
}

template <typename T>
int launch_propose(int C, int p, const T* beta, const T* score, const T* chol, const T* step,
                   const T* z, const T* x_eval, T* prop, T* logq, cudaStream_t stream) {
  if (C < 1 || p < 1 || p > LSLB_MAX_IWLS_P || !beta || !score || !chol || !step || !logq)
    return LSLB_ERR_INVALID;
  if ((z == nullptr) == (x_eval == nullptr)) return LSLB_ERR_INVALID;
  if (z != nullptr && !prop) return LSLB_ERR_INVALID;
  int wpb = 4;
  size_t smem = (size_t)wpb * (p * (p + 1) + 3 * p) * sizeof(T);
  auto kern = propose_kernel<T>;
  if (smem > 48 * 1024)
    LSLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(C + wpb - 1) / wpb, wpb * 32, smem, stream>>>(C, p, beta, score, chol, step, z, x_eval,
                                                        prop, logq);
  LSLB_LAUNCH_CHECK();
  return LSLB_OK;
}

// ---------------------------------------------------------------------------------------
// beta^T K beta
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void quadform_kernel(int C, int P, int off, int k, const T* __restrict__ params,
                                const T* __restrict__ K, T* __restrict__ out) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const T* b = params + (size_t)c * P + off;
  double q = 0.0;
  for (int j = 0; j < k; ++j) {
    double xk = 0.0;
    for (int i = 0; i < k; ++i) xk += (double)b[i] * (double)K[(size_t)i * k + j];
    q += xk * (double)b[j];
  }
  out[c] = (T)q;
}

template <typename T>
int launch_quadform(const lslb_plan* plan, int C, int block, const T* params, T* out,
                    cudaStream_t stream) {
  if (!plan || C < 1 || block < 0 || block >= plan->desc.nblocks || !params || !out)
    return LSLB_ERR_INVALID;
  if ((plan->desc.dtype == LSLB_F32) != (sizeof(T) == 4)) return LSLB_ERR_INVALID;
  const lslb_block_desc& d = plan->blk[block];
  if (d.prior != LSLB_PRIOR_MVND) return LSLB_ERR_INVALID;
  quadform_kernel<T><<<(C + 127) / 128, 128, 0, stream>>>(C, plan->P, plan->off[block], d.ncoef,
                                                          params, reinterpret_cast<const T*>(d.K), out);
  LSLB_LAUNCH_CHECK();
  return LSLB_OK;
}

// ---------------------------------------------------------------------------------------
// host side of the IWLS pass
// ---------------------------------------------------------------------------------------
static bool iwls_target(const lslb_plan* plan, int block, int& tkind, int& tidx) {
  if (block < 0 || block >= plan->desc.nblocks) return false;
  for (int m = 0; m < plan->ns; ++m)
    if (plan->sp[m] == block) { tkind = TK_SPLINE; tidx = m; return true; }
  for (int m = 0; m < plan->ni; ++m)
    if (plan->ic[m] == block) { tkind = TK_INTERCEPT; tidx = m; return true; }
  for (int m = 0; m < plan->nd; ++m)
    if (plan->dn[m] == block) { tkind = TK_DENSE; tidx = m; return true; }
  return false;  // group blocks: a G x G information matrix is not materialised
}

static size_t iwls_smem_per_chain(const lslb_plan* plan, int nq) {
  return (size_t)(plan->Psmall + nq) * (plan->desc.dtype == LSLB_F32 ? 4 : 8);
}

template <typename T>
struct IwlsWs {
  T *pT, *partial, *bT, *gbT, *gpl;
  size_t bytes;
};

template <typename T>
static IwlsWs<T> iwls_layout(const lslb_plan* plan, const LaunchCfg& k, int nq, void* ws) {
  IwlsWs<T> L;
  size_t off = 0;
  auto take = [&](size_t n) {
    T* ptr = reinterpret_cast<T*>(reinterpret_cast<char*>(ws) + off);
    off += align_up(n * sizeof(T), 256);
    return ptr;
  };
  L.pT = take((size_t)plan->Psmall * k.Cpad);
  L.partial = take((size_t)k.nchunks * nq * k.Cpad);
  L.bT = L.gbT = L.gpl = nullptr;
  if (plan->grp >= 0) {
    int G = plan->blk[plan->grp].ncoef;
    L.bT = take((size_t)G * k.Cpad);
    L.gbT = take((size_t)G * k.Cpad);
    L.gpl = take((size_t)((G + 31) / 32) * k.Cpad);
  }
  L.bytes = off;
  return L;
}

int launch_banded_iwls_x2(const lslb_plan* plan, const EvalParams<float>& ep, int tk, int tm, int tp,
                          int p, int nq, float* partial, const BandedCfg& k, cudaStream_t st);  // banded_iwls_x2.cu
BandedCfg banded_iwls_cfg(const lslb_plan* plan, int C);

// packed TMA run-length IWLS kernel: float32 banded Gaussian plans, spline or intercept target
static bool iwls_fast(const lslb_plan* plan, int tkind) {
  // the packed IWLS kernel implements the Normal families only; others take the generic kernel
  return plan->desc.dtype == LSLB_F32 && banded_fast_eligible(plan) &&
         (plan->fam == FAM_NORMAL_LS || plan->fam == FAM_NORMAL_FIXED) &&
         (tkind == TK_SPLINE || tkind == TK_INTERCEPT);
}

// dense target of a "dense block in registers" plan with p <= 8: accumulators in registers
static bool iwls_dense_reg(const lslb_plan* plan, int tkind, int p) {
  static const bool off = [] {  // LSLB_NO_IWLS_DENSE_REG=1: the generic kernel (A/B runs)
    const char* e = getenv("LSLB_NO_IWLS_DENSE_REG");
    return e && atoi(e) != 0;
  }();
  return !off && tkind == TK_DENSE && plan->variant == VAR_DENSE_REG && plan->ns == 0 && plan->nd == 1 && p <= 8;
}

static LaunchCfg iwls_cfg(const lslb_plan* plan, int C, int tkind, int nq) {
  if (iwls_fast(plan, tkind)) {
    BandedCfg b = banded_iwls_cfg(plan, C);
    LaunchCfg k;
    k.TC = b.TC;
    k.Cpad = b.Cpad;
    k.ntiles = b.ntiles_chain;
    k.fstride = 0;
    k.nchunks = b.nchunks;
    return k;
  }
  LaunchCfg k = choose_cfg(plan, C, iwls_smem_per_chain(plan, nq));
  if (tkind == TK_DENSE && plan->nd == 1 && iwls_dense_reg(plan, tkind, plan->blk[plan->dn[0]].ncoef) && k.fstride > 8) {
    // iwls_dense_reg_kernel sums a whole row chunk in fp32 registers: at most 8 fine chunks (n / (16 SMs'
    // worth) rows each: 17k rows at S5) per accumulator, whatever the chain count - the regime the full-size
    // parity test covers (test_full_size_s5_iwls_both_blocks); the chunk partials are summed in double
    k.fstride = 8;
    k.nchunks = (plan->nchunks + k.fstride - 1) / k.fstride;
  }
  return k;
}

size_t iwls_workspace_bytes(const lslb_plan* plan, int C) {
  size_t best = 0;
  for (int b = 0; b < plan->desc.nblocks; ++b) {
    int tkind, tidx;
    if (!iwls_target(plan, b, tkind, tidx)) continue;
    int p = plan->blk[b].ncoef;
    if (p > LSLB_MAX_IWLS_P) continue;
    int nq = iwls_nq(tkind, p);
    LaunchCfg k = iwls_cfg(plan, C, tkind, nq);
    size_t bytes = plan->desc.dtype == LSLB_F32 ? iwls_layout<float>(plan, k, nq, nullptr).bytes
                                                : iwls_layout<double>(plan, k, nq, nullptr).bytes;
    if (bytes > best) best = bytes;
  }
  return best;
}

template <typename T, int FAM>
static int iwls_dispatch(const lslb_plan* plan, const EvalParams<T>& ep, const IwlsParams& q,
                         const LaunchCfg& k, size_t smem, cudaStream_t st) {
  dim3 grid(k.nchunks, k.ntiles), block(k.TC);
  if (iwls_dense_reg(plan, q.tkind, q.p)) {
    const lslb_block_desc& d = plan->blk[plan->dn[0]];
    const int pd = q.p <= 4 ? 4 : 8;
    const int vec = d.ncoef == pd && d.ld % 4 == 0 && (reinterpret_cast<uintptr_t>(d.data) & (4 * sizeof(T) - 1)) == 0;
    static const bool no_x2 = [] {  // LSLB_NO_IWLS_DENSE_X2=1: the scalar kernel (A/B runs)
      const char* e = getenv("LSLB_NO_IWLS_DENSE_X2");
      return e && atoi(e) != 0;
    }();
    if constexpr (sizeof(T) == 4) {
      // rows exactly PD wide (q.p == PD): the register triangle IS the layout. From 512 chains: A/B at S5 on one
      // box, ms per pass: 1024 chains 24.8 -> 21.3, 256 chains 6.68 -> 6.81 (no gain: not issue-bound there)
      if (vec && !no_x2 && k.TC % 64 == 0 && k.Cpad >= 512) {
        dim3 block2(k.TC / 2);
        if (plan->grp >= 0) {
          if (pd == 4) iwls_dense_reg_x2_kernel<FAM, 4, true><<<grid, block2, 0, st>>>(ep, q, k.fstride);
          else iwls_dense_reg_x2_kernel<FAM, 8, true><<<grid, block2, 0, st>>>(ep, q, k.fstride);
        } else {
          if (pd == 4) iwls_dense_reg_x2_kernel<FAM, 4, false><<<grid, block2, 0, st>>>(ep, q, k.fstride);
          else iwls_dense_reg_x2_kernel<FAM, 8, false><<<grid, block2, 0, st>>>(ep, q, k.fstride);
        }
        LSLB_LAUNCH_CHECK();
        return LSLB_OK;
      }
    }
    if (plan->grp >= 0) {
      if (q.p <= 4) iwls_dense_reg_kernel<T, FAM, 4, true><<<grid, block, 0, st>>>(ep, q, k.fstride, vec);
      else iwls_dense_reg_kernel<T, FAM, 8, true><<<grid, block, 0, st>>>(ep, q, k.fstride, vec);
    } else {
      if (q.p <= 4) iwls_dense_reg_kernel<T, FAM, 4, false><<<grid, block, 0, st>>>(ep, q, k.fstride, vec);
      else iwls_dense_reg_kernel<T, FAM, 8, false><<<grid, block, 0, st>>>(ep, q, k.fstride, vec);
    }
    LSLB_LAUNCH_CHECK();
    return LSLB_OK;
  }
#define LSLB_IWLS_LAUNCH(NS, GRP)                                                              \
  do {                                                                                         \
    auto kern = iwls_kernel<T, FAM, NS, GRP>;                                                  \
    if (smem > 48 * 1024)                                                                      \
      LSLB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<grid, block, smem, st>>>(ep, q, k.fstride);                                         \
  } while (0)
  if (plan->grp >= 0) LSLB_IWLS_LAUNCH(8, true);
  else if (plan->ns <= 2) LSLB_IWLS_LAUNCH(2, false);
  else LSLB_IWLS_LAUNCH(8, false);
#undef LSLB_IWLS_LAUNCH
  LSLB_LAUNCH_CHECK();
  return LSLB_OK;
}

template <typename T>
int launch_iwls(const lslb_plan* plan, int C, int block, const T* params, const T* aux, T* score,
                T* info, T* chol, unsigned flags, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (!plan || C < 1 || !params || !score || !info || !ws) return LSLB_ERR_INVALID;
  if (plan->A > 0 && !aux) return LSLB_ERR_INVALID;
  if ((plan->desc.dtype == LSLB_F32) != (sizeof(T) == 4)) return LSLB_ERR_INVALID;
  int tkind, tidx;
  if (!iwls_target(plan, block, tkind, tidx)) return LSLB_ERR_UNSUPPORTED;
  const int p = plan->blk[block].ncoef;
  if (p > LSLB_MAX_IWLS_P) return LSLB_ERR_UNSUPPORTED;
  // the centring correction of the target block uses u = G 1, which needs basis rows summing to one
  if (plan->blk[block].center && !plan->unit_rows) return LSLB_ERR_UNSUPPORTED;
  const int nq = iwls_nq(tkind, p);
  const bool fast = iwls_fast(plan, tkind);
  LaunchCfg k = iwls_cfg(plan, C, tkind, nq);
  size_t smem = fast ? 0 : iwls_smem_per_chain(plan, nq) * k.TC;
  if (smem > 200 * 1024) return LSLB_ERR_UNSUPPORTED;
  IwlsWs<T> L = iwls_layout<T>(plan, k, nq, ws);
  if (L.bytes > ws_bytes) return LSLB_ERR_WORKSPACE;

  FinalizeParams fp;
  fill_finalize_params(plan, fp, C, k.Cpad, k.nchunks, flags);
  EvalParams<T> ep;
  fill_eval_params<T>(plan, ep, C, k.Cpad);
  ep.pT = L.pT;
  ep.partial = nullptr;
  ep.bT = L.bT;
  ep.gbT = L.gbT;
  IwlsParams q;
  q.tkind = tkind;
  q.tidx = tidx;
  q.p = p;
  q.tpred = plan->blk[block].predictor;
  q.nq = nq;
  q.partial = L.partial;

  if (int gst = launch_gather_small<T>(fp, params, L.pT, stream)) return gst;
  if (plan->grp >= 0) {
    const lslb_block_desc& d = plan->blk[plan->grp];
    dim3 grid((d.ncoef + 31) / 32, (k.Cpad + 31) / 32), blockd(32, 8);
    group_init_kernel<T><<<grid, blockd, 0, stream>>>(params, plan->P, plan->off[plan->grp], d.ncoef,
                                                      C, k.Cpad, (T)d.prior_m, T(0), false, L.bT,
                                                      L.gbT, L.gpl);
    LSLB_LAUNCH_CHECK();
  }
  int st = LSLB_ERR_INVALID;
  debug_record_start(stream);
  if (fast) {
    if constexpr (sizeof(T) == 4)
      st = launch_banded_iwls_x2(plan, ep, tkind == TK_SPLINE ? 0 : 1, tidx, q.tpred, p, nq,
                                 reinterpret_cast<float*>(L.partial), banded_iwls_cfg(plan, C), stream);
  } else
  switch (plan->fam) {
    case FAM_NORMAL_LS: st = iwls_dispatch<T, FAM_NORMAL_LS>(plan, ep, q, k, smem, stream); break;
    case FAM_NORMAL_FIXED: st = iwls_dispatch<T, FAM_NORMAL_FIXED>(plan, ep, q, k, smem, stream); break;
    case FAM_BERNOULLI: st = iwls_dispatch<T, FAM_BERNOULLI>(plan, ep, q, k, smem, stream); break;
    case FAM_POISSON: st = iwls_dispatch<T, FAM_POISSON>(plan, ep, q, k, smem, stream); break;
  }
  debug_record_stop(stream);
  if (st != LSLB_OK) return st;
  long long ne = (long long)(p + p * p) * k.Cpad;
  const T* center = reinterpret_cast<const T*>(plan->blk[block].center);
  if (k.nchunks > 1) {
    dim3 rgrid(k.Cpad / 32, nq);
    if ((long long)rgrid.x * rgrid.y >= 4LL * plan->sm_count)
      iwls_reduce_kernel<T, 8><<<rgrid, dim3(32, 8), 0, stream>>>(L.partial, nq, k.Cpad, k.nchunks);
    else
      iwls_reduce_kernel<T, 32><<<rgrid, dim3(32, 32), 0, stream>>>(L.partial, nq, k.Cpad, k.nchunks);
    LSLB_LAUNCH_CHECK();
  }
  fp.nchunks = 1;
  iwls_finalize_kernel<T><<<(unsigned)((ne + 255) / 256), 256, 0, stream>>>(
      fp, block, tkind, p, nq, L.partial, L.pT, aux, score, info, center);
  LSLB_LAUNCH_CHECK();
  if (chol) return launch_chol<T>(C, p, info, chol, stream);
  return LSLB_OK;
}
template int launch_iwls<float>(const lslb_plan*, int, int, const float*, const float*, float*,
                                float*, float*, unsigned, void*, size_t, cudaStream_t);
template int launch_iwls<double>(const lslb_plan*, int, int, const double*, const double*, double*,
                                 double*, double*, unsigned, void*, size_t, cudaStream_t);

// ---------------------------------------------------------------------------------------
// Metropolis-Hastings decision + state selection (reference goose/mh.py:48-68), one CTA per chain:
//   log_acc = proposed - current + correction;  NaN -> (-inf, error code 90);
//   acc = min(exp(log_acc), 1);  accept <=> u <= acc
// and, fused, what the kernels do with the decision (iwls.py:337-342, mh.py:62-66): the accepted
// proposal replaces the chain's row of `params` and its carried log-posterior.
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128)
mh_step_kernel(int P, const T* __restrict__ u, T* __restrict__ cur_lp, const T* __restrict__ prop_lp,
               const T* __restrict__ log_corr, T* __restrict__ params, const T* __restrict__ params_prop,
               int* __restrict__ code, T* __restrict__ acc, unsigned char* __restrict__ moved) {
  const int c = blockIdx.x;
  T la = prop_lp[c] - cur_lp[c] + (log_corr ? log_corr[c] : T(0));
  const bool nan = la != la;
  if (nan) la = -INFINITY;
  T a = exp(la);  // exp(-inf) = 0, exp(+inf) = inf
  if (a > T(1)) a = T(1);
  const bool accept = u[c] <= a;
  __syncthreads();  // every thread has read cur_lp[c]
  if (threadIdx.x == 0) {
    if (code) code[c] = nan ? 90 : 0;
    if (acc) acc[c] = a;
    if (moved) moved[c] = accept ? 1 : 0;
    if (accept) cur_lp[c] = prop_lp[c];
  }
  if (accept && params && params_prop) {
    T* dst = params + (size_t)c * P;
    const T* src = params_prop + (size_t)c * P;
    for (int j = threadIdx.x; j < P; j += blockDim.x) dst[j] = src[j];
  }
}

template <typename T>
int launch_mh_step(int C, int P, const T* u, T* cur_lp, const T* prop_lp, const T* log_corr, T* params,
                   const T* params_prop, int* code, T* acc, unsigned char* moved, cudaStream_t stream) {
  if (C < 1 || P < 0 || !u || !cur_lp || !prop_lp) return LSLB_ERR_INVALID;
  if ((params == nullptr) != (params_prop == nullptr)) return LSLB_ERR_INVALID;
  mh_step_kernel<T><<<C, 128, 0, stream>>>(P, u, cur_lp, prop_lp, log_corr, params, params_prop, code, acc, moved);
  LSLB_LAUNCH_CHECK();
  return LSLB_OK;
}

}  // namespace lslb

#define LSLB_S(x) ((cudaStream_t)(x))
extern "C" int lslb_mh_step_f32(int C, int P, const float* u, float* cur_lp, const float* prop_lp,
                                const float* log_corr, float* params, const float* params_prop, int* code,
                                float* acc, unsigned char* moved, lslb_stream_t s) {
  return lslb::launch_mh_step<float>(C, P, u, cur_lp, prop_lp, log_corr, params, params_prop, code, acc, moved, LSLB_S(s));
}
extern "C" int lslb_mh_step_f64(int C, int P, const double* u, double* cur_lp, const double* prop_lp,
                                const double* log_corr, double* params, const double* params_prop, int* code,
                                double* acc, unsigned char* moved, lslb_stream_t s) {
  return lslb::launch_mh_step<double>(C, P, u, cur_lp, prop_lp, log_corr, params, params_prop, code, acc, moved, LSLB_S(s));
}
extern "C" int lslb_iwls_f32(const lslb_plan* plan, int C, int block, const float* params,
                             const float* aux, float* score, float* info, float* chol,
                             uint32_t flags, void* ws, size_t ws_bytes, lslb_stream_t s) {
  return lslb::launch_iwls<float>(plan, C, block, params, aux, score, info, chol, flags, ws, ws_bytes, LSLB_S(s));
}
extern "C" int lslb_iwls_f64(const lslb_plan* plan, int C, int block, const double* params,
                             const double* aux, double* score, double* info, double* chol,
                             uint32_t flags, void* ws, size_t ws_bytes, lslb_stream_t s) {
  return lslb::launch_iwls<double>(plan, C, block, params, aux, score, info, chol, flags, ws, ws_bytes, LSLB_S(s));
}
extern "C" int lslb_chol_f32(int C, int p, const float* info, float* chol, lslb_stream_t s) {
  return lslb::launch_chol<float>(C, p, info, chol, LSLB_S(s));
}
extern "C" int lslb_chol_f64(int C, int p, const double* info, double* chol, lslb_stream_t s) {
  return lslb::launch_chol<double>(C, p, info, chol, LSLB_S(s));
}
extern "C" int lslb_quadform_f32(const lslb_plan* plan, int C, int block, const float* params,
                                 float* out, lslb_stream_t s) {
  return lslb::launch_quadform<float>(plan, C, block, params, out, LSLB_S(s));
}
extern "C" int lslb_quadform_f64(const lslb_plan* plan, int C, int block, const double* params,
                                 double* out, lslb_stream_t s) {
  return lslb::launch_quadform<double>(plan, C, block, params, out, LSLB_S(s));
}
extern "C" int lslb_iwls_propose_f32(int C, int p, const float* beta, const float* score,
                                     const float* chol, const float* step, const float* z,
                                     const float* x_eval, float* prop, float* logq,
                                     lslb_stream_t s) {
  return lslb::launch_propose<float>(C, p, beta, score, chol, step, z, x_eval, prop, logq, LSLB_S(s));
}
extern "C" int lslb_iwls_propose_f64(int C, int p, const double* beta, const double* score,
                                     const double* chol, const double* step, const double* z,
                                     const double* x_eval, double* prop, double* logq,
                                     lslb_stream_t s) {
  return lslb::launch_propose<double>(C, p, beta, score, chol, step, z, x_eval, prop, logq, LSLB_S(s));
}
