// Banded B-spline basis (order 0..3; cubic is the default) on the device.
//
// Same recursion and operation order as the reference's `_build_basis_vector`
// (src/liesel/contrib/splines.py:73-106): half-open base case `x >= t[i] && x < t[i+1]`
// (:86-90) decided by COMPARISON against the stored knot array (never by floor((x-t0)/h)),
// then b1 = (x - t[i]) / (t[i+m] - t[i]) * bv[i], b2 = (t[i+m+1] - x) / (t[i+m+1] - t[i+1]) * bv[i+1],
// bv[i] = b1 + b2 (:93-98). Only the <= 4 entries that can be non-zero are carried; the
// entries the dense recursion multiplies by an exact 0 are skipped, which leaves every
// result bit-identical. Round-to-nearest intrinsics keep the compiler from contracting
// a*b + c*d into an FMA, so the weights agree bit for bit with a NumPy evaluation.
#include "common.cuh"

namespace lslb {

__device__ __forceinline__ float op_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float op_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float op_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float op_div(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double op_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double op_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double op_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double op_div(double a, double b) { return __ddiv_rn(a, b); }

// ORD = spline order (degree) 0..3. The band written is always FOUR wide (what every evaluation
// kernel streams): the ORD + 1 weights that can be non-zero sit inside the window
// [start, start + 3] with start = clip(span - ORD, 0, k - 4), the rest of the window is exact zeros.
template <typename T, int ORD>
__global__ void spline_basis_kernel(const T* __restrict__ x, long long n, const T* __restrict__ knots,
                                    int nk, int* __restrict__ start, T* __restrict__ weights) {
  extern __shared__ __align__(16) unsigned char smraw[];
  T* t = reinterpret_cast<T*>(smraw);
  for (int i = threadIdx.x; i < nk; i += blockDim.x) t[i] = knots[i];
  __syncthreads();
  const int k = nk - ORD - 1;  // number of basis functions
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n;
       r += (long long)gridDim.x * blockDim.x) {
    const T xv = x[r];
    // j = (number of knots <= x) - 1, by binary search on the stored knots
    int lo = 0, hi = nk;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (t[mid] <= xv) lo = mid + 1; else hi = mid;
    }
    const int j = lo - 1;
    T w[4] = {T(0), T(0), T(0), T(0)};
    int s = 0;
    if (j >= 0 && j <= nk - 2 && xv < t[j + 1]) {
      // v[u] = value of basis function i = j - ORD + u at the current level
      T v[ORD + 1];
#pragma unroll
      for (int u = 0; u < ORD; ++u) v[u] = T(0);
      v[ORD] = T(1);
#pragma unroll
      for (int m = 1; m <= ORD; ++m) {
        T nv[ORD + 1];
#pragma unroll
        for (int u = 0; u <= ORD; ++u) nv[u] = T(0);
#pragma unroll
        for (int u = 0; u <= ORD; ++u) {
          if (u < ORD - m) continue;
          const int i = j - ORD + u;
          if (i < 0 || i + m + 1 > nk - 1) continue;  // outside the valid triangle
          T b1 = T(0), b2 = T(0);
          const bool has1 = u > ORD - m;  // bv[i] non-zero at level m-1
          const bool has2 = u < ORD;      // bv[i+1] non-zero at level m-1
          if (has1) b1 = op_mul(op_div(op_sub(xv, t[i]), op_sub(t[i + m], t[i])), v[u]);
          if (has2) {
            const T vn = v[u + 1 <= ORD ? u + 1 : ORD];
            b2 = op_mul(op_div(op_sub(t[i + m + 1], xv), op_sub(t[i + m + 1], t[i + 1])), vn);
          }
          nv[u] = has1 ? (has2 ? op_add(b1, b2) : b1) : b2;
        }
#pragma unroll
        for (int u = 0; u <= ORD; ++u) v[u] = nv[u];
      }
      s = j - ORD;
      if (s < 0) s = 0;
      if (s > k - 4) s = k - 4;
#pragma unroll
      for (int tt = 0; tt < 4; ++tt) {
        const int i = s + tt;            // basis function index
        const int u = i - (j - ORD);     // position in v
        T val = T(0);
#pragma unroll
        for (int uu = 0; uu <= ORD; ++uu)
          if (uu == u && i < k) val = v[uu];
        w[tt] = val;
      }
    }
    start[r] = s;
    T* dst = weights + 4 * r;
    dst[0] = w[0]; dst[1] = w[1]; dst[2] = w[2]; dst[3] = w[3];
  }
}

template <typename T>
int launch_spline_basis(const T* x, long long n, const T* knots, int nk, int order, int* start, T* weights,
                        cudaStream_t stream) {
  if (n < 0 || order < 0 || order > 3 || nk < order + 5 || nk > 4096 || !knots) return LSLB_ERR_INVALID;
  if (n == 0) return LSLB_OK;
  if (!x || !start || !weights) return LSLB_ERR_INVALID;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  const unsigned g = (unsigned)blocks;
  const size_t sm = nk * sizeof(T);
  switch (order) {
    case 0: spline_basis_kernel<T, 0><<<g, 256, sm, stream>>>(x, n, knots, nk, start, weights); break;
    case 1: spline_basis_kernel<T, 1><<<g, 256, sm, stream>>>(x, n, knots, nk, start, weights); break;
    case 2: spline_basis_kernel<T, 2><<<g, 256, sm, stream>>>(x, n, knots, nk, start, weights); break;
    default: spline_basis_kernel<T, 3><<<g, 256, sm, stream>>>(x, n, knots, nk, start, weights); break;
  }
  LSLB_LAUNCH_CHECK();
  return LSLB_OK;
}

}  // namespace lslb

extern "C" int lslb_spline_basis_f32(const float* x, int64_t n, const float* knots, int nknots,
                                     int32_t* start, float* weights, lslb_stream_t s) {
  return lslb::launch_spline_basis<float>(x, n, knots, nknots, 3, start, weights, (cudaStream_t)s);
}
extern "C" int lslb_spline_basis_f64(const double* x, int64_t n, const double* knots, int nknots,
                                     int32_t* start, double* weights, lslb_stream_t s) {
  return lslb::launch_spline_basis<double>(x, n, knots, nknots, 3, start, weights, (cudaStream_t)s);
}
extern "C" int lslb_spline_basis_order_f32(const float* x, int64_t n, const float* knots, int nknots, int order,
                                           int32_t* start, float* weights, lslb_stream_t s) {
  return lslb::launch_spline_basis<float>(x, n, knots, nknots, order, start, weights, (cudaStream_t)s);
}
extern "C" int lslb_spline_basis_order_f64(const double* x, int64_t n, const double* knots, int nknots, int order,
                                           int32_t* start, double* weights, lslb_stream_t s) {
  return lslb::launch_spline_basis<double>(x, n, knots, nknots, order, start, weights, (cudaStream_t)s);
}
