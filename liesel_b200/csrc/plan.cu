// Plan construction: validates the model description, fixes the parameter layout, derives
// the row chunking (aligned to group boundaries when a random intercept is present) and the
// row-independent likelihood constant. Setup-time code: synchronous by design.
#include <stdio.h>
#include <string.h>

#include <new>

#include "common.cuh"

namespace lslb {

static thread_local char g_err[256] = "";
void set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
}

void set_comm_error(const char* msg) { snprintf(g_err, sizeof(g_err), "%s", msg); }

// fine chunk boundaries: fine[f] = f*n/F moved forward to the next group start
__global__ void fine_bounds_kernel(long long n, int F, const int* __restrict__ gidx,
                                   long long* __restrict__ fine) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f > F) return;
  long long r = (f == F) ? n : (long long)((double)f * (double)n / (double)F);
  if (r > n) r = n;
  if (gidx != nullptr && f != 0 && f != F) {
    while (r < n && r > 0 && gidx[r] == gidx[r - 1]) ++r;
  }
  fine[f] = r;
}

// number of rows whose span differs from the previous row's (one block, setup time)
__global__ void span_changes_kernel(const int* __restrict__ start, long long n, unsigned long long* out) {
  __shared__ unsigned long long sh[1024];
  unsigned long long acc = 0;
  for (long long i = 1 + threadIdx.x; i < n; i += blockDim.x) acc += start[i] != start[i - 1];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sh[0];
}

// out[0] = min, out[1] = max of an index array, out[2] = number of descents (one block, setup time)
__global__ void index_range_kernel(const int* __restrict__ idx, long long n, int* out) {
  __shared__ int lo[1024], hi[1024], dn[1024];
  int a = 0x7fffffff, b = -0x7fffffff - 1, d = 0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const int v = idx[i];
    a = min(a, v);
    b = max(b, v);
    if (i > 0 && idx[i - 1] > v) ++d;
  }
  lo[threadIdx.x] = a;
  hi[threadIdx.x] = b;
  dn[threadIdx.x] = d;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      lo[threadIdx.x] = min(lo[threadIdx.x], lo[threadIdx.x + s]);
      hi[threadIdx.x] = max(hi[threadIdx.x], hi[threadIdx.x + s]);
      dn[threadIdx.x] += dn[threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = lo[0]; out[1] = hi[0]; out[2] = dn[0]; }
}

// max_i |w_i0 + w_i1 + w_i2 + w_i3 - 1| over the rows of one spline block (one block, setup time)
template <typename T>
__global__ void unit_rows_kernel(const T* __restrict__ w, long long n, double* out) {
  __shared__ double sh[1024];
  double acc = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const double s = (double)w[4 * i] + (double)w[4 * i + 1] + (double)w[4 * i + 2] + (double)w[4 * i + 3];
    acc = fmax(acc, fabs(s - 1.0));
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sh[0];
}

// number of rows in which two spline blocks differ (span or any of the four weights, bit for bit)
__global__ void band_diff_kernel(const int* __restrict__ s0, const int* __restrict__ s1, const float* __restrict__ w0,
                                 const float* __restrict__ w1, long long n, unsigned long long* out) {
  unsigned long long cnt = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    bool d = s0[i] != s1[i];
#pragma unroll
    for (int u = 0; u < 4; ++u) d = d || (__float_as_uint(w0[4 * i + u]) != __float_as_uint(w1[4 * i + u]));
    cnt += d;
  }
  if (cnt) atomicAdd(out, cnt);
}

// gstart[g] = first row whose group index is >= g (rows ascending by group), gstart[G] = n: thread i looks at
// the boundary between rows i - 1 and i and fills the entries of the groups that start there (a group
// without rows gets an empty range)
__global__ void group_starts_kernel(const int* __restrict__ idx, long long n, int G, int* __restrict__ gstart) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  const int prev = i == 0 ? -1 : idx[i - 1];
  const int cur = i == n ? G : idx[i];
  for (int g = prev + 1; g <= cur; ++g) gstart[g] = (int)i;
}

// per 128-row tile the sums of the four weight columns of one spline block (one block per tile). Entry 1 is
// not a sum (the closed form of banded_x2.cu needs columns 0, 2, 3 only): it carries, as an int bit pattern,
// the tile's knot span when all of its rows share one (and the tile is full), else -1 - the kernel's test
// "is the span constant over this tile" becomes one shared-memory read instead of a loop over the rows
__global__ void tile_weight_sums_kernel(const float* __restrict__ w, const int* __restrict__ start, long long n,
                                        float* __restrict__ out) {
  const long long r0 = (long long)blockIdx.x * 128;
  const int u = threadIdx.x & 3, q = threadIdx.x >> 2;  // 32 row slots x 4 columns
  double acc = 0.0;
  const int first = start[r0];
  int same = 1;
  for (int r = q; r < 128 && r0 + r < n; r += 32) {
    acc += (double)w[4 * (r0 + r) + u];
    same &= start[r0 + r] == first;
  }
  __shared__ double sh[128];
  sh[threadIdx.x] = acc;
  const int all_same = __syncthreads_and(same);
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (int j = 0; j < 32; ++j) t += sh[4 * j + threadIdx.x];  // fixed order
    float v = (float)t;
    if (threadIdx.x == 1) v = __int_as_float((all_same && r0 + 128 <= n) ? first : -1);
    out[4 * (long long)blockIdx.x + threadIdx.x] = v;
  }
}

// per row the ten products w_a w_b (a <= b) of one spline block's weights, padded to 12 floats: the
// IWLS kernel's Gram update reads them as broadcast scalars (they are the same for every chain)
__global__ void weight_products_kernel(const float* __restrict__ w, long long n, float* __restrict__ out) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const float4 v = *reinterpret_cast<const float4*>(w + 4 * r);
  float4* o = reinterpret_cast<float4*>(out + 12 * r);
  o[0] = make_float4(v.x * v.x, v.x * v.y, v.x * v.z, v.x * v.w);
  o[1] = make_float4(v.y * v.y, v.y * v.z, v.y * v.w, v.z * v.z);
  o[2] = make_float4(v.z * v.w, v.w * v.w, 0.f, 0.f);
}

// sum_i lgamma(y_i + 1) in double, one block, fixed order (deterministic)
template <typename T>
__global__ void lgamma_sum_kernel(const T* __restrict__ y, long long n, double* out) {
  __shared__ double sh[1024];
  double acc = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) acc += lgamma((double)y[i] + 1.0);
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sh[0];
}

}  // namespace lslb

using namespace lslb;

extern "C" int lslb_version(void) { return LSLB_VERSION; }

extern "C" const char* lslb_status_string(int s) {
  switch (s) {
    case LSLB_OK: return "ok";
    case LSLB_ERR_INVALID: return "invalid argument or unsupported model structure";
    case LSLB_ERR_WORKSPACE: return "workspace too small";
    case LSLB_ERR_CUDA: return "CUDA runtime error";
    case LSLB_ERR_UNSUPPORTED: return "not implemented for this configuration";
    case LSLB_ERR_COMM: return "collective (NCCL) error";
    default: return "unknown status";
  }
}

extern "C" const char* lslb_last_cuda_error(void) { return g_err; }

extern "C" int lslb_plan_create(const lslb_model_desc* desc, lslb_plan** out) {
  if (!desc || !out) return LSLB_ERR_INVALID;
  *out = nullptr;
  if (desc->nblocks < 1 || desc->nblocks > LSLB_MAX_BLOCKS || desc->n < 0 || !desc->blocks)
    return LSLB_ERR_INVALID;
  if (desc->dtype != LSLB_F32 && desc->dtype != LSLB_F64) return LSLB_ERR_INVALID;
  if (desc->family < LSLB_NORMAL || desc->family > LSLB_POISSON) return LSLB_ERR_INVALID;
  if (desc->n > 0 && !desc->y) return LSLB_ERR_INVALID;

  lslb_plan* p = new (std::nothrow) lslb_plan();
  if (!p) return LSLB_ERR_INVALID;
  memset(p, 0, sizeof(*p));
  p->desc = *desc;
  p->grp = -1;
  int off = 0, soff = 0, aux = 0;
  for (int b = 0; b < desc->nblocks; ++b) {
    const lslb_block_desc& d = desc->blocks[b];
    p->blk[b] = d;
    bool bad = d.ncoef < 1 || d.predictor < 0 || d.predictor > 1 ||
               (d.predictor == 1 && desc->family != LSLB_NORMAL) || d.prior < LSLB_PRIOR_NONE ||
               d.prior > LSLB_PRIOR_MVND;
    if (d.prior == LSLB_PRIOR_MVND && (!d.K || d.kind == LSLB_GROUP)) bad = true;
    if (d.prior == LSLB_PRIOR_NORMAL && !(d.prior_s > 0)) bad = true;
    switch (d.kind) {
      case LSLB_DENSE:
        if (!d.data || d.ld < d.ncoef || p->nd >= LSLB_MAX_DNS) bad = true;
        else p->dn[p->nd++] = b;
        break;
      case LSLB_SPLINE:
        if (!d.data || !d.index || d.ncoef < 4 || p->ns >= LSLB_MAX_SPL) bad = true;
        else p->sp[p->ns++] = b;
        break;
      case LSLB_GROUP:
        if (!d.index || p->grp >= 0) bad = true;
        else p->grp = b;
        break;
      case LSLB_INTERCEPT:
        if (d.ncoef != 1 || p->ni >= LSLB_MAX_INT - 2) bad = true;
        else p->ic[p->ni++] = b;
        break;
      default: bad = true;
    }
    if (d.center && d.kind != LSLB_SPLINE) bad = true;
    if (bad) {
      delete p;
      return LSLB_ERR_INVALID;
    }
    p->off[b] = off;
    off += d.ncoef;
    if (d.kind == LSLB_GROUP) {
      p->soff[b] = -1;
    } else {
      p->soff[b] = soff;
      soff += d.ncoef;
    }
    if (d.prior == LSLB_PRIOR_MVND && d.tau2_slot >= 0) aux = aux > d.tau2_slot + 1 ? aux : d.tau2_slot + 1;
    if (d.predictor == 1) p->has_scale = true;
  }
  p->desc.blocks = p->blk;
  p->P = off;
  p->Preal = soff;
  // centred spline blocks: (B - 1 cbar^T) beta = B beta + v with v = -cbar^T beta constant over the
  // rows, so each predictor gets ONE virtual intercept row; its gradient (the residual sum) is
  // folded back into the blocks' gradients by the finalize step.
  p->vrow[0] = p->vrow[1] = -1;
  for (int b = 0; b < desc->nblocks; ++b)
    if (p->blk[b].center && p->vrow[p->blk[b].predictor] < 0) p->vrow[p->blk[b].predictor] = soff++;
  p->Psmall = soff;
  p->A = aux;
  p->fam = desc->family == LSLB_NORMAL ? (p->has_scale ? FAM_NORMAL_LS : FAM_NORMAL_FIXED)
                                       : desc->family;
  if (p->fam == FAM_NORMAL_FIXED && !(desc->fixed_scale > 0)) {
    delete p;
    return LSLB_ERR_INVALID;
  }

  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) {
    set_cuda_error(e, "device query");
    delete p;
    return LSLB_ERR_CUDA;
  }

  // fine chunking: at most 16 chunks per SM, at least 32 rows per chunk
  long long n = desc->n;
  long long F = (long long)p->sm_count * 16;
  if (F > (n + 31) / 32) F = (n + 31) / 32;
  if (F < 1) F = 1;
  p->nchunks = (int)F;
  e = cudaMalloc(&p->d_chunk_start, sizeof(long long) * (F + 1));
  if (e != cudaSuccess) {
    set_cuda_error(e, "cudaMalloc(chunk_start)");
    delete p;
    return LSLB_ERR_CUDA;
  }
  const int* gidx = p->grp >= 0 ? p->blk[p->grp].index : nullptr;
  fine_bounds_kernel<<<(unsigned)((F + 1 + 255) / 256), 256>>>(n, (int)F, gidx, p->d_chunk_start);

  // row-independent likelihood constant
  const double half_log_2pi = 0.91893853320467274178;
  p->lik_const = 0.0;
  if (desc->family == LSLB_NORMAL) {
    p->lik_const = -half_log_2pi * (double)n;
    if (p->fam == FAM_NORMAL_FIXED) p->lik_const -= log(desc->fixed_scale) * (double)n;
  } else if (desc->family == LSLB_POISSON && n > 0) {
    double* d_out = nullptr;
    e = cudaMalloc(&d_out, sizeof(double));
    if (e == cudaSuccess) {
      if (desc->dtype == LSLB_F32)
        lgamma_sum_kernel<float><<<1, 1024>>>((const float*)desc->y, n, d_out);
      else
        lgamma_sum_kernel<double><<<1, 1024>>>((const double*)desc->y, n, d_out);
      double h = 0.0;
      e = cudaMemcpy(&h, d_out, sizeof(double), cudaMemcpyDeviceToHost);
      p->lik_const = -h;
      cudaFree(d_out);
    }
  }
  // The kernels index coefficient storage with the caller's integers (spline: start .. start + 3,
  // group: idx): validate their range once, here — an out-of-range value would be an out-of-bounds
  // shared-/global-memory access in every later call. Group rows must also be ascending (the
  // segment sums are deterministic because a group's rows are contiguous).
  bool bad_index = false;
  if (n > 0 && (p->ns > 0 || p->grp >= 0) && e == cudaSuccess) {
    int* d_rng = nullptr;
    e = cudaMalloc(&d_rng, sizeof(int) * 3 * (LSLB_MAX_SPL + 1));
    if (e == cudaSuccess) {
      int h[3 * (LSLB_MAX_SPL + 1)];
      for (int m = 0; m < p->ns; ++m) index_range_kernel<<<1, 1024>>>(p->blk[p->sp[m]].index, n, d_rng + 3 * m);
      if (p->grp >= 0) index_range_kernel<<<1, 1024>>>(p->blk[p->grp].index, n, d_rng + 3 * p->ns);
      e = cudaMemcpy(h, d_rng, sizeof(int) * 3 * (p->ns + 1), cudaMemcpyDeviceToHost);
      if (e == cudaSuccess) {
        for (int m = 0; m < p->ns; ++m)
          if (h[3 * m] < 0 || h[3 * m + 1] > p->blk[p->sp[m]].ncoef - 4) bad_index = true;
        if (p->grp >= 0) {
          const int* g = h + 3 * p->ns;
          if (g[0] < 0 || g[1] >= p->blk[p->grp].ncoef || g[2] != 0) bad_index = true;
        }
      }
      cudaFree(d_rng);
    }
  }
  if (bad_index) {
    cudaFree(p->d_chunk_start);
    delete p;
    return LSLB_ERR_INVALID;
  }
  // random intercept: row range of every group, for the diagonal IWLS pass (group_iwls.cu)
  p->d_gstart = nullptr;
  if (p->grp >= 0 && e == cudaSuccess && n < 2147483647LL) {
    const int G = p->blk[p->grp].ncoef;
    e = cudaMalloc(&p->d_gstart, sizeof(int) * ((size_t)G + 1));
    if (e == cudaSuccess)
      group_starts_kernel<<<(unsigned)((n + 1 + 255) / 256), 256>>>(p->blk[p->grp].index, n, G, p->d_gstart);
  }
  // are the rows ordered for long constant-span runs?
  p->rows_sorted = p->ns > 0;
  if (p->ns > 0 && e == cudaSuccess) {
    unsigned long long* d_cnt = nullptr;
    e = cudaMalloc(&d_cnt, sizeof(unsigned long long) * LSLB_MAX_SPL);
    if (e == cudaSuccess) {
      unsigned long long h[LSLB_MAX_SPL] = {0};
      for (int m = 0; m < p->ns; ++m)
        span_changes_kernel<<<1, 1024>>>(p->blk[p->sp[m]].index, n, d_cnt + m);
      e = cudaMemcpy(h, d_cnt, sizeof(unsigned long long) * p->ns, cudaMemcpyDeviceToHost);
      for (int m = 0; m < p->ns; ++m)
        if ((long long)h[m] > n / 64 + 64) p->rows_sorted = false;
      cudaFree(d_cnt);
    }
  }
  // do the four weights of every spline row sum to one? (true when all x lie inside the knot range)
  p->unit_rows = false;
  if (p->ns > 0 && e == cudaSuccess) {
    double* d_mx = nullptr;
    e = cudaMalloc(&d_mx, sizeof(double) * LSLB_MAX_SPL);
    if (e == cudaSuccess) {
      double h[LSLB_MAX_SPL] = {0};
      for (int m = 0; m < p->ns; ++m) {
        if (desc->dtype == LSLB_F32)
          unit_rows_kernel<float><<<1, 1024>>>((const float*)p->blk[p->sp[m]].data, n, d_mx + m);
        else
          unit_rows_kernel<double><<<1, 1024>>>((const double*)p->blk[p->sp[m]].data, n, d_mx + m);
      }
      e = cudaMemcpy(h, d_mx, sizeof(double) * p->ns, cudaMemcpyDeviceToHost);
      p->unit_rows = true;
      for (int m = 0; m < p->ns; ++m)
        if (!(h[m] <= (desc->dtype == LSLB_F32 ? 2e-6 : 1e-12))) p->unit_rows = false;
      cudaFree(d_mx);
    }
  }
  // two smooths of the same covariate on the same knots (mean and log-sd smooths of a GAMLSS): one band
  p->shared_band = false;
  {
    const char* nos = getenv("LSLB_NO_SHARED_BAND");
    if (p->ns == 2 && desc->dtype == LSLB_F32 && e == cudaSuccess && !(nos && atoi(nos) != 0) &&
        p->blk[p->sp[0]].ncoef == p->blk[p->sp[1]].ncoef) {
      unsigned long long* d_cnt = nullptr;
      e = cudaMalloc(&d_cnt, sizeof(unsigned long long));
      if (e == cudaSuccess) e = cudaMemset(d_cnt, 0, sizeof(unsigned long long));
      if (e == cudaSuccess) {
        unsigned long long h = 1;
        band_diff_kernel<<<296, 256>>>(p->blk[p->sp[0]].index, p->blk[p->sp[1]].index,
                                       (const float*)p->blk[p->sp[0]].data, (const float*)p->blk[p->sp[1]].data, n, d_cnt);
        e = cudaMemcpy(&h, d_cnt, sizeof(h), cudaMemcpyDeviceToHost);
        p->shared_band = e == cudaSuccess && h == 0;
      }
      if (d_cnt) cudaFree(d_cnt);
    }
  }
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_cuda_error(e, "plan setup kernels");
    cudaFree(p->d_chunk_start);
    if (p->d_gstart) cudaFree(p->d_gstart);
    delete p;
    return LSLB_ERR_CUDA;
  }

  // kernel variant
  if (p->grp < 0 && p->nd == 0 && p->ns >= 1) {
    p->variant = VAR_BANDED;
    p->variant_name = banded_fast_eligible(p) ? "banded_tma_runlength"
                      : (msmooth_x2_eligible(p) ? "multi_smooth_x2" : "banded_runlength");
  } else if (p->ns == 0 && p->nd == 1 && p->blk[p->dn[0]].ncoef <= 16) {
    p->variant = VAR_DENSE_REG;
    p->variant_name = dgroup_x2_eligible(p) ? "dense_group_x2_tma"
                      : (tiny_eligible(p, 1) ? "dense_single_launch" : "dense_registers");
  } else {
    p->variant = VAR_GENERIC;
    p->variant_name = dense_x2_eligible(p) ? "dense_x2_register_sliced" : "generic_smem";
  }
  dense_tc_prepare(p);
  if (p->d_XT) p->variant_name = "dense_tcgen05_tf32x3";
  if (p->variant == VAR_BANDED) {
    if (msort_prepare(p) != LSLB_OK) {
      lslb_plan_destroy(p);
      return LSLB_ERR_CUDA;
    }
    if (msort_x2_eligible(p)) p->variant_name = "multi_smooth_chunk_sorted_x2";
  }
  if (p->variant == VAR_BANDED && desc->dtype == LSLB_F32 && n > 0) {
    const long long ntiles = (n + 127) / 128;
    for (int m = 0; m < p->ns && e == cudaSuccess; ++m) {
      e = cudaMalloc(&p->d_tw[m], sizeof(float) * 4 * ntiles);
      if (e == cudaSuccess)
        tile_weight_sums_kernel<<<(unsigned)ntiles, 128>>>((const float*)p->blk[p->sp[m]].data, p->blk[p->sp[m]].index, n,
                                                           p->d_tw[m]);
    }
    // IWLS of the Normal families runs the packed kernel (banded_iwls_x2.cu): weight products of every
    // smooth, 48 bytes per row (LSLB_NO_IWLS_PAIRS=1 skips them; the kernel then multiplies per chain)
    const char* nop = getenv("LSLB_NO_IWLS_PAIRS");
    if (!(nop && atoi(nop) != 0) && banded_fast_eligible(p) &&
        (p->fam == FAM_NORMAL_LS || p->fam == FAM_NORMAL_FIXED)) {
      for (int m = 0; m < p->ns && e == cudaSuccess; ++m) {
        if ((reinterpret_cast<uintptr_t>(p->blk[p->sp[m]].data) & 15) != 0) continue;
        e = cudaMalloc(&p->d_wp[m], sizeof(float) * 12 * (size_t)n);
        if (e == cudaSuccess)
          weight_products_kernel<<<(unsigned)((n + 255) / 256), 256>>>((const float*)p->blk[p->sp[m]].data, n,
                                                                        p->d_wp[m]);
      }
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      set_cuda_error(e, "tile weight sums");
      lslb_plan_destroy(p);
      return LSLB_ERR_CUDA;
    }
  }
  *out = p;
  return LSLB_OK;
}

extern "C" void lslb_plan_destroy(lslb_plan* p) {
  if (!p) return;
  if (p->d_chunk_start) cudaFree(p->d_chunk_start);
  if (p->d_gstart) cudaFree(p->d_gstart);
  for (int m = 0; m < LSLB_MAX_SPL; ++m)
    if (p->d_tw[m]) cudaFree(p->d_tw[m]);
  for (int m = 0; m < LSLB_MAX_SPL; ++m)
    if (p->d_wp[m]) cudaFree(p->d_wp[m]);
  lslb::dense_tc_release(p);
  lslb::msort_release(p);
  delete p;
}

extern "C" int lslb_plan_num_params(const lslb_plan* p) { return p ? p->P : -1; }
extern "C" int lslb_plan_num_aux(const lslb_plan* p) { return p ? p->A : -1; }
extern "C" int lslb_plan_dtype(const lslb_plan* p) { return p ? p->desc.dtype : -1; }
extern "C" int lslb_plan_block_ncoef(const lslb_plan* p, int b) {
  return (p && b >= 0 && b < p->desc.nblocks) ? p->blk[b].ncoef : -1;
}
extern "C" const char* lslb_plan_variant(const lslb_plan* p) { return p ? p->variant_name : ""; }
extern "C" size_t lslb_workspace_bytes(const lslb_plan* p, int C) {
  return p ? lslb::workspace_bytes(p, C) : 0;
}
