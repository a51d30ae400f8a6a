// Row sharding over the GPUs of one box (BASELINE config 4): every rank evaluates ALL
// chains on ITS rows, and the per-rank partial sums logp [C], grad [C x P] (IWLS: score, info) are
// summed across ranks. Two transports behind the same C ABI:
//
//  * peer memory over NVLink/NVSwitch (lslb_peer_*): every rank owns one cudaMalloc'ed exchange
//    buffer that all other ranks map through CUDA IPC. The finalize kernel of the evaluation writes
//    its partial results STRAIGHT into that buffer (no staging copy), and one reduce kernel
//    signals "epoch e published" into every peer's flag word, waits for the W flags of its own
//    rank, then sums the W partials through P2P loads in RANK ORDER — every rank obtains the
//    bitwise identical sum, with no library call and no host synchronisation on the path.
//    Two slots alternate by epoch: a rank can only publish epoch e+2 after all peers signalled
//    e+1, i.e. after they finished reading epoch e, so slot reuse needs no second barrier.
//    A failure is fatal for the WHOLE group: a rank whose wait times out (or whose evaluation
//    could not be launched) writes the poison value into its flag word on every peer; a rank that
//    reads poison, or a flag two epochs ahead of its own (impossible in a healthy group), fails
//    the same way. The error word is sticky: every later call of a failed rank publishes poison
//    and returns NaN, until all ranks call lslb_peer_reset (a collective, bracketed by barriers).
//  * NCCL (lslb_*_nccl_*): `ncclAllReduce(sum)` on the caller's communicator and stream, resolved
//    from the process's libnccl.so.2 at first use (dlopen; the library is not a link dependency).
//
// The reference has no multi-device path (its only parallelism is jax.vmap over chains,
// src/liesel/goose/engine.py:442-448); what is summed here is the row sum of
// `_reduced_sum` (src/liesel/model/model.py:50-53) and its gradient.
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include <new>
#include <type_traits>

#include "common.cuh"

struct lslb_peer {
  int rank, world;
  size_t max_elems;    // capacity of one slot in elements of 8 bytes
  size_t slot_bytes;
  unsigned epoch;      // host mirror: calls issued so far
  char* local;         // this rank's exchange buffer
  char* mapped[LSLB_MAX_PEERS];  // every rank's buffer as seen from here (mapped[rank] == local)
  bool connected;
  bool poisoned;       // host mirror: a launch failed between peer_begin and peer_finish
};

namespace lslb {

constexpr size_t PEER_FLAGS_OFF = 0;    // unsigned flags[W]: flags[r] = last epoch rank r published
constexpr size_t PEER_ERR_OFF = 256;    // unsigned: set when a wait timed out
constexpr size_t PEER_DATA_OFF = 512;
constexpr unsigned PEER_POISON = 0xFFFFFFFFu;  // never a valid epoch

struct PeerDev {
  const char* slot[LSLB_MAX_PEERS];  // slot of this epoch inside every rank's buffer
  unsigned* flag_at[LSLB_MAX_PEERS]; // &flags[me] inside every rank's buffer
  unsigned* my_flags;
  unsigned* err;
  int rank, world;
  unsigned epoch;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// One region of the slot: out[i] = sum_r slot_r[off + i], r ascending.
__device__ __forceinline__ void acc_add(float& a, const float& v) { a += v; }
__device__ __forceinline__ void acc_add(double& a, const double& v) { a += v; }
__device__ __forceinline__ void acc_add(float4& a, const float4& v) { a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
__device__ __forceinline__ void acc_add(double2& a, const double2& v) { a.x += v.x; a.y += v.y; }

template <typename V>
__device__ __forceinline__ void reduce_region(const PeerDev& pd, size_t byte_off, V* __restrict__ out,
                                              size_t count, size_t tid, size_t nthreads) {
  for (size_t i = tid; i < count; i += nthreads) {
    V acc;
    // batches of 8 ranks: all loads of a batch are in flight together (one NVLink round trip),
    // the additions then run in rank order
    for (int r0 = 0; r0 < pd.world; r0 += 8) {
      V v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (r0 + j < pd.world) v[j] = __ldcv(reinterpret_cast<const V*>(pd.slot[r0 + j] + byte_off) + i);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (r0 + j < pd.world) {
          if (r0 + j == 0) acc = v[0];
          else acc_add(acc, v[j]);
        }
    }
    out[i] = acc;
  }
}

// Slot layout: region A (na elements, padded to 16 bytes) then region B (nb elements).
template <typename T>
__global__ void __launch_bounds__(256)
peer_reduce_kernel(const __grid_constant__ PeerDev pd, T* __restrict__ out_a, size_t na,
                   T* __restrict__ out_b, size_t nb, size_t off_b, int vec_b) {
  // a rank that failed before stays failed (sticky error word) and tells its peers so
  __shared__ int failed;
  if (threadIdx.x == 0) failed = *reinterpret_cast<volatile unsigned*>(pd.err) != 0u;
  __syncthreads();
  const bool dead = failed != 0;
  // publish: the partials were written by the preceding kernel on this stream
  if (blockIdx.x == 0 && (int)threadIdx.x < pd.world) {
    __threadfence_system();
    st_release_sys(pd.flag_at[threadIdx.x], dead ? PEER_POISON : pd.epoch);
  }
  // wait until every rank published this epoch. The wait is bounded (~10 s of SM clock): a rank that
  // never arrives must not hang the GPU. A failed call is LOUD and fatal for the group: the error
  // flag is set, poison goes to every peer and every output of this block becomes NaN (the
  // reference's convention for a failed evaluation).
  if (!dead && (int)threadIdx.x < pd.world) {
    const long long t0 = clock64();
    for (;;) {
      const unsigned f = ld_acquire_sys(pd.my_flags + threadIdx.x);
      const int ahead = (int)(f - pd.epoch);
      if (f == PEER_POISON || ahead >= 2) { failed = 1; break; }  // peer failed / slot already overwritten
      if (ahead >= 0) break;
      if (clock64() - t0 > 20000000000LL) { failed = 1; break; }
      __nanosleep(64);
    }
  }
  __syncthreads();
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nth = (size_t)gridDim.x * blockDim.x;
  if (failed) {
    if (!dead && (int)threadIdx.x < pd.world) {
      if (threadIdx.x == 0) *pd.err = 1u;
      __threadfence_system();
      st_release_sys(pd.flag_at[threadIdx.x], PEER_POISON);
    }
    const T nan = (T)__int_as_float(0x7fc00000);
    for (size_t i = tid; i < na; i += nth) out_a[i] = nan;
    for (size_t i = tid; i < nb; i += nth) out_b[i] = nan;
    return;
  }
  reduce_region<T>(pd, 0, out_a, na, tid, nth);
  if (vec_b) {
    using V = typename std::conditional<sizeof(T) == 4, float4, double2>::type;
    reduce_region<V>(pd, off_b, reinterpret_cast<V*>(out_b), nb / (16 / sizeof(T)), tid, nth);
  } else {
    reduce_region<T>(pd, off_b, out_b, nb, tid, nth);
  }
}

// Marks this rank failed and poisons its flag word on every peer (lslb_peer_abort and the error
// returns between peer_begin and peer_finish).
__global__ void peer_poison_kernel(const __grid_constant__ PeerDev pd) {
  if ((int)threadIdx.x < pd.world) {
    if (threadIdx.x == 0) *pd.err = 1u;
    __threadfence_system();
    st_release_sys(pd.flag_at[threadIdx.x], PEER_POISON);
  }
}

static void fill_peer_dev(const lslb_peer* pr, unsigned e, PeerDev& pd) {
  memset(&pd, 0, sizeof(pd));
  for (int r = 0; r < pr->world; ++r) {
    pd.slot[r] = pr->mapped[r] + PEER_DATA_OFF + (size_t)(e & 1u) * pr->slot_bytes;
    pd.flag_at[r] = reinterpret_cast<unsigned*>(pr->mapped[r] + PEER_FLAGS_OFF) + pr->rank;
  }
  pd.my_flags = reinterpret_cast<unsigned*>(pr->local + PEER_FLAGS_OFF);
  pd.err = reinterpret_cast<unsigned*>(pr->local + PEER_ERR_OFF);
  pd.rank = pr->rank;
  pd.world = pr->world;
  pd.epoch = e;
}

static int peer_poison(lslb_peer* pr, cudaStream_t st) {
  if (!pr || !pr->connected) return LSLB_ERR_INVALID;
  pr->poisoned = true;
  PeerDev pd;
  fill_peer_dev(pr, pr->epoch, pd);
  peer_poison_kernel<<<1, 32, 0, st>>>(pd);
  LSLB_LAUNCH_CHECK();
  return LSLB_OK;
}

static size_t region_b_offset(size_t na, size_t elem) { return (na * elem + 15) / 16 * 16; }

// Pointers into the local slot of the NEXT epoch for the producing kernels.
template <typename T>
static int peer_begin(lslb_peer* pr, size_t na, size_t nb, cudaStream_t st, T** a, T** b) {
  if (!pr || !pr->connected) return LSLB_ERR_INVALID;
  if (pr->poisoned) {
    set_comm_error("peer group failed earlier (launch error or lslb_peer_abort); call lslb_peer_reset on all ranks");
    return LSLB_ERR_COMM;
  }
  if (region_b_offset(na, sizeof(T)) + nb * sizeof(T) > pr->slot_bytes) return LSLB_ERR_WORKSPACE;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  LSLB_CUDA(cudaStreamIsCapturing(st, &cs));
  if (cs != cudaStreamCaptureStatusNone) return LSLB_ERR_UNSUPPORTED;  // the epoch is a launch argument
  const unsigned e = pr->epoch + 1;
  char* slot = pr->local + PEER_DATA_OFF + (size_t)(e & 1u) * pr->slot_bytes;
  *a = reinterpret_cast<T*>(slot);
  *b = reinterpret_cast<T*>(slot + region_b_offset(na, sizeof(T)));
  return LSLB_OK;
}

template <typename T>
static int peer_finish(lslb_peer* pr, T* out_a, size_t na, T* out_b, size_t nb, cudaStream_t st) {
  const unsigned e = ++pr->epoch;
  PeerDev pd;
  fill_peer_dev(pr, e, pd);
  const size_t per = 16 / sizeof(T);
  const int vec_b = nb % per == 0 && reinterpret_cast<uintptr_t>(out_b) % 16 == 0;
  const size_t work = na + (vec_b ? nb / per : nb);
  unsigned blocks = (unsigned)((work + 255) / 256);
  if (blocks < 1) blocks = 1;
  if (blocks > 592) blocks = 592;  // 4 x 148: enough loads in flight to cover the NVLink round trip
  peer_reduce_kernel<T><<<blocks, 256, 0, st>>>(pd, out_a, na, out_b, nb, region_b_offset(na, sizeof(T)), vec_b);
  LSLB_LAUNCH_CHECK();
  return LSLB_OK;
}

template <typename T>
static int logp_grad_peer(const lslb_plan* plan, int C, const T* params, const T* aux, T* logp, T* grad,
                          unsigned flags, void* ws, size_t ws_bytes, lslb_peer* pr, cudaStream_t st) {
  if (!plan || !logp || C < 1) return LSLB_ERR_INVALID;
  const size_t nb = grad ? (size_t)C * plan->P : 0;
  T *pa = nullptr, *pb = nullptr;
  int s = peer_begin<T>(pr, (size_t)C, nb, st, &pa, &pb);
  if (s != LSLB_OK) return s;
  if (pr->rank != 0) flags |= LSLB_SKIP_PRIOR;
  s = launch_logp_grad<T>(plan, C, params, aux, pa, grad ? pb : nullptr, flags, ws, ws_bytes, st);
  if (s != LSLB_OK) {  // the peers are (or will be) waiting for this epoch: fail the group, do not desynchronise
    peer_poison(pr, st);
    return s;
  }
  return peer_finish<T>(pr, logp, (size_t)C, grad, nb, st);
}

template <typename T>
static int iwls_peer(const lslb_plan* plan, int C, int block, const T* params, const T* aux, T* score,
                     T* info, T* chol, unsigned flags, void* ws, size_t ws_bytes, lslb_peer* pr,
                     cudaStream_t st) {
  if (!plan || !score || !info || C < 1 || block < 0 || block >= plan->desc.nblocks) return LSLB_ERR_INVALID;
  const size_t p = plan->blk[block].ncoef;
  const size_t na = (size_t)C * p, nb = (size_t)C * p * p;
  T *pa = nullptr, *pb = nullptr;
  int s = peer_begin<T>(pr, na, nb, st, &pa, &pb);
  if (s != LSLB_OK) return s;
  if (pr->rank != 0) flags |= LSLB_SKIP_PRIOR;
  s = launch_iwls<T>(plan, C, block, params, aux, pa, pb, nullptr, flags, ws, ws_bytes, st);
  if (s != LSLB_OK) {
    peer_poison(pr, st);
    return s;
  }
  s = peer_finish<T>(pr, score, na, info, nb, st);
  if (s != LSLB_OK || !chol) return s;
  return launch_chol<T>(C, (int)p, info, chol, st);
}

// ---------------------------------------------------------------------------------------
// NCCL, resolved at run time from the process's own libnccl.so.2
// ---------------------------------------------------------------------------------------
struct NcclId { char bytes[128]; };
struct NcclApi {
  int (*GetUniqueId)(NcclId*);
  int (*CommInitRank)(void**, int, NcclId, int);
  int (*CommDestroy)(void*);
  int (*CommUserRank)(void*, int*);
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
  int (*GroupStart)();
  int (*GroupEnd)();
  const char* (*GetErrorString)(int);
  bool ok;
};

static NcclApi* nccl_api() {
  static NcclApi api = [] {
    NcclApi a;
    memset(&a, 0, sizeof(a));
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return a;
#define LSLB_SYM(name) *(void**)(&a.name) = dlsym(h, "nccl" #name)
    LSLB_SYM(GetUniqueId); LSLB_SYM(CommInitRank); LSLB_SYM(CommDestroy); LSLB_SYM(CommUserRank);
    LSLB_SYM(AllReduce); LSLB_SYM(GroupStart); LSLB_SYM(GroupEnd); LSLB_SYM(GetErrorString);
#undef LSLB_SYM
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.CommUserRank && a.AllReduce &&
           a.GroupStart && a.GroupEnd && a.GetErrorString;
    return a;
  }();
  return &api;
}

static int nccl_fail(NcclApi* n, int rc, const char* where) {
  char msg[200];
  snprintf(msg, sizeof(msg), "%s: %s", where, n->GetErrorString ? n->GetErrorString(rc) : "?");
  set_comm_error(msg);
  return LSLB_ERR_COMM;
}

template <typename T>
static int nccl_sum2(void* comm, T* a, size_t na, T* b, size_t nb, cudaStream_t st) {
  NcclApi* n = nccl_api();
  const int dt = sizeof(T) == 4 ? 7 /* ncclFloat32 */ : 8 /* ncclFloat64 */;
  int rc = n->GroupStart();
  if (rc) return nccl_fail(n, rc, "ncclGroupStart");
  rc = n->AllReduce(a, a, na, dt, 0 /* ncclSum */, comm, st);
  if (!rc && b && nb) rc = n->AllReduce(b, b, nb, dt, 0, comm, st);
  const int rc2 = n->GroupEnd();
  if (rc) return nccl_fail(n, rc, "ncclAllReduce");
  if (rc2) return nccl_fail(n, rc2, "ncclGroupEnd");
  return LSLB_OK;
}

template <typename T>
static int logp_grad_nccl(const lslb_plan* plan, int C, const T* params, const T* aux, T* logp, T* grad,
                          unsigned flags, void* ws, size_t ws_bytes, void* comm, cudaStream_t st) {
  NcclApi* n = nccl_api();
  if (!n->ok) {
    set_comm_error("libnccl.so.2 not found in this process");
    return LSLB_ERR_COMM;
  }
  if (!plan || !comm) return LSLB_ERR_INVALID;
  int rank = 0, rc = n->CommUserRank(comm, &rank);
  if (rc) return nccl_fail(n, rc, "ncclCommUserRank");
  if (rank != 0) flags |= LSLB_SKIP_PRIOR;
  int s = launch_logp_grad<T>(plan, C, params, aux, logp, grad, flags, ws, ws_bytes, st);
  if (s != LSLB_OK) return s;
  return nccl_sum2<T>(comm, logp, (size_t)C, grad, grad ? (size_t)C * plan->P : 0, st);
}

template <typename T>
static int iwls_nccl(const lslb_plan* plan, int C, int block, const T* params, const T* aux, T* score,
                     T* info, T* chol, unsigned flags, void* ws, size_t ws_bytes, void* comm,
                     cudaStream_t st) {
  NcclApi* n = nccl_api();
  if (!n->ok) {
    set_comm_error("libnccl.so.2 not found in this process");
    return LSLB_ERR_COMM;
  }
  if (!plan || !comm || block < 0 || block >= plan->desc.nblocks) return LSLB_ERR_INVALID;
  int rank = 0, rc = n->CommUserRank(comm, &rank);
  if (rc) return nccl_fail(n, rc, "ncclCommUserRank");
  if (rank != 0) flags |= LSLB_SKIP_PRIOR;
  int s = launch_iwls<T>(plan, C, block, params, aux, score, info, nullptr, flags, ws, ws_bytes, st);
  if (s != LSLB_OK) return s;
  const size_t p = plan->blk[block].ncoef;
  s = nccl_sum2<T>(comm, score, (size_t)C * p, info, (size_t)C * p * p, st);
  if (s != LSLB_OK || !chol) return s;
  return launch_chol<T>(C, (int)p, info, chol, st);
}

}  // namespace lslb

using namespace lslb;

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
extern "C" int lslb_peer_create(int rank, int world, size_t max_elems, lslb_peer** out, void* handle) {
  if (!out || !handle || world < 1 || world > LSLB_MAX_PEERS || rank < 0 || rank >= world || max_elems < 1)
    return LSLB_ERR_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == LSLB_PEER_HANDLE_BYTES, "handle size");
  *out = nullptr;
  lslb_peer* pr = new (std::nothrow) lslb_peer();
  if (!pr) return LSLB_ERR_INVALID;
  memset(pr, 0, sizeof(*pr));
  pr->rank = rank;
  pr->world = world;
  pr->max_elems = max_elems;
  pr->slot_bytes = (max_elems * 8 + 16 + 255) / 256 * 256;
  const size_t bytes = PEER_DATA_OFF + 2 * pr->slot_bytes;
  cudaError_t e = cudaMalloc(&pr->local, bytes);
  if (e == cudaSuccess) e = cudaMemset(pr->local, 0, bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, pr->local);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    set_cuda_error(e, "lslb_peer_create");
    if (pr->local) cudaFree(pr->local);
    delete pr;
    return LSLB_ERR_CUDA;
  }
  memcpy(handle, &h, sizeof(h));
  pr->mapped[rank] = pr->local;
  pr->connected = world == 1;
  *out = pr;
  return LSLB_OK;
}

extern "C" int lslb_peer_connect(lslb_peer* pr, const void* handles) {
  if (!pr || !handles) return LSLB_ERR_INVALID;
  if (pr->connected) return LSLB_OK;
  for (int r = 0; r < pr->world; ++r) {
    if (r == pr->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)r * LSLB_PEER_HANDLE_BYTES, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      set_cuda_error(e, "cudaIpcOpenMemHandle (are the GPUs P2P-connected?)");
      for (int q = 0; q < r; ++q)
        if (q != pr->rank && pr->mapped[q]) {
          cudaIpcCloseMemHandle(pr->mapped[q]);
          pr->mapped[q] = nullptr;
        }
      return LSLB_ERR_CUDA;
    }
    pr->mapped[r] = (char*)p;
  }
  pr->connected = true;
  return LSLB_OK;
}

extern "C" void lslb_peer_destroy(lslb_peer* pr) {
  if (!pr) return;
  for (int r = 0; r < pr->world; ++r)
    if (r != pr->rank && pr->mapped[r]) cudaIpcCloseMemHandle(pr->mapped[r]);
  if (pr->local) cudaFree(pr->local);
  delete pr;
}

extern "C" int lslb_peer_timed_out(lslb_peer* pr, int* timed_out) {
  if (!pr || !timed_out) return LSLB_ERR_INVALID;
  unsigned v = 0;
  LSLB_CUDA(cudaMemcpy(&v, pr->local + PEER_ERR_OFF, sizeof(v), cudaMemcpyDeviceToHost));
  *timed_out = (int)v;
  return LSLB_OK;
}

extern "C" int lslb_peer_abort(lslb_peer* pr, lslb_stream_t s) { return peer_poison(pr, (cudaStream_t)s); }

extern "C" int lslb_peer_reset(lslb_peer* pr) {
  if (!pr || !pr->connected) return LSLB_ERR_INVALID;
  LSLB_CUDA(cudaDeviceSynchronize());
  LSLB_CUDA(cudaMemset(pr->local, 0, PEER_DATA_OFF));  // flag words and the error word
  LSLB_CUDA(cudaDeviceSynchronize());
  pr->epoch = 0;
  pr->poisoned = false;
  return LSLB_OK;
}

#define LSLB_S(s) ((cudaStream_t)(s))
extern "C" int lslb_logp_grad_peer_f32(const lslb_plan* plan, int C, const float* params, const float* aux,
                                       float* logp, float* grad, uint32_t flags, void* ws, size_t ws_bytes,
                                       lslb_peer* peer, lslb_stream_t s) {
  return logp_grad_peer<float>(plan, C, params, aux, logp, grad, flags, ws, ws_bytes, peer, LSLB_S(s));
}
extern "C" int lslb_logp_grad_peer_f64(const lslb_plan* plan, int C, const double* params, const double* aux,
                                       double* logp, double* grad, uint32_t flags, void* ws, size_t ws_bytes,
                                       lslb_peer* peer, lslb_stream_t s) {
  return logp_grad_peer<double>(plan, C, params, aux, logp, grad, flags, ws, ws_bytes, peer, LSLB_S(s));
}
extern "C" int lslb_iwls_peer_f32(const lslb_plan* plan, int C, int block, const float* params,
                                  const float* aux, float* score, float* info, float* chol, uint32_t flags,
                                  void* ws, size_t ws_bytes, lslb_peer* peer, lslb_stream_t s) {
  return iwls_peer<float>(plan, C, block, params, aux, score, info, chol, flags, ws, ws_bytes, peer, LSLB_S(s));
}
extern "C" int lslb_iwls_peer_f64(const lslb_plan* plan, int C, int block, const double* params,
                                  const double* aux, double* score, double* info, double* chol,
                                  uint32_t flags, void* ws, size_t ws_bytes, lslb_peer* peer, lslb_stream_t s) {
  return iwls_peer<double>(plan, C, block, params, aux, score, info, chol, flags, ws, ws_bytes, peer, LSLB_S(s));
}

extern "C" int lslb_nccl_unique_id(void* id) {
  NcclApi* n = nccl_api();
  if (!id) return LSLB_ERR_INVALID;
  if (!n->ok) {
    set_comm_error("libnccl.so.2 not found in this process");
    return LSLB_ERR_COMM;
  }
  NcclId u;
  int rc = n->GetUniqueId(&u);
  if (rc) return nccl_fail(n, rc, "ncclGetUniqueId");
  memcpy(id, &u, sizeof(u));
  return LSLB_OK;
}
extern "C" int lslb_nccl_comm_create(int rank, int world, const void* id, void** comm) {
  NcclApi* n = nccl_api();
  if (!id || !comm || world < 1 || rank < 0 || rank >= world) return LSLB_ERR_INVALID;
  if (!n->ok) {
    set_comm_error("libnccl.so.2 not found in this process");
    return LSLB_ERR_COMM;
  }
  NcclId u;
  memcpy(&u, id, sizeof(u));
  int rc = n->CommInitRank(comm, world, u, rank);
  if (rc) return nccl_fail(n, rc, "ncclCommInitRank");
  return LSLB_OK;
}
extern "C" void lslb_nccl_comm_destroy(void* comm) {
  NcclApi* n = nccl_api();
  if (comm && n->ok) n->CommDestroy(comm);
}
extern "C" int lslb_logp_grad_nccl_f32(const lslb_plan* plan, int C, const float* params, const float* aux,
                                       float* logp, float* grad, uint32_t flags, void* ws, size_t ws_bytes,
                                       void* comm, lslb_stream_t s) {
  return logp_grad_nccl<float>(plan, C, params, aux, logp, grad, flags, ws, ws_bytes, comm, LSLB_S(s));
}
extern "C" int lslb_logp_grad_nccl_f64(const lslb_plan* plan, int C, const double* params, const double* aux,
                                       double* logp, double* grad, uint32_t flags, void* ws, size_t ws_bytes,
                                       void* comm, lslb_stream_t s) {
  return logp_grad_nccl<double>(plan, C, params, aux, logp, grad, flags, ws, ws_bytes, comm, LSLB_S(s));
}
extern "C" int lslb_iwls_nccl_f32(const lslb_plan* plan, int C, int block, const float* params,
                                  const float* aux, float* score, float* info, float* chol, uint32_t flags,
                                  void* ws, size_t ws_bytes, void* comm, lslb_stream_t s) {
  return iwls_nccl<float>(plan, C, block, params, aux, score, info, chol, flags, ws, ws_bytes, comm, LSLB_S(s));
}
extern "C" int lslb_iwls_nccl_f64(const lslb_plan* plan, int C, int block, const double* params,
                                  const double* aux, double* score, double* info, double* chol,
                                  uint32_t flags, void* ws, size_t ws_bytes, void* comm, lslb_stream_t s) {
  return iwls_nccl<double>(plan, C, block, params, aux, score, info, chol, flags, ws, ws_bytes, comm, LSLB_S(s));
}
