// XLA FFI shim: re-exports the C ABI of include/liesel_b200.h as jax.ffi custom-call targets
// (`jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(lib.<symbol>), platform="CUDA")`, see
// INTEGRATION.md). It is the binding behind the reference's plug-in points: `ModelInterface.log_prob`
// (src/liesel/goose/types.py:72-102) under `jax.vmap` over chains (goose/engine.py:442-448), and
// `IWLSKernel(chol_info_fn=...)` (goose/iwls.py:130-145,241-243).
//
// Build. The XLA FFI headers ship inside jaxlib, which is not installable in this image, so
// build.sh compiles this file against the TEST STUB tests/stubs/xla/ffi/api/ffi.h into
// build/liblslb_xla_ffi_stub.so on every build (it cannot rot, and the tests drive the handlers
// through `lslb_stub_call` below: decoding, checks and the calls into the C ABI run for real; only
// XLA's binary call frame is not reproduced). With a JAX environment:
//   g++ -shared -fPIC -std=c++17 -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())")
//       -I include -I /usr/local/cuda/include xla_ffi_shim.cc -L. -lliesel_b200 -o liblslb_xla_ffi.so
//
// Operand layout = what `jax.vmap` over chains produces with vmap_method="broadcast_all":
// params [C, P], aux [C, A] (A may be 0) -> logp [C], grad [C, P]; IWLS: score [C, p], info and
// chol [C, p, p]. float32 and float64 (`jax_enable_x64`) targets; the dtype must match the plan's.
// The plan handle travels as an int64 attribute (the plan is immutable and outlives the jitted
// computation); scratch comes from XLA's ScratchAllocator, so a call allocates nothing itself and
// runs asynchronously on XLA's stream. Errors are returned as ffi::Error (never thrown); numerical
// failure stays NaN/Inf in the outputs, as in the reference.
#if __has_include("xla/ffi/api/ffi.h")
#include <cstdint>
#include <string>
#include <type_traits>

#if __has_include(<cuda_runtime_api.h>)
#include <cuda_runtime_api.h>
#else
typedef struct CUstream_st* cudaStream_t;
#endif

#include "xla/ffi/api/ffi.h"

#include "liesel_b200.h"

namespace ffi = xla::ffi;

namespace {

template <ffi::DataType DT>
using Native = typename std::conditional<DT == ffi::F32, float, double>::type;

ffi::Error Invalid(const std::string& m) { return ffi::Error(ffi::ErrorCode::kInvalidArgument, "liesel_b200: " + m); }

template <typename Dims>
bool dims_are(const Dims& d, std::initializer_list<int64_t> want) {
  if (d.size() != want.size()) return false;
  size_t i = 0;
  for (int64_t w : want)
    if (d[i++] != w) return false;
  return true;
}

// Checks shared by all targets; returns the chain count through *C.
template <ffi::DataType DT>
ffi::Error check_common(const lslb_plan* plan, const ffi::Buffer<DT>& params, const ffi::Buffer<DT>& aux, int* C) {
  if (plan == nullptr) return Invalid("null plan handle");
  const int want_dtype = DT == ffi::F32 ? LSLB_F32 : LSLB_F64;
  if (lslb_plan_dtype(plan) != want_dtype) return Invalid("operand dtype differs from the plan's dtype");
  const int P = lslb_plan_num_params(plan), A = lslb_plan_num_aux(plan);
  const auto pd = params.dimensions();
  if (pd.size() != 2 || pd[0] < 1 || pd[1] != P) return Invalid("params must be [C, P]");
  if (pd[0] > 0x7fffffff) return Invalid("too many chains");
  if (A > 0 && !dims_are(aux.dimensions(), {pd[0], (int64_t)A})) return Invalid("aux must be [C, A]");
  *C = static_cast<int>(pd[0]);
  return ffi::Error::Success();
}

template <ffi::DataType DT>
ffi::Error LogpGradImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, int64_t plan_handle,
                        ffi::Buffer<DT> params, ffi::Buffer<DT> aux, ffi::ResultBuffer<DT> logp,
                        ffi::ResultBuffer<DT> grad) {
  const lslb_plan* plan = reinterpret_cast<const lslb_plan*>(plan_handle);
  int C = 0;
  ffi::Error e = check_common<DT>(plan, params, aux, &C);
  if (e.failure()) return e;
  if (!dims_are(logp->dimensions(), {(int64_t)C})) return Invalid("logp must be [C]");
  if (!dims_are(grad->dimensions(), {(int64_t)C, (int64_t)lslb_plan_num_params(plan)})) return Invalid("grad must be [C, P]");
  const size_t ws_bytes = lslb_workspace_bytes(plan, C);
  auto ws = scratch.Allocate(ws_bytes);
  if (!ws.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "liesel_b200: workspace");
  const Native<DT>* ax = lslb_plan_num_aux(plan) > 0 ? aux.typed_data() : nullptr;
  int st;
  if constexpr (DT == ffi::F32)
    st = lslb_logp_grad_f32(plan, C, params.typed_data(), ax, logp->typed_data(), grad->typed_data(), 0u, *ws,
                            ws_bytes, reinterpret_cast<lslb_stream_t>(stream));
  else
    st = lslb_logp_grad_f64(plan, C, params.typed_data(), ax, logp->typed_data(), grad->typed_data(), 0u, *ws,
                            ws_bytes, reinterpret_cast<lslb_stream_t>(stream));
  if (st != LSLB_OK) return ffi::Error(ffi::ErrorCode::kInternal, lslb_status_string(st));
  return ffi::Error::Success();
}

// value only (what `mh_step` needs, goose/mh.py:48-50): grad == NULL in the C ABI
template <ffi::DataType DT>
ffi::Error LogpImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, int64_t plan_handle, ffi::Buffer<DT> params,
                    ffi::Buffer<DT> aux, ffi::ResultBuffer<DT> logp) {
  const lslb_plan* plan = reinterpret_cast<const lslb_plan*>(plan_handle);
  int C = 0;
  ffi::Error e = check_common<DT>(plan, params, aux, &C);
  if (e.failure()) return e;
  if (!dims_are(logp->dimensions(), {(int64_t)C})) return Invalid("logp must be [C]");
  const size_t ws_bytes = lslb_workspace_bytes(plan, C);
  auto ws = scratch.Allocate(ws_bytes);
  if (!ws.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "liesel_b200: workspace");
  const Native<DT>* ax = lslb_plan_num_aux(plan) > 0 ? aux.typed_data() : nullptr;
  int st;
  if constexpr (DT == ffi::F32)
    st = lslb_logp_grad_f32(plan, C, params.typed_data(), ax, logp->typed_data(), nullptr, 0u, *ws, ws_bytes,
                            reinterpret_cast<lslb_stream_t>(stream));
  else
    st = lslb_logp_grad_f64(plan, C, params.typed_data(), ax, logp->typed_data(), nullptr, 0u, *ws, ws_bytes,
                            reinterpret_cast<lslb_stream_t>(stream));
  if (st != LSLB_OK) return ffi::Error(ffi::ErrorCode::kInternal, lslb_status_string(st));
  return ffi::Error::Success();
}

// score + observed information + Cholesky factor (with the reference's ridge) of one block
template <ffi::DataType DT>
ffi::Error IwlsImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, int64_t plan_handle, int64_t block,
                    ffi::Buffer<DT> params, ffi::Buffer<DT> aux, ffi::ResultBuffer<DT> score,
                    ffi::ResultBuffer<DT> info, ffi::ResultBuffer<DT> chol) {
  const lslb_plan* plan = reinterpret_cast<const lslb_plan*>(plan_handle);
  int C = 0;
  ffi::Error e = check_common<DT>(plan, params, aux, &C);
  if (e.failure()) return e;
  const int p = lslb_plan_block_ncoef(plan, static_cast<int>(block));
  if (p < 1) return Invalid("block index out of range");
  if (!dims_are(score->dimensions(), {(int64_t)C, (int64_t)p})) return Invalid("score must be [C, p]");
  if (!dims_are(info->dimensions(), {(int64_t)C, (int64_t)p, (int64_t)p}) ||
      !dims_are(chol->dimensions(), {(int64_t)C, (int64_t)p, (int64_t)p}))
    return Invalid("info and chol must be [C, p, p]");
  const size_t ws_bytes = lslb_workspace_bytes(plan, C);
  auto ws = scratch.Allocate(ws_bytes);
  if (!ws.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "liesel_b200: workspace");
  const Native<DT>* ax = lslb_plan_num_aux(plan) > 0 ? aux.typed_data() : nullptr;
  int st;
  if constexpr (DT == ffi::F32)
    st = lslb_iwls_f32(plan, C, static_cast<int>(block), params.typed_data(), ax, score->typed_data(),
                       info->typed_data(), chol->typed_data(), 0u, *ws, ws_bytes,
                       reinterpret_cast<lslb_stream_t>(stream));
  else
    st = lslb_iwls_f64(plan, C, static_cast<int>(block), params.typed_data(), ax, score->typed_data(),
                       info->typed_data(), chol->typed_data(), 0u, *ws, ws_bytes,
                       reinterpret_cast<lslb_stream_t>(stream));
  if (st != LSLB_OK) return ffi::Error(ffi::ErrorCode::kInternal, lslb_status_string(st));
  return ffi::Error::Success();
}

// score + DIAGONAL of the observed information of the random-intercept block (lslb_iwls_group_*): what an
// IWLSKernel whose position is that block obtains from grad / -jacfwd(grad) (goose/iwls.py:206-243), without
// the G x G matrix
template <ffi::DataType DT>
ffi::Error IwlsGroupImpl(cudaStream_t stream, int64_t plan_handle, ffi::Buffer<DT> params, ffi::Buffer<DT> aux,
                         ffi::ResultBuffer<DT> score, ffi::ResultBuffer<DT> info_diag) {
  const lslb_plan* plan = reinterpret_cast<const lslb_plan*>(plan_handle);
  int C = 0;
  ffi::Error e = check_common<DT>(plan, params, aux, &C);
  if (e.failure()) return e;
  const auto sd = score->dimensions();
  if (sd.size() != 2 || sd[0] != C || sd[1] < 1) return Invalid("score must be [C, G]");
  if (!dims_are(info_diag->dimensions(), {sd[0], sd[1]})) return Invalid("info_diag must be [C, G]");
  const Native<DT>* ax = lslb_plan_num_aux(plan) > 0 ? aux.typed_data() : nullptr;
  int st;
  if constexpr (DT == ffi::F32)
    st = lslb_iwls_group_f32(plan, C, params.typed_data(), ax, score->typed_data(), info_diag->typed_data(), nullptr,
                             0u, reinterpret_cast<lslb_stream_t>(stream));
  else
    st = lslb_iwls_group_f64(plan, C, params.typed_data(), ax, score->typed_data(), info_diag->typed_data(), nullptr,
                             0u, reinterpret_cast<lslb_stream_t>(stream));
  if (st != LSLB_OK) return ffi::Error(ffi::ErrorCode::kInternal, lslb_status_string(st));
  return ffi::Error::Success();
}

// beta^T K beta of one penalised block (tau2_gibbs_kernel, model/distreg.py:291)
template <ffi::DataType DT>
ffi::Error QuadformImpl(cudaStream_t stream, int64_t plan_handle, int64_t block, ffi::Buffer<DT> params,
                        ffi::ResultBuffer<DT> out) {
  const lslb_plan* plan = reinterpret_cast<const lslb_plan*>(plan_handle);
  if (plan == nullptr) return Invalid("null plan handle");
  if (lslb_plan_dtype(plan) != (DT == ffi::F32 ? LSLB_F32 : LSLB_F64)) return Invalid("operand dtype differs from the plan's dtype");
  const auto pd = params.dimensions();
  if (pd.size() != 2 || pd[0] < 1 || pd[1] != lslb_plan_num_params(plan)) return Invalid("params must be [C, P]");
  if (!dims_are(out->dimensions(), {pd[0]})) return Invalid("out must be [C]");
  int st;
  if constexpr (DT == ffi::F32)
    st = lslb_quadform_f32(plan, (int)pd[0], (int)block, params.typed_data(), out->typed_data(),
                           reinterpret_cast<lslb_stream_t>(stream));
  else
    st = lslb_quadform_f64(plan, (int)pd[0], (int)block, params.typed_data(), out->typed_data(),
                           reinterpret_cast<lslb_stream_t>(stream));
  if (st != LSLB_OK) return ffi::Error(ffi::ErrorCode::kInternal, lslb_status_string(st));
  return ffi::Error::Success();
}

}  // namespace

#define LSLB_BIND_LOGP_GRAD(DT)                     \
  ffi::Ffi::Bind()                                  \
      .Ctx<ffi::PlatformStream<cudaStream_t>>()     \
      .Ctx<ffi::ScratchAllocator>()                 \
      .Attr<int64_t>("plan")                        \
      .Arg<ffi::Buffer<DT>>()                       \
      .Arg<ffi::Buffer<DT>>()                       \
      .Ret<ffi::Buffer<DT>>()                       \
      .Ret<ffi::Buffer<DT>>()
#define LSLB_BIND_LOGP(DT)                          \
  ffi::Ffi::Bind()                                  \
      .Ctx<ffi::PlatformStream<cudaStream_t>>()     \
      .Ctx<ffi::ScratchAllocator>()                 \
      .Attr<int64_t>("plan")                        \
      .Arg<ffi::Buffer<DT>>()                       \
      .Arg<ffi::Buffer<DT>>()                       \
      .Ret<ffi::Buffer<DT>>()
#define LSLB_BIND_IWLS(DT)                          \
  ffi::Ffi::Bind()                                  \
      .Ctx<ffi::PlatformStream<cudaStream_t>>()     \
      .Ctx<ffi::ScratchAllocator>()                 \
      .Attr<int64_t>("plan")                        \
      .Attr<int64_t>("block")                       \
      .Arg<ffi::Buffer<DT>>()                       \
      .Arg<ffi::Buffer<DT>>()                       \
      .Ret<ffi::Buffer<DT>>()                       \
      .Ret<ffi::Buffer<DT>>()                       \
      .Ret<ffi::Buffer<DT>>()
#define LSLB_BIND_IWLS_GROUP(DT)                    \
  ffi::Ffi::Bind()                                  \
      .Ctx<ffi::PlatformStream<cudaStream_t>>()     \
      .Attr<int64_t>("plan")                        \
      .Arg<ffi::Buffer<DT>>()                       \
      .Arg<ffi::Buffer<DT>>()                       \
      .Ret<ffi::Buffer<DT>>()                       \
      .Ret<ffi::Buffer<DT>>()
#define LSLB_BIND_QUADFORM(DT)                      \
  ffi::Ffi::Bind()                                  \
      .Ctx<ffi::PlatformStream<cudaStream_t>>()     \
      .Attr<int64_t>("plan")                        \
      .Attr<int64_t>("block")                       \
      .Arg<ffi::Buffer<DT>>()                       \
      .Ret<ffi::Buffer<DT>>()

XLA_FFI_DEFINE_HANDLER_SYMBOL(LslbLogpGrad, LogpGradImpl<ffi::F32>, LSLB_BIND_LOGP_GRAD(ffi::F32));
XLA_FFI_DEFINE_HANDLER_SYMBOL(LslbLogpGradF64, LogpGradImpl<ffi::F64>, LSLB_BIND_LOGP_GRAD(ffi::F64));
XLA_FFI_DEFINE_HANDLER_SYMBOL(LslbLogp, LogpImpl<ffi::F32>, LSLB_BIND_LOGP(ffi::F32));
XLA_FFI_DEFINE_HANDLER_SYMBOL(LslbLogpF64, LogpImpl<ffi::F64>, LSLB_BIND_LOGP(ffi::F64));
XLA_FFI_DEFINE_HANDLER_SYMBOL(LslbIwls, IwlsImpl<ffi::F32>, LSLB_BIND_IWLS(ffi::F32));
XLA_FFI_DEFINE_HANDLER_SYMBOL(LslbIwlsF64, IwlsImpl<ffi::F64>, LSLB_BIND_IWLS(ffi::F64));
XLA_FFI_DEFINE_HANDLER_SYMBOL(LslbIwlsGroup, IwlsGroupImpl<ffi::F32>, LSLB_BIND_IWLS_GROUP(ffi::F32));
XLA_FFI_DEFINE_HANDLER_SYMBOL(LslbIwlsGroupF64, IwlsGroupImpl<ffi::F64>, LSLB_BIND_IWLS_GROUP(ffi::F64));
XLA_FFI_DEFINE_HANDLER_SYMBOL(LslbQuadform, QuadformImpl<ffi::F32>, LSLB_BIND_QUADFORM(ffi::F32));
XLA_FFI_DEFINE_HANDLER_SYMBOL(LslbQuadformF64, QuadformImpl<ffi::F64>, LSLB_BIND_QUADFORM(ffi::F64));

#ifdef LSLB_XLA_FFI_STUB
// Test driver (stub builds only): builds the stub's call frame from plain C arguments and invokes a
// handler symbol. Buffers: nargs operands then nrets results; dtypes[i] in {11 = F32, 12 = F64};
// dims are concatenated, ranks[i] entries each. Scratch comes from cudaMalloc and is released after
// a stream synchronise. Returns 0 on success, else the ffi::ErrorCode, message in errbuf.
#include <cuda_runtime_api.h>
#include <cstring>
#include <vector>
extern "C" int lslb_stub_call(XLA_FFI_Error* (*handler)(XLA_FFI_CallFrame*), void* stream, int64_t plan,
                              int64_t block, int nargs, int nrets, const int* dtypes, void* const* datas,
                              const int* ranks, const int64_t* dims, char* errbuf, int errlen) {
  xla::ffi::StubCallFrame f;
  f.stream = stream;
  std::vector<void*> scratch;
  f.alloc = [&](size_t n) -> void* {
    void* p = nullptr;
    if (cudaMalloc(&p, n ? n : 1) != cudaSuccess) return nullptr;
    scratch.push_back(p);
    return p;
  };
  f.attrs = {{"plan", plan}, {"block", block}};
  size_t off = 0;
  for (int i = 0; i < nargs + nrets; ++i) {
    xla::ffi::StubBuffer b{static_cast<xla::ffi::DataType>(dtypes[i]), datas[i],
                           std::vector<int64_t>(dims + off, dims + off + ranks[i])};
    off += ranks[i];
    (i < nargs ? f.args : f.rets).push_back(std::move(b));
  }
  handler(&f);
  cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
  for (void* p : scratch) cudaFree(p);
  if (errbuf && errlen > 0) {
    strncpy(errbuf, f.error.message().c_str(), (size_t)errlen - 1);
    errbuf[errlen - 1] = 0;
  }
  return static_cast<int>(f.error.code());
}
#endif  // LSLB_XLA_FFI_STUB
#endif  // __has_include("xla/ffi/api/ffi.h")
