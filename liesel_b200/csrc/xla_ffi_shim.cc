// XLA FFI shim: re-exports the C ABI of include/liesel_b200.h as jax.ffi custom-call targets.
//
// NOT compiled by build.sh: the XLA FFI headers (xla/ffi/api/ffi.h) ship inside jaxlib, which is not
// installed in this image (SURVEY "environment" table), so this file is unexercised here. With a
// JAX environment:  g++ -shared -fPIC -std=c++17 -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") \
//                       -I include xla_ffi_shim.cc -L. -lliesel_b200 -o liblslb_xla_ffi.so
//
// Operand layout matches what `jax.vmap` over chains produces with vmap_method="broadcast_all":
// params [C, P] f32, aux [C, A] f32 -> logp [C], grad [C, P]. The plan handle travels as an int64
// attribute (the plan is immutable and outlives the jitted computation); scratch comes from XLA's
// ScratchAllocator so the call allocates nothing itself.
#if __has_include("xla/ffi/api/ffi.h")
#include "xla/ffi/api/ffi.h"

#include "liesel_b200.h"

namespace ffi = xla::ffi;

static ffi::Error LogpGradImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, int64_t plan_handle,
                               ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F32> aux,
                               ffi::ResultBuffer<ffi::F32> logp, ffi::ResultBuffer<ffi::F32> grad) {
  const lslb_plan* plan = reinterpret_cast<const lslb_plan*>(plan_handle);
  const int C = static_cast<int>(params.dimensions()[0]);
  const size_t ws_bytes = lslb_workspace_bytes(plan, C);
  auto ws = scratch.Allocate(ws_bytes);
  if (!ws.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "lslb workspace");
  const int st = lslb_logp_grad_f32(plan, C, params.typed_data(), aux.typed_data(), logp->typed_data(),
                                    grad->typed_data(), 0u, *ws, ws_bytes,
                                    reinterpret_cast<lslb_stream_t>(stream));
  if (st != LSLB_OK) return ffi::Error(ffi::ErrorCode::kInternal, lslb_status_string(st));
  return ffi::Error::Success();
}

static ffi::Error LogpImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, int64_t plan_handle,
                           ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F32> aux,
                           ffi::ResultBuffer<ffi::F32> logp) {
  const lslb_plan* plan = reinterpret_cast<const lslb_plan*>(plan_handle);
  const int C = static_cast<int>(params.dimensions()[0]);
  const size_t ws_bytes = lslb_workspace_bytes(plan, C);
  auto ws = scratch.Allocate(ws_bytes);
  if (!ws.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "lslb workspace");
  const int st = lslb_logp_grad_f32(plan, C, params.typed_data(), aux.typed_data(), logp->typed_data(),
                                    nullptr, 0u, *ws, ws_bytes, reinterpret_cast<lslb_stream_t>(stream));
  if (st != LSLB_OK) return ffi::Error(ffi::ErrorCode::kInternal, lslb_status_string(st));
  return ffi::Error::Success();
}

static ffi::Error IwlsImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, int64_t plan_handle,
                           int64_t block, ffi::Buffer<ffi::F32> params, ffi::Buffer<ffi::F32> aux,
                           ffi::ResultBuffer<ffi::F32> score, ffi::ResultBuffer<ffi::F32> info,
                           ffi::ResultBuffer<ffi::F32> chol) {
  const lslb_plan* plan = reinterpret_cast<const lslb_plan*>(plan_handle);
  const int C = static_cast<int>(params.dimensions()[0]);
  const size_t ws_bytes = lslb_workspace_bytes(plan, C);
  auto ws = scratch.Allocate(ws_bytes);
  if (!ws.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "lslb workspace");
  const int st = lslb_iwls_f32(plan, C, static_cast<int>(block), params.typed_data(), aux.typed_data(),
                               score->typed_data(), info->typed_data(), chol->typed_data(), 0u, *ws,
                               ws_bytes, reinterpret_cast<lslb_stream_t>(stream));
  if (st != LSLB_OK) return ffi::Error(ffi::ErrorCode::kInternal, lslb_status_string(st));
  return ffi::Error::Success();
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(LslbLogpGrad, LogpGradImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Ctx<ffi::ScratchAllocator>()
                                  .Attr<int64_t>("plan")
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(LslbLogp, LogpImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Ctx<ffi::ScratchAllocator>()
                                  .Attr<int64_t>("plan")
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(LslbIwls, IwlsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Ctx<ffi::ScratchAllocator>()
                                  .Attr<int64_t>("plan")
                                  .Attr<int64_t>("block")
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());
#endif  // __has_include("xla/ffi/api/ffi.h")
