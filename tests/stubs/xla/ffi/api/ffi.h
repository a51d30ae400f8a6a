// TEST STUB — NOT the XLA header. A small, self-contained stand-in for the public C++ surface of
// jaxlib's `xla/ffi/api/ffi.h` that liesel_b200/csrc/xla_ffi_shim.cc uses (Ffi::Bind() with
// Ctx/Attr/Arg/Ret, Buffer<dtype>, ResultBuffer, ScratchAllocator, PlatformStream, Error,
// XLA_FFI_DEFINE_HANDLER_SYMBOL). jaxlib is not installable in this image, so this stub lets the shim
// be COMPILED on every build (it cannot rot) and DRIVEN from the tests through the stub's own call
// frame (`lslb_stub_call`, below): argument decoding, dtype/shape checks and the calls into the C
// ABI run for real; only XLA's binary call-frame layout is not reproduced. Written from the
// documented API, for tests only — the product shim is compiled against jaxlib's real header.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <optional>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#define LSLB_XLA_FFI_STUB 1

namespace xla {
namespace ffi {

enum class DataType : uint8_t { INVALID = 0, S32 = 4, S64 = 5, F32 = 11, F64 = 12 };
inline constexpr DataType S32 = DataType::S32;
inline constexpr DataType S64 = DataType::S64;
inline constexpr DataType F32 = DataType::F32;
inline constexpr DataType F64 = DataType::F64;

template <DataType dt> struct NativeTypeOf;
template <> struct NativeTypeOf<DataType::F32> { using type = float; };
template <> struct NativeTypeOf<DataType::F64> { using type = double; };
template <> struct NativeTypeOf<DataType::S32> { using type = int32_t; };
template <> struct NativeTypeOf<DataType::S64> { using type = int64_t; };

template <typename T>
class Span {
 public:
  Span(const T* d, size_t n) : d_(d), n_(n) {}
  size_t size() const { return n_; }
  const T& operator[](size_t i) const { return d_[i]; }
  const T* begin() const { return d_; }
  const T* end() const { return d_ + n_; }

 private:
  const T* d_;
  size_t n_;
};

enum class ErrorCode : uint8_t { kOk = 0, kInvalidArgument = 3, kResourceExhausted = 8, kUnimplemented = 12, kInternal = 13 };

class Error {
 public:
  Error() = default;
  Error(ErrorCode c, std::string m) : code_(c), msg_(std::move(m)) {}
  static Error Success() { return Error(); }
  static Error InvalidArgument(std::string m) { return Error(ErrorCode::kInvalidArgument, std::move(m)); }
  static Error Internal(std::string m) { return Error(ErrorCode::kInternal, std::move(m)); }
  bool success() const { return code_ == ErrorCode::kOk; }
  bool failure() const { return !success(); }
  ErrorCode code() const { return code_; }
  const std::string& message() const { return msg_; }

 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string msg_;
};

// ---- stub call frame ------------------------------------------------------------------------
struct StubBuffer {
  DataType dtype;
  void* data;
  std::vector<int64_t> dims;
};
struct StubCallFrame {
  void* stream = nullptr;
  std::function<void*(size_t)> alloc;  // scratch allocator of the "runtime"
  std::vector<std::pair<std::string, int64_t>> attrs;
  std::vector<StubBuffer> args, rets;
  Error error;
};

template <DataType dt>
class Buffer {
 public:
  using T = typename NativeTypeOf<dt>::type;
  explicit Buffer(const StubBuffer* b) : b_(b) {}
  T* typed_data() const { return static_cast<T*>(b_->data); }
  void* untyped_data() const { return b_->data; }
  Span<int64_t> dimensions() const { return Span<int64_t>(b_->dims.data(), b_->dims.size()); }
  size_t element_count() const {
    size_t n = 1;
    for (int64_t d : b_->dims) n *= (size_t)d;
    return n;
  }
  DataType element_type() const { return dt; }

 private:
  const StubBuffer* b_;
};

template <typename T>
class Result {
 public:
  explicit Result(T v) : v_(v) {}
  T& operator*() { return v_; }
  T* operator->() { return &v_; }

 private:
  T v_;
};
template <DataType dt>
using ResultBuffer = Result<Buffer<dt>>;

class ScratchAllocator {
 public:
  explicit ScratchAllocator(StubCallFrame* f) : f_(f) {}
  std::optional<void*> Allocate(size_t size, size_t /*alignment*/ = 1) {
    void* p = f_->alloc ? f_->alloc(size) : nullptr;
    if (!p) return std::nullopt;
    return p;
  }

 private:
  StubCallFrame* f_;
};

template <typename S>
struct PlatformStream {};

// ---- binding ----------------------------------------------------------------------------------
namespace internal {
template <typename T> struct CtxTag {};
template <typename T> struct AttrTag { std::string name; };
template <typename T> struct ArgTag {};
template <typename T> struct RetTag {};

struct Cursor {
  StubCallFrame* f;
  size_t arg = 0, ret = 0;
  bool ok = true;
  std::string why;
};

template <typename S>
S decode(Cursor& c, const CtxTag<PlatformStream<S>>&) { return reinterpret_cast<S>(c.f->stream); }
inline ScratchAllocator decode(Cursor& c, const CtxTag<ScratchAllocator>&) { return ScratchAllocator(c.f); }
inline int64_t decode(Cursor& c, const AttrTag<int64_t>& t) {
  for (auto& kv : c.f->attrs)
    if (kv.first == t.name) return kv.second;
  c.ok = false;
  c.why = "missing attribute " + t.name;
  return 0;
}
template <DataType dt>
Buffer<dt> decode(Cursor& c, const ArgTag<Buffer<dt>>&) {
  static StubBuffer empty{dt, nullptr, {}};
  if (c.arg >= c.f->args.size()) { c.ok = false; c.why = "too few operands"; return Buffer<dt>(&empty); }
  const StubBuffer* b = &c.f->args[c.arg++];
  if (b->dtype != dt) { c.ok = false; c.why = "operand dtype mismatch"; }
  return Buffer<dt>(b);
}
template <DataType dt>
Result<Buffer<dt>> decode(Cursor& c, const RetTag<Buffer<dt>>&) {
  static StubBuffer empty{dt, nullptr, {}};
  if (c.ret >= c.f->rets.size()) { c.ok = false; c.why = "too few results"; return Result<Buffer<dt>>(Buffer<dt>(&empty)); }
  const StubBuffer* b = &c.f->rets[c.ret++];
  if (b->dtype != dt) { c.ok = false; c.why = "result dtype mismatch"; }
  return Result<Buffer<dt>>(Buffer<dt>(b));
}
}  // namespace internal

using Handler = std::function<Error(StubCallFrame*)>;

template <typename... Tags>
class Binding {
 public:
  explicit Binding(std::tuple<Tags...> t) : tags_(std::move(t)) {}
  template <typename T> auto Ctx() && { return ext(internal::CtxTag<T>{}); }
  template <typename T> auto Attr(std::string name) && { return ext(internal::AttrTag<T>{std::move(name)}); }
  template <typename T> auto Arg() && { return ext(internal::ArgTag<T>{}); }
  template <typename T> auto Ret() && { return ext(internal::RetTag<T>{}); }

  template <typename Fn>
  std::unique_ptr<Handler> To(Fn fn) && {
    auto tags = tags_;
    return std::make_unique<Handler>([tags, fn](StubCallFrame* f) -> Error {
      internal::Cursor c{f};
      // braced initialisation: decoded strictly left to right
      auto vals = std::apply([&](const auto&... t) { return std::tuple{internal::decode(c, t)...}; }, tags);
      if (!c.ok) return Error(ErrorCode::kInvalidArgument, c.why);
      if (c.arg != f->args.size() || c.ret != f->rets.size())
        return Error(ErrorCode::kInvalidArgument, "operand/result count mismatch");
      return std::apply(fn, std::move(vals));
    });
  }

 private:
  template <typename Tag>
  auto ext(Tag t) {
    return Binding<Tags..., Tag>(std::tuple_cat(std::move(tags_), std::make_tuple(std::move(t))));
  }
  std::tuple<Tags...> tags_;
};

struct Ffi {
  static Binding<> Bind() { return Binding<>(std::tuple<>{}); }
};

}  // namespace ffi
}  // namespace xla

struct XLA_FFI_Error;
using XLA_FFI_CallFrame = xla::ffi::StubCallFrame;

// In the stub a handler reports its error through the frame and returns NULL on success.
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(fn, impl, binding)                                   \
  extern "C" XLA_FFI_Error* fn(XLA_FFI_CallFrame* call_frame) {                            \
    static auto* handler = (binding).To(impl).release();                                   \
    call_frame->error = (*handler)(call_frame);                                            \
    return call_frame->error.success() ? nullptr : reinterpret_cast<XLA_FFI_Error*>(call_frame); \
  }
