// XLA FFI handlers for the projector pair: the thin layer between XLA's custom-call ABI and the
// C ABI in include/scico_b200_xray.h.  NOT built by default: it needs the headers shipped inside
// jaxlib (`python -c "import jax.ffi; print(jax.ffi.include_dir())"`), and JAX is not installed
// in this image.  Build (where JAX is available):
//
//   g++ -O2 -fPIC -shared -std=c++17 -I$(python -c 'import jax.ffi;print(jax.ffi.include_dir())')
//       -I/usr/local/cuda/include -Iinclude scico_b200/csrc/xct_ffi.cc
//       -Lscico_b200 -lscico_b200_xray -Wl,-rpath,'$ORIGIN' -o scico_b200/libscico_b200_ffi.so
//
// Contract (XLA side): operands are read-only device buffers, result buffers are preallocated and
// uninitialised, everything is enqueued on the stream XLA passes, no host synchronisation, no
// allocation -- exactly what xct_forward / xct_adjoint guarantee.  The plan (geometry tables on
// the device, built once per operator by xct2d/3d_plan_create) is passed as an int64 attribute.
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "scico_b200_xray.h"
#include "xla/ffi/api/c_api.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

ffi::Error check(int rc) {
  if (rc == XCT_OK) return ffi::Error::Success();
  return ffi::Error(rc == XCT_ERR_INVALID ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal,
                    std::string("scico_b200: ") + xct_last_error());
}

ffi::Error ForwardImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> x, ffi::ResultBuffer<ffi::F32> y,
                       int64_t plan, int32_t batch) {
  xct_plan_info info;
  const xct_plan* pl = reinterpret_cast<const xct_plan*>(static_cast<intptr_t>(plan));
  if (ffi::Error e = check(xct_plan_get_info(pl, &info)); e.failure()) return e;
  if (static_cast<int64_t>(x.element_count()) != info.in_elems * batch ||
      static_cast<int64_t>(y->element_count()) != info.out_elems * batch)
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "scico_b200: buffer sizes do not match the plan");
  return check(xct_forward(pl, x.typed_data(), y->typed_data(), batch, stream));
}

ffi::Error AdjointImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> y, ffi::ResultBuffer<ffi::F32> x,
                       int64_t plan, int32_t batch) {
  xct_plan_info info;
  const xct_plan* pl = reinterpret_cast<const xct_plan*>(static_cast<intptr_t>(plan));
  if (ffi::Error e = check(xct_plan_get_info(pl, &info)); e.failure()) return e;
  if (static_cast<int64_t>(y.element_count()) != info.out_elems * batch ||
      static_cast<int64_t>(x->element_count()) != info.in_elems * batch)
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "scico_b200: buffer sizes do not match the plan");
  return check(xct_adjoint(pl, y.typed_data(), x->typed_data(), batch, stream));
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(XctForwardFfi, ForwardImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<int64_t>("plan")
                                  .Attr<int32_t>("batch"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(XctAdjointFfi, AdjointImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<int64_t>("plan")
                                  .Attr<int32_t>("batch"));
