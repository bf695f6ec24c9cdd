// XLA FFI handlers for the projector pair: the thin layer between XLA's custom-call ABI and the operator
// registry of the C ABI (include/scico_b200_xray.h, xct_op_*).  NOT built by default: it needs the headers
// shipped inside jaxlib (`python -c "import jax.ffi; print(jax.ffi.include_dir())"`), and JAX is not installed
// in this image, so this file has never been compiled.  Everything with logic in it -- the per-device plan
// table, the batch rule, lifetime -- lives in xct_api.cu behind xct_op_plan / xct_op_apply, which ARE built and
// tested (tests/test_host.py, tests/test_gpu_op_registry.py); this file only unpacks XLA's arguments.
// Build (where JAX is available):
//
//   g++ -O2 -fPIC -shared -std=c++17 -I$(python -c 'import jax.ffi;print(jax.ffi.include_dir())')
//       -I/usr/local/cuda/include -Iinclude scico_b200/csrc/xct_ffi.cc
//       -Lscico_b200 -lscico_b200_xray -Wl,-rpath,'$ORIGIN' -o scico_b200/libscico_b200_ffi.so
//
// Contract (XLA side): operands are read-only device buffers, result buffers are preallocated and
// uninitialised, everything is enqueued on the stream XLA passes, no host synchronisation and no allocation
// in the execute stage -- what xct_forward / xct_adjoint guarantee.  The operator is an int64 ATTRIBUTE holding
// a registry id (not a pointer): the executable can run on any device of the process (the handler asks XLA for
// the device ordinal and uses that device's plan), and a released operator yields an error, not a dangling
// pointer.  The plan of a device is created in the INITIALIZE stage (once per executable and device), where XLA
// allows allocation.  The leading batch axes jax.vmap adds (vmap_method="expand_dims") are folded into the
// batch count by xct_op_apply.
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

// initialize stage: make sure the operator has a plan on the device this executable runs on
ffi::Error InitImpl(int32_t device, int64_t op) {
  const xct_plan* pl = nullptr;
  return check(xct_op_plan(op, device, &pl));
}

ffi::Error ForwardImpl(cudaStream_t stream, int32_t device, ffi::Buffer<ffi::F32> x, ffi::ResultBuffer<ffi::F32> y, int64_t op) {
  return check(xct_op_apply(op, device, 1, x.typed_data(), y->typed_data(), static_cast<int64_t>(x.element_count()), stream));
}

ffi::Error AdjointImpl(cudaStream_t stream, int32_t device, ffi::Buffer<ffi::F32> y, ffi::ResultBuffer<ffi::F32> x, int64_t op) {
  return check(xct_op_apply(op, device, 0, y.typed_data(), x->typed_data(), static_cast<int64_t>(y.element_count()), stream));
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(XctInitFfi, InitImpl,
                              ffi::Ffi::Bind<ffi::ExecutionStage::kInitialize>().Ctx<ffi::DeviceOrdinal>().Attr<int64_t>("op"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(XctForwardFfi, ForwardImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Ctx<ffi::DeviceOrdinal>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<int64_t>("op"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(XctAdjointFfi, AdjointImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Ctx<ffi::DeviceOrdinal>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<int64_t>("op"));
