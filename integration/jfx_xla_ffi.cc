// XLA FFI handlers that forward jax.ffi custom calls to the jfx C ABI (include/jfx.h).
//
// NOT BUILT IN THIS REPOSITORY'S IMAGE: it needs the jaxlib headers (xla/ffi/api/ffi.h, shipped inside
// the jaxlib wheel: `python -c "import jax.ffi; print(jax.ffi.include_dir())"`), which are absent here
// (no jax, no network).  It is the binding a jaxfun maintainer adds; see INTEGRATION.md.  Build:
//
//   g++ -O2 -fPIC -shared -std=c++17 -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") \
//       -Iinclude integration/jfx_xla_ffi.cc -Ljaxfun_b200 -ljfx -lcudart -o libjfx_xla.so
//
// Contract: buffers belong to XLA; scratch comes from XLA's ScratchAllocator; handlers
// are re-entrant (plans are immutable, created at trace time in Python and passed as an int64 attribute);
// nothing synchronises the device, so the calls are CUDA-graph (command-buffer) compatible.
#include <cuda_runtime_api.h>

#include <cstdint>

#include "jfx.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error JfxExecuteImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::AnyBuffer x,
                                 ffi::Result<ffi::AnyBuffer> y, int64_t plan_handle) {
  const jfx_plan* plan = reinterpret_cast<const jfx_plan*>(static_cast<intptr_t>(plan_handle));
  size_t ws_bytes = 0;
  if (jfx_plan_workspace_bytes(plan, &ws_bytes) != JFX_OK)
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, jfx_last_error());
  void* ws = nullptr;
  if (ws_bytes) {
    auto got = scratch.Allocate(ws_bytes);
    if (!got.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "jfx: scratch allocation failed");
    ws = *got;
  }
  const int rc = jfx_execute(plan, stream, x.untyped_data(), y->untyped_data(), ws);
  if (rc != JFX_OK) return ffi::Error(ffi::ErrorCode::kInternal, jfx_last_error());
  return ffi::Error::Success();
}

// forward / backward / scalar_product / backward_primitive / evaluate: one handler, the plan says which
XLA_FFI_DEFINE_HANDLER_SYMBOL(JfxExecute, JfxExecuteImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Ctx<ffi::ScratchAllocator>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Attr<int64_t>("plan"),
                              {xla::ffi::Traits::kCmdBufferCompatible});

static ffi::Error JfxNonlinearImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::AnyBuffer uh,
                                   ffi::Result<ffi::AnyBuffer> out, int64_t handle) {
  const jfx_nonlinear* nl = reinterpret_cast<const jfx_nonlinear*>(static_cast<intptr_t>(handle));
  size_t ws_bytes = 0;
  if (jfx_nonlinear_workspace_bytes(nl, &ws_bytes) != JFX_OK)
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, jfx_last_error());
  auto got = scratch.Allocate(ws_bytes ? ws_bytes : 1);
  if (!got.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "jfx: scratch allocation failed");
  const int rc = jfx_nonlinear_execute(nl, stream, uh.untyped_data(), out->untyped_data(), *got);
  if (rc != JFX_OK) return ffi::Error(ffi::ErrorCode::kInternal, jfx_last_error());
  return ffi::Error::Success();
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(JfxNonlinear, JfxNonlinearImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Ctx<ffi::ScratchAllocator>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Attr<int64_t>("nonlinear"),
                              {xla::ffi::Traits::kCmdBufferCompatible});
