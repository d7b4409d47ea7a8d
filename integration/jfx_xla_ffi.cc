// XLA FFI handlers that forward jax.ffi custom calls to the jfx C ABI (include/jfx.h).
//
// NOT BUILT IN THIS REPOSITORY'S IMAGE: it needs the jaxlib headers (xla/ffi/api/ffi.h, shipped inside
// the jaxlib wheel: `python -c "import jax.ffi; print(jax.ffi.include_dir())"`), which are absent here
// (no jax, no network).  It is the binding a jaxfun maintainer adds; see INTEGRATION.md.  Build:
//
//   g++ -O2 -fPIC -shared -std=c++17 -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") \
//       -Iinclude integration/jfx_xla_ffi.cc -Ljaxfun_b200 -ljfx -lcudart -o libjfx_xla.so
//
// Contract: buffers belong to XLA; scratch comes from XLA's ScratchAllocator; handlers are re-entrant;
// nothing in the execute stage synchronises the device, so the calls are CUDA-graph (command-buffer) compatible.
//
// Plans travel as a 64-bit REGISTRY KEY (jfx_registry_register, include/jfx.h), never as a pointer: the key is a hash of
// the plan descriptor and its table contents, so it is the same in every process and can sit in a serialised executable;
// the plan object itself is created per device in the handler's `initialize` stage (jfx_registry_acquire on the device
// XLA made current), which is what a program running on 8 devices under shard_map needs: one descriptor, eight plans.
// A key that was never registered in the process (an executable loaded from a compilation cache before jaxfun imported
// its spaces) fails loudly with the registry's message.
#include <cuda_runtime_api.h>

#include <cstdint>

#include "jfx.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

// ---- initialize stage: make sure the plan of this key exists on the current device (may allocate and synchronise) ------
static ffi::Error JfxInitImpl(int64_t plan_key) {
  const jfx_plan* plan = nullptr;
  if (jfx_registry_acquire(static_cast<uint64_t>(plan_key), &plan) != JFX_OK)
    return ffi::Error(ffi::ErrorCode::kFailedPrecondition, jfx_last_error());
  return ffi::Error::Success();
}
XLA_FFI_DEFINE_HANDLER(kJfxInit, JfxInitImpl, ffi::Ffi::Bind<ffi::ExecutionStage::kInitialize>().Attr<int64_t>("plan_key"));

// ---- execute stage: forward / backward / scalar_product / backward_primitive / apply: the plan says which -----------------
static ffi::Error JfxExecuteImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::AnyBuffer x,
                                 ffi::Result<ffi::AnyBuffer> y, int64_t plan_key) {
  const jfx_plan* plan = nullptr;
  if (jfx_registry_acquire(static_cast<uint64_t>(plan_key), &plan) != JFX_OK)   // a lookup: the plan exists since initialize
    return ffi::Error(ffi::ErrorCode::kFailedPrecondition, jfx_last_error());
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
XLA_FFI_DEFINE_HANDLER(kJfxExecute, JfxExecuteImpl,
                       ffi::Ffi::Bind()
                           .Ctx<ffi::PlatformStream<cudaStream_t>>()
                           .Ctx<ffi::ScratchAllocator>()
                           .Arg<ffi::AnyBuffer>()
                           .Ret<ffi::AnyBuffer>()
                           .Attr<int64_t>("plan_key"),
                       {xla::ffi::Traits::kCmdBufferCompatible});

// One exported symbol per op family; jax.ffi.register_ffi_target takes the bundle {"initialize": ..., "execute": ...}.
extern "C" XLA_FFI_Error* JfxExecuteInitialize(XLA_FFI_CallFrame* f) { return kJfxInit->Call(f); }
extern "C" XLA_FFI_Error* JfxExecute(XLA_FFI_CallFrame* f) { return kJfxExecute->Call(f); }

// ---- slab transform (sharding.py:43-105): phase 1 with the exchange fused into its last pass, then phase 2 --------------
// Inside shard_map every device calls this handler on its local block.  `recv` is this device's receive buffer and
// `peers` the table of all devices' receive-buffer addresses as mapped on THIS device (peer-mapped / symmetric memory,
// exchanged once at start-up — jax.experimental's multi-process utilities or NCCL's ncclCommWindowRegister give them).
// The cross-device ordering between the peer stores of phase 1 and the reads of phase 2 is an XLA-level barrier: the
// Python glue issues the two halves as two ffi_calls with `jax.lax.psum(0, axis)` (an all-reduce on a scalar) between
// them, so no handler ever blocks on another device.
static ffi::Error JfxSlabPhase1Impl(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::AnyBuffer x,
                                    ffi::Buffer<ffi::DataType::S64> peers, ffi::Result<ffi::AnyBuffer> token,
                                    int64_t plan_key, int64_t rank, int64_t split_axis) {
  const jfx_plan* plan = nullptr;
  if (jfx_registry_acquire(static_cast<uint64_t>(plan_key), &plan) != JFX_OK)
    return ffi::Error(ffi::ErrorCode::kFailedPrecondition, jfx_last_error());
  const int parts = static_cast<int>(peers.element_count());
  if (!jfx_plan_scatter_supported(plan, parts, static_cast<int>(split_axis)))
    return ffi::Error(ffi::ErrorCode::kUnimplemented, "jfx: plan has no scatter epilogue; use jfx_execute + lax.all_to_all");
  size_t ws_bytes = 0;
  jfx_plan_workspace_bytes(plan, &ws_bytes);
  void* ws = nullptr;
  if (ws_bytes) {
    auto got = scratch.Allocate(ws_bytes);
    if (!got.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "jfx: scratch allocation failed");
    ws = *got;
  }
  // the peer table is a small HOST-visible constant of the program (device pointers as int64): pinned by the Python glue
  void* const* peer_ptrs = reinterpret_cast<void* const*>(peers.typed_data());
  const int rc = jfx_execute_scatter(plan, stream, x.untyped_data(), peer_ptrs, parts, static_cast<int>(rank),
                                     static_cast<int>(split_axis), ws);
  (void)token;
  if (rc != JFX_OK) return ffi::Error(ffi::ErrorCode::kInternal, jfx_last_error());
  return ffi::Error::Success();
}
XLA_FFI_DEFINE_HANDLER(kJfxSlabPhase1, JfxSlabPhase1Impl,
                       ffi::Ffi::Bind()
                           .Ctx<ffi::PlatformStream<cudaStream_t>>()
                           .Ctx<ffi::ScratchAllocator>()
                           .Arg<ffi::AnyBuffer>()
                           .Arg<ffi::Buffer<ffi::DataType::S64>>()
                           .Ret<ffi::AnyBuffer>()
                           .Attr<int64_t>("plan_key")
                           .Attr<int64_t>("rank")
                           .Attr<int64_t>("split_axis"));
extern "C" XLA_FFI_Error* JfxSlabTransformPhase1(XLA_FFI_CallFrame* f) { return kJfxSlabPhase1->Call(f); }
// phase 2 is an ordinary JfxExecute on the receive buffer.

// ---- the whole slab transform as ONE handler (JfxSlabTransform) ------------------------------------------------
// `slab` is a jfx_slab created and bound by the Python glue at start-up (jfx_slab_create with slab_rank / slab_size,
// jfx_slab_bind with the peer-mapped receive buffers and flag pads of all devices); like the nonlinear objects it is looked
// up by a per-process integer id.  jfx_slab_execute enqueues phase 1 (peer stores from the contraction epilogue, or one
// strided peer copy per device), a device-side flag barrier and phase 2 on XLA's stream: no NCCL call, no host
// synchronisation, and the handler never blocks the host thread on another device (the wait happens on the GPU, bounded).
// Every device of the mesh must run the handler the same number of times — which shard_map guarantees.  Not marked
// command-buffer compatible: the object alternates between its two receive buffers on successive calls.
static ffi::Error JfxSlabTransformImpl(cudaStream_t stream, ffi::ScratchAllocator scratch, ffi::AnyBuffer x,
                                       ffi::Result<ffi::AnyBuffer> out, int64_t handle) {
  jfx_slab* slab = reinterpret_cast<jfx_slab*>(static_cast<intptr_t>(handle));
  size_t ws_bytes = 0;
  if (jfx_slab_sizes(slab, nullptr, nullptr, &ws_bytes, nullptr) != JFX_OK)
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, jfx_last_error());
  auto got = scratch.Allocate(ws_bytes ? ws_bytes : 1);
  if (!got.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "jfx: scratch allocation failed");
  const int rc = jfx_slab_execute(slab, stream, x.untyped_data(), out->untyped_data(), *got);
  if (rc != JFX_OK) return ffi::Error(ffi::ErrorCode::kInternal, jfx_last_error());
  return ffi::Error::Success();
}
XLA_FFI_DEFINE_HANDLER(kJfxSlabTransform, JfxSlabTransformImpl,
                       ffi::Ffi::Bind()
                           .Ctx<ffi::PlatformStream<cudaStream_t>>()
                           .Ctx<ffi::ScratchAllocator>()
                           .Arg<ffi::AnyBuffer>()
                           .Ret<ffi::AnyBuffer>()
                           .Attr<int64_t>("handle"));
extern "C" XLA_FFI_Error* JfxSlabTransform(XLA_FFI_CallFrame* f) { return kJfxSlabTransform->Call(f); }

// ---- nonlinear term ---------------------------------------------------------------------------------------------------
// jfx_nonlinear objects are created by the Python glue at trace time and looked up by an integer id it keeps per process
// (nonlinear descriptors hold leaf plan descriptors; a registry for them follows the plan registry's pattern).
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

// ---- wavenumber-batched banded solve (la/tpmatrix.py:982-1013: TPMatricesWavenumberSolver.solve) ---------------------------
// The jfx_banded object (assembled and factored on this device by jfx_banded_create when the solver object is built on the
// Python side, from W / P_data_stack / poly_offsets of tpmats_wavenumber_factor) travels as a per-process handle like the
// nonlinear objects.  `outer` / `inner` place the polynomial axis inside the array (outer * n * inner elements): the
// transposes of tpmatrix.py:963-977 do not exist here.  One launch, no allocation, no synchronisation.
static ffi::Error JfxBandedSolveImpl(cudaStream_t stream, ffi::AnyBuffer rhs, ffi::Result<ffi::AnyBuffer> out, int64_t handle,
                                     int64_t outer, int64_t inner) {
  const jfx_banded* b = reinterpret_cast<const jfx_banded*>(static_cast<intptr_t>(handle));
  const int rc = jfx_banded_solve(b, stream, rhs.untyped_data(), out->untyped_data(), outer, inner);
  if (rc != JFX_OK) return ffi::Error(rc == JFX_ERR_INVALID ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal, jfx_last_error());
  return ffi::Error::Success();
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(JfxBandedSolve, JfxBandedSolveImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Attr<int64_t>("handle")
                                  .Attr<int64_t>("outer")
                                  .Attr<int64_t>("inner"),
                              {xla::ffi::Traits::kCmdBufferCompatible});
