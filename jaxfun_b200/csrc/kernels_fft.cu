// Fast transform kernels (Chebyshev DCT, Fourier FFT).  Placeholder until the radix kernels land:
// reports "no fast path" so every axis is served by the dense table kernels.
#include "jfx_common.h"

namespace jfx {

struct FastTables { int dummy; };

bool fast_available(int, int, int) { return false; }
int fast_tables_create(const FastParams&, int, FastTables** out) {
  *out = nullptr;
  set_error("no fast transform kernel for this size");
  return JFX_ERR_UNSUPPORTED;
}
void fast_tables_destroy(FastTables* t) { delete t; }
int launch_fast_axis(cudaStream_t, const AxisGeom&, int, const FastParams&, const FastTables*,
                     const void*, void*) {
  set_error("no fast transform kernel for this size");
  return JFX_ERR_UNSUPPORTED;
}

}  // namespace jfx
