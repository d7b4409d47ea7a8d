// Shared declarations of the jfx engine (internal; the public ABI is include/jfx.h).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/jfx.h"

namespace jfx {

// ---- error plumbing -------------------------------------------------------------------
void set_error(const char* fmt, ...);
const char* get_error();

#define JFX_CUDA_OK(expr)                                                                 \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      jfx::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,    \
                     __LINE__);                                                           \
      return JFX_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

#define JFX_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      jfx::set_error(__VA_ARGS__);    \
      return (code);                  \
    }                                 \
  } while (0)

inline bool dtype_is_complex(int dt) { return dt == JFX_C64 || dt == JFX_C128; }
inline bool dtype_is_double(int dt) { return dt == JFX_F64 || dt == JFX_C128; }
inline size_t dtype_size(int dt) {
  switch (dt) {
    case JFX_F32: return 4;
    case JFX_F64: return 8;
    case JFX_C64: return 8;
    default: return 16;
  }
}

// ---- one pass along one axis ------------------------------------------------------------
// The array is viewed as [outer, n_in, inner] -> [outer, n_out, inner] (row-major).
struct AxisGeom {
  int64_t outer;
  int64_t inner;
  int n_in;
  int n_out;
};

// Dense table apply: out[o, r, i] = sum_c T[r, c] * in[o, c, i].
// table: device pointer, row-major [n_out, n_in], real (table_complex = 0) or complex.
// data dtype as jfx_dtype.  Picks the DMMA kernel when eligible, the generic one otherwise.
int launch_table_apply(cudaStream_t s, const AxisGeom& g, int dtype, const void* table,
                       bool table_complex, const void* in, void* out, int* used_dmma);
bool table_apply_uses_dmma(const AxisGeom& g, int dtype, bool table_complex);

// Fast transforms (kernels_fft.cu).  kind: see FastKind.
enum FastKind {
  FAST_CHEB_BACKWARD = 0,   // c[N] -> u[n]   (Chebyshev.py:225-241), optional chebder^k prologue
  FAST_CHEB_FORWARD = 1,    // u[n] -> c[N]   (Chebyshev.py:243-260)
  FAST_CHEB_SCALAR = 2,     // u[n] -> c[N]   (Chebyshev.py:262-279)
  FAST_FOURIER_BACKWARD = 3,  // c[N] -> u[n] (Fourier.py:126-148), optional (i m)^k prologue
  FAST_FOURIER_FORWARD = 4,   // u[n] -> c[N] (Fourier.py:165-180)
  FAST_FOURIER_SCALAR = 5     // u[n] -> c[N] (Fourier.py:150-163)
};
struct FastParams {
  int kind;
  int n_modes;   // N
  int n_quad;    // n (transform length)
  int deriv;     // k
  double domain_factor;
  int reverse;   // walk the tiles last-to-first: the pass starts on what the previous pass wrote last (still in L2)
};
bool fast_available(int basis, int n, int dtype);
bool fast_geometry_ok(const AxisGeom& g, int dtype);
// twiddle tables etc. are owned by a FastTables object created at plan time
struct FastTables;
int fast_tables_create(const FastParams& p, int dtype, FastTables** out);
void fast_tables_destroy(FastTables* t);
int launch_fast_axis(cudaStream_t s, const AxisGeom& g, int dtype, const FastParams& p,
                     const FastTables* t, const void* in, void* out);

// Pointwise / layout kernels (kernels_pointwise.cu)
int launch_slab_pack(cudaStream_t s, const void* in, void* out, const int64_t* shape, int ndim,
                     int split_axis, int parts, int dtype);
int launch_slab_unpack(cudaStream_t s, const void* in, void* out, const int64_t* shape_out,
                       int ndim, int concat_axis, int parts, int dtype);
int launch_point_contract(cudaStream_t s, const void* y, const void* w, void* out, int64_t outer, int n, int64_t P, int dtype,
                          int w_is_complex);
int launch_axpby_diag(cudaStream_t s, int n_terms, const void* const* coeff, const double* alpha,
                      const void* const* x, void* out, int64_t n, int dtype, int coeff_is_complex);
// Polynomial normal form of a pointwise program:  sum_t coeff_t * prod_f x_f,  x_f = leaf or conj(leaf).
// Every nonlinear term of the reference's examples and tests has this form (u u_x, (u+u_x)^2,
// 6u(u_x^2+u_y^2)+3u^2(u_xx+u_yy), |u|^2 u, ...); it evaluates with three fixed registers instead of an
// operand stack.  program_to_poly() derives it from the postfix program by symbolic execution.
#define JFX_POLY_MAX_TERMS 16
#define JFX_POLY_MAX_FACTORS 8
struct PolyTerm {
  double cre, cim;
  int nf;
  unsigned char fac[JFX_POLY_MAX_FACTORS];   // leaf index | 0x80 = conjugated
  int pad_;
};
struct PolyProgram {
  int n_terms;   // 0 = not a polynomial: use the stack machine
  int pad_;
  PolyTerm t[JFX_POLY_MAX_TERMS];
};

struct PointwiseProgram {
  int n_instr;
  jfx_pw_instr instr[JFX_MAX_PROGRAM];
  int n_consts;
  double consts[32][2];
  int n_leaves;
};
int validate_program(const PointwiseProgram& prog, const void* const* statics);
int program_depth(const PointwiseProgram& prog);
bool program_to_poly(const PointwiseProgram& prog, PolyProgram* out);
int launch_pointwise(cudaStream_t s, const PointwiseProgram& prog, const void* const* leaves,
                     const void* const* statics, void* out, int64_t n, int dtype);

// Row-fused nonlinear term along a Fourier last axis (kernels_fused.cu)
struct FusedRowArgs {
  const void* src[JFX_MAX_LEAVES];     // coefficient rows per leaf group: [rows, n_coeff] complex
  const void* mult[JFX_MAX_LEAVES];    // per-leaf spectral multiplier over the coefficient index, or null
  const void* statics[JFX_MAX_LEAVES]; // mesh-sampled statics: [rows, N] complex
  int leaf_group[JFX_MAX_LEAVES];
  int n_leaves;
  void* out;                           // [rows, n_out] complex
  void* scratch;                       // leaf lines of the resident CTAs (bytes: see launch_fused_rows query)
  const void* tw;                      // exp(-2 pi i m / N)
  long long rows;
  int n_coeff, n_out;
  double scale;
  int n_instr;
  int depth;                           // operand-stack depth of the program
  PolyProgram poly;                    // polynomial normal form (n_terms = 0: none)
  jfx_pw_instr instr[JFX_MAX_PROGRAM];
  double consts[32][2];
};
int launch_fused_rows(cudaStream_t s, int dtype, int n, const FusedRowArgs& a, size_t* scratch_query);

int calibrate_dmma(cudaStream_t s, int iters, double* tflops);
int calibrate_dfma(cudaStream_t s, int iters, double* tflops);

}  // namespace jfx
