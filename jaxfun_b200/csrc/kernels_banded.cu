// Wavenumber-batched banded systems  B_s x_s = b_s  (s = 0 .. n_sys-1)  of Fourier x polynomial tensor-product problems:
// the device side of `TPMatricesWavenumberSolver` / `tpmats_wavenumber_factor` (la/tpmatrix.py:590-1014, 1236-1354 of the
// reference) and of the no-pivot banded LU it factors with (`_lu_banded_no_pivot_kernel`, la/diamatrix.py:1937-1973).
//
// Every Fourier wavenumber combination s owns one banded matrix  B_s = sum_t W[t, s] * P_t  of the polynomial axis
// (tpmatrix.py:1306-1347).  The reference assembles B as [n_sys, n_diags, n], factors it with a vmapped scan and solves with
// two vmapped scans over n (forward elimination, back substitution) after transposing the right-hand side to [n_sys, n].
//
// Here:
//  * storage is [band row][column j][system s] with s FASTEST: one thread owns one system and walks j, so every load of a
//    matrix entry is a coalesced 32-system row segment, for assembly, factorisation and both sweeps;
//  * assembly happens on the device from the separable form (W: n_terms x n_sys, P: n_terms x n_diags x n): the host never
//    materialises the n_sys matrices;
//  * the right-hand side stays where it is: the array is addressed as [outer, n, inner] (system s = o * inner + i), so
//    there are no transposes, and the solve may be in place;
//  * one launch does both sweeps; the recurrence window (the last p / q unknowns) lives in registers, and the loads of a
//    chunk of U steps (right-hand side and matrix entries: independent of the recurrence) are issued together before the
//    chunk's dependent arithmetic, so a thread has U * (1 + p) loads in flight instead of one;
//  * two kernels: `banded_solve_kernel` (one thread per system, any layout, window widths 2 / 4 / 8 or the generic read-back
//    variant for wider bands) and `banded_solve_rows_kernel` (polynomial axis last, bandwidth <= 4, >= 18 944 systems: one warp
//    moves [32 rows x 8 steps] tiles through shared memory so that the global accesses are row-contiguous).
// The per-system bodies live in banded.cuh as __host__ __device__ functions: tools/banded_emul.cpp runs exactly that code on
// the CPU (tests/test_banded_emul.py).  Measurements and the variants that lost: profiles/r2_banded.txt.
// HBM-bound: per system and sweep pair, n * (2 * sizeof(rhs element) [read b, write x] + sizeof(rhs element) * 2 [y written
// and read back, L2-resident for fields <= ~100 MB] + (p + q + 1) * sizeof(band element)) bytes.
#include <memory>
#include <new>
#include <type_traits>

#include "banded.cuh"
#include "jfx_common.h"

struct jfx_banded {
  int dtype = 0;
  int band_complex = 0;
  int p = 0, q = 0;
  int64_t n = 0, n_sys = 0;
  void* lu = nullptr;   // device, [p + q + 1][n][n_sys] band elements (real or complex of the dtype's precision)
  size_t lu_bytes = 0;
};

namespace jfx {
namespace {

using banded::BA;
using banded::BandElem;

template <typename R, bool EC>
__global__ void __launch_bounds__(256) banded_assemble_kernel(BandElem<R, EC>* __restrict__ lu, const double* __restrict__ W,
                                                              const double* __restrict__ P, const int* __restrict__ rows,
                                                              int n_terms, int n_diags, int64_t n, int64_t n_sys) {
  const int64_t total = (int64_t)n_diags * n * n_sys;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x)
    banded::assemble_entry<R, EC>(lu, W, P, rows, n_terms, n_diags, n, n_sys, idx);
}

template <typename R, bool EC>
__global__ void __launch_bounds__(128) banded_factor_kernel(BandElem<R, EC>* lu, int64_t n, int64_t n_sys, int p, int q,
                                                            int* flag) {
  const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s >= n_sys) return;
  if (banded::factor_system<R, EC>(lu, n, n_sys, p, q, s)) atomicOr(flag, 1);
}

template <typename R, bool EC, bool XC, int W, int U, bool EXACT>
__global__ void __launch_bounds__(128) banded_solve_kernel(const BandElem<R, EC>* __restrict__ lu,
                                                           const typename BA<R, EC, XC>::X* rhs,
                                                           typename BA<R, EC, XC>::X* out, int64_t n, int64_t n_sys,
                                                           int64_t inner, int p, int q) {
  const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s >= n_sys) return;
  banded::solve_system<R, EC, XC, W, U, EXACT>(lu, rhs, out, n, n_sys, inner, p, q, s);
}

// polynomial axis last: one warp per 32 consecutive rows, tiles staged through shared memory (banded.cuh: solve_rows_warp)
template <typename R, bool EC, bool XC, int W>
__global__ void __launch_bounds__(128) banded_solve_rows_kernel(const BandElem<R, EC>* __restrict__ lu,
                                                                const typename BA<R, EC, XC>::X* rhs,
                                                                typename BA<R, EC, XC>::X* out, int64_t n, int64_t n_sys,
                                                                int p, int q) {
  using X = typename BA<R, EC, XC>::X;
  constexpr int TILE = 32 * banded::rows_pitch<R, EC, XC>();
  __shared__ X tiles[4 * TILE];
  const int warp = (int)(threadIdx.x >> 5);
  const int64_t s0 = (blockIdx.x * (int64_t)(blockDim.x >> 5) + warp) * 32;
  if (s0 >= n_sys) return;
  banded::solve_rows_warp<R, EC, XC, W>(lu, rhs, out, n, n_sys, p, q, s0, tiles + warp * TILE);
}

template <typename R, bool EC, bool XC>
int launch_solve_t(cudaStream_t st, const jfx_banded* b, const void* rhs, void* out, int64_t inner) {
  using A = BA<R, EC, XC>;
  using E = typename A::E;
  using X = typename A::X;
  if (banded::rows_variant_applies(inner, b->p, b->q, b->n_sys)) {
    const int wpc = 4;   // 4 warps of 32 rows per CTA
    const unsigned nb = (unsigned)((b->n_sys + 32 * wpc - 1) / (32 * wpc));
    const E* lu_ = static_cast<const E*>(b->lu);
    if (b->p <= 2 && b->q <= 2)
      banded_solve_rows_kernel<R, EC, XC, 2><<<nb, 32 * wpc, 0, st>>>(lu_, static_cast<const X*>(rhs), static_cast<X*>(out), b->n,
                                                                    b->n_sys, b->p, b->q);
    else
      banded_solve_rows_kernel<R, EC, XC, 4><<<nb, 32 * wpc, 0, st>>>(lu_, static_cast<const X*>(rhs), static_cast<X*>(out), b->n,
                                                                    b->n_sys, b->p, b->q);
    JFX_CUDA_OK(cudaGetLastError());
    return JFX_OK;
  }
  // few systems: small CTAs so that they spread over the SMs; many: 128 threads
  const int threads = b->n_sys >= 148 * 128 ? 128 : (b->n_sys >= 148 * 64 ? 64 : 32);
  const unsigned blocks = (unsigned)((b->n_sys + threads - 1) / threads);
  const E* lu = static_cast<const E*>(b->lu);
  const X* r = static_cast<const X*>(rhs);
  X* o = static_cast<X*>(out);
  banded::dispatch_window(b->p, b->q, [&](auto w, auto u, auto exact) {
    banded_solve_kernel<R, EC, XC, decltype(w)::value, decltype(u)::value, decltype(exact)::value>
        <<<blocks, threads, 0, st>>>(lu, r, o, b->n, b->n_sys, inner, b->p, b->q);
  });
  JFX_CUDA_OK(cudaGetLastError());
  return JFX_OK;
}

template <typename R, bool EC>
int build_t(jfx_banded* b, const double* dW, const double* dP, const int* drows, int n_terms, int n_diags, int* dflag) {
  using E = BandElem<R, EC>;
  E* lu = static_cast<E*>(b->lu);
  const int64_t total = (int64_t)n_diags * b->n * b->n_sys;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  banded_assemble_kernel<R, EC><<<(unsigned)blocks, 256>>>(lu, dW, dP, drows, n_terms, n_diags, b->n, b->n_sys);
  JFX_CUDA_OK(cudaGetLastError());
  const int threads = b->n_sys >= 148 * 128 ? 128 : 32;
  banded_factor_kernel<R, EC><<<(unsigned)((b->n_sys + threads - 1) / threads), threads>>>(lu, b->n, b->n_sys, b->p, b->q, dflag);
  JFX_CUDA_OK(cudaGetLastError());
  return JFX_OK;
}

struct DevBuf {   // scoped device allocation for the temporaries of jfx_banded_create
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
};

}  // namespace
}  // namespace jfx

extern "C" {

int jfx_banded_create(const jfx_banded_desc* d, jfx_banded** out) {
  using namespace jfx;
  JFX_REQUIRE(d && out, JFX_ERR_INVALID, "null argument");
  *out = nullptr;
  JFX_REQUIRE(d->abi_version == JFX_ABI_VERSION, JFX_ERR_INVALID, "ABI version mismatch (%d, library %d)", d->abi_version,
              JFX_ABI_VERSION);
  JFX_REQUIRE(d->dtype >= JFX_F32 && d->dtype <= JFX_C128, JFX_ERR_INVALID, "unknown dtype %d", d->dtype);
  JFX_REQUIRE(d->n >= 1 && d->n_sys >= 1, JFX_ERR_INVALID, "system length n = %lld and count n_sys = %lld must be positive",
              (long long)d->n, (long long)d->n_sys);
  JFX_REQUIRE(d->n_terms >= 1 && d->n_terms <= JFX_BANDED_MAX_TERMS, JFX_ERR_INVALID, "n_terms must be 1..%d",
              JFX_BANDED_MAX_TERMS);
  JFX_REQUIRE(d->n_diags >= 1 && d->offsets && d->weights && d->diags, JFX_ERR_INVALID,
              "offsets, weights and diags are required (n_diags >= 1)");
  JFX_REQUIRE(!(d->band_complex && !dtype_is_complex(d->dtype)), JFX_ERR_INVALID,
              "complex matrices need a complex right-hand-side dtype");
  int p = 0, q = 0;
  bool has_main = false;
  for (int k = 0; k < d->n_diags; ++k) {
    const int off = d->offsets[k];
    JFX_REQUIRE(k == 0 || off > d->offsets[k - 1], JFX_ERR_INVALID, "offsets must be strictly increasing");
    JFX_REQUIRE((off < 0 ? -off : off) < d->n, JFX_ERR_INVALID, "offset %d lies outside an n = %lld matrix", off,
                (long long)d->n);
    if (off == 0) has_main = true;
    if (-off > p) p = -off;
    if (off > q) q = off;
  }
  JFX_REQUIRE(has_main, JFX_ERR_INVALID, "the main diagonal (offset 0) is required: LU without pivoting");
  const int bw = p + q + 1;
  const size_t esize = (dtype_is_double(d->dtype) ? 8 : 4) * (d->band_complex ? 2 : 1);
  const double elems = (double)bw * (double)d->n * (double)d->n_sys;
  JFX_REQUIRE(elems * esize < 1.6e11, JFX_ERR_NOMEM, "factor storage of %.3g bytes does not fit one device", elems * esize);
  JFX_REQUIRE(jfx_device_count() > 0, JFX_ERR_CUDA, "no CUDA device: the jfx engine has no CPU fallback");

  std::unique_ptr<jfx_banded> b(new (std::nothrow) jfx_banded);
  JFX_REQUIRE(b, JFX_ERR_NOMEM, "out of host memory");
  b->dtype = d->dtype;
  b->band_complex = d->band_complex ? 1 : 0;
  b->p = p;
  b->q = q;
  b->n = d->n;
  b->n_sys = d->n_sys;
  b->lu_bytes = (size_t)bw * (size_t)d->n * (size_t)d->n_sys * esize;

  const size_t csize = d->band_complex ? 16 : 8;
  const size_t w_bytes = (size_t)d->n_terms * (size_t)d->n_sys * csize;
  const size_t p_bytes = (size_t)d->n_terms * (size_t)d->n_diags * (size_t)d->n * csize;
  std::vector<int> rows(d->n_diags);
  for (int k = 0; k < d->n_diags; ++k) rows[k] = p + d->offsets[k];
  DevBuf dW, dP, dR, dF, lu;
  JFX_CUDA_OK(cudaMalloc(&lu.p, b->lu_bytes));
  JFX_CUDA_OK(cudaMalloc(&dW.p, w_bytes));
  JFX_CUDA_OK(cudaMalloc(&dP.p, p_bytes));
  JFX_CUDA_OK(cudaMalloc(&dR.p, rows.size() * sizeof(int)));
  JFX_CUDA_OK(cudaMalloc(&dF.p, sizeof(int)));
  JFX_CUDA_OK(cudaMemcpy(dW.p, d->weights, w_bytes, cudaMemcpyHostToDevice));
  JFX_CUDA_OK(cudaMemcpy(dP.p, d->diags, p_bytes, cudaMemcpyHostToDevice));
  JFX_CUDA_OK(cudaMemcpy(dR.p, rows.data(), rows.size() * sizeof(int), cudaMemcpyHostToDevice));
  JFX_CUDA_OK(cudaMemset(dF.p, 0, sizeof(int)));
  JFX_CUDA_OK(cudaMemset(lu.p, 0, b->lu_bytes));
  b->lu = lu.p;
  int rc;
  const double* w = static_cast<const double*>(dW.p);
  const double* pp = static_cast<const double*>(dP.p);
  const int* rr = static_cast<const int*>(dR.p);
  int* ff = static_cast<int*>(dF.p);
  if (dtype_is_double(d->dtype))
    rc = d->band_complex ? build_t<double, true>(b.get(), w, pp, rr, d->n_terms, d->n_diags, ff)
                         : build_t<double, false>(b.get(), w, pp, rr, d->n_terms, d->n_diags, ff);
  else
    rc = d->band_complex ? build_t<float, true>(b.get(), w, pp, rr, d->n_terms, d->n_diags, ff)
                         : build_t<float, false>(b.get(), w, pp, rr, d->n_terms, d->n_diags, ff);
  if (rc != JFX_OK) { b->lu = nullptr; return rc; }
  int flag = 0;
  // creation may synchronise (like jfx_plan_create); jfx_banded_solve never does
  cudaError_t e = cudaMemcpy(&flag, dF.p, sizeof(int), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    b->lu = nullptr;
    set_error("banded factorisation failed: %s", cudaGetErrorString(e));
    return JFX_ERR_CUDA;
  }
  if (flag) {
    b->lu = nullptr;
    set_error("a banded system is singular or has a zero / non-finite pivot in LU without pivoting");
    return JFX_ERR_UNSUPPORTED;
  }
  lu.p = nullptr;   // ownership moves to the object
  *out = b.release();
  return JFX_OK;
}

void jfx_banded_destroy(jfx_banded* b) {
  if (!b) return;
  if (b->lu) cudaFree(b->lu);
  delete b;
}

int jfx_banded_info(const jfx_banded* b, int32_t* p, int32_t* q, size_t* factor_bytes) {
  using namespace jfx;
  JFX_REQUIRE(b, JFX_ERR_INVALID, "null argument");
  if (p) *p = b->p;
  if (q) *q = b->q;
  if (factor_bytes) *factor_bytes = b->lu_bytes;
  return JFX_OK;
}

int jfx_banded_factors(const jfx_banded* b, void* lu_host) {
  using namespace jfx;
  JFX_REQUIRE(b && lu_host, JFX_ERR_INVALID, "null argument");
  JFX_CUDA_OK(cudaMemcpy(lu_host, b->lu, b->lu_bytes, cudaMemcpyDeviceToHost));
  return JFX_OK;
}

int jfx_banded_solve(const jfx_banded* b, void* stream, const void* rhs, void* out, int64_t outer, int64_t inner) {
  using namespace jfx;
  JFX_REQUIRE(b && rhs && out, JFX_ERR_INVALID, "null argument");
  JFX_REQUIRE(outer >= 1 && inner >= 1 && outer * inner == b->n_sys, JFX_ERR_INVALID,
              "outer * inner = %lld * %lld does not match the %lld factored systems", (long long)outer, (long long)inner,
              (long long)b->n_sys);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (b->dtype) {
    case JFX_F32: return launch_solve_t<float, false, false>(st, b, rhs, out, inner);
    case JFX_F64: return launch_solve_t<double, false, false>(st, b, rhs, out, inner);
    case JFX_C64:
      return b->band_complex ? launch_solve_t<float, true, true>(st, b, rhs, out, inner)
                             : launch_solve_t<float, false, true>(st, b, rhs, out, inner);
    default:
      return b->band_complex ? launch_solve_t<double, true, true>(st, b, rhs, out, inner)
                             : launch_solve_t<double, false, true>(st, b, rhs, out, inner);
  }
}

}  // extern "C"
