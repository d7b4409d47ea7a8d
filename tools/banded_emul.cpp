// Host execution of the device code of the wavenumber-batched banded solver: the per-system bodies in
// jaxfun_b200/csrc/banded.cuh are __host__ __device__, so this file runs EXACTLY the arithmetic and indexing of the CUDA
// kernels (assembly, LU without pivoting, both substitution sweeps, register-window and generic variants) with a loop over
// the systems in place of the thread grid.  Test infrastructure (tests/test_banded_emul.py builds it with g++ and compares
// with the oracle on hosts without a GPU); never loaded by the product.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../jaxfun_b200/csrc/banded.cuh"

using namespace jfx::banded;

template <typename R, bool EC, bool XC>
static int run(int n_terms, int64_t n, int64_t n_sys, int n_diags, const int* offsets, const double* W, const double* P,
               const void* rhs, void* out, int64_t inner, void* lu_out, int use_rows) {
  using E = BandElem<R, EC>;
  using X = typename BA<R, EC, XC>::X;
  int p = 0, q = 0;
  for (int k = 0; k < n_diags; ++k) {
    if (-offsets[k] > p) p = -offsets[k];
    if (offsets[k] > q) q = offsets[k];
  }
  std::vector<int> rows(n_diags);
  for (int k = 0; k < n_diags; ++k) rows[k] = p + offsets[k];
  const int64_t bw = p + q + 1;
  std::vector<E> lu((size_t)(bw * n * n_sys));
  std::memset(lu.data(), 0, lu.size() * sizeof(E));
  for (int64_t idx = 0; idx < (int64_t)n_diags * n * n_sys; ++idx)
    assemble_entry<R, EC>(lu.data(), W, P, rows.data(), n_terms, n_diags, n, n_sys, idx);
  int flag = 0;
  for (int64_t s = 0; s < n_sys; ++s)
    if (factor_system<R, EC>(lu.data(), n, n_sys, p, q, s)) flag = 1;
  if (lu_out) std::memcpy(lu_out, lu.data(), lu.size() * sizeof(E));
  if (flag) return 1;
  if (inner == 1 && (p > q ? p : q) <= 4 && use_rows == 1) {      // the launch logic of kernels_banded.cu: warps of 32 consecutive rows
    std::vector<X> tile((size_t)32 * rows_pitch<R, EC, XC>());
    for (int64_t s0 = 0; s0 < n_sys; s0 += 32) {
      if (p <= 2 && q <= 2) solve_rows_warp<R, EC, XC, 2>(lu.data(), static_cast<const X*>(rhs), static_cast<X*>(out), n, n_sys, p, q, s0, tile.data());
      else solve_rows_warp<R, EC, XC, 4>(lu.data(), static_cast<const X*>(rhs), static_cast<X*>(out), n, n_sys, p, q, s0, tile.data());
    }
    return 0;
  }
  dispatch_window(p, q, [&](auto w, auto u, auto exact) {
    for (int64_t s = 0; s < n_sys; ++s)
      solve_system<R, EC, XC, decltype(w)::value, decltype(u)::value, decltype(exact)::value>(lu.data(), static_cast<const X*>(rhs), static_cast<X*>(out),
                                                                      n, n_sys, inner, p, q, s);
  });
  return 0;
}

// dtype: 0 f32, 1 f64, 2 c64, 3 c128 (jfx_dtype).  use_rows = 1: the row-tile variant where it can run (polynomial axis last,
// bandwidth <= 4; the device additionally asks for enough systems); else one thread per system with register chunks.  Returns 1 for a zero / non-finite pivot, -1 for a bad combination.
extern "C" int banded_emul(int dtype, int band_complex, int n_terms, int64_t n, int64_t n_sys, int n_diags, const int* offsets,
                           const double* W, const double* P, const void* rhs, void* out, int64_t inner, void* lu_out, int use_rows) {
  switch (dtype) {
    case 0: return band_complex ? -1 : run<float, false, false>(n_terms, n, n_sys, n_diags, offsets, W, P, rhs, out, inner, lu_out, use_rows);
    case 1: return band_complex ? -1 : run<double, false, false>(n_terms, n, n_sys, n_diags, offsets, W, P, rhs, out, inner, lu_out, use_rows);
    case 2:
      return band_complex ? run<float, true, true>(n_terms, n, n_sys, n_diags, offsets, W, P, rhs, out, inner, lu_out, use_rows)
                          : run<float, false, true>(n_terms, n, n_sys, n_diags, offsets, W, P, rhs, out, inner, lu_out, use_rows);
    case 3:
      return band_complex ? run<double, true, true>(n_terms, n, n_sys, n_diags, offsets, W, P, rhs, out, inner, lu_out, use_rows)
                          : run<double, false, true>(n_terms, n, n_sys, n_diags, offsets, W, P, rhs, out, inner, lu_out, use_rows);
  }
  return -1;
}
