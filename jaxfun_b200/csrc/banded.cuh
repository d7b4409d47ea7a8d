// Per-system bodies of the wavenumber-batched banded solver (see kernels_banded.cu for the design).  They are
// __host__ __device__ so that tools/banded_emul.cu can run exactly this code on the CPU (tests/test_banded_emul.py): the
// kernels in kernels_banded.cu are thin wrappers that map one thread to one system (or one band entry).
#pragma once
#include <cmath>
#include <cstdint>
#include <type_traits>

#ifdef __CUDACC__
#define JFX_HD __host__ __device__ __forceinline__
#else
#define JFX_HD inline
#endif

namespace jfx {
namespace banded {

template <typename T> struct Cx { T re, im; };   // layout of an interleaved complex element

template <typename R, bool EC> using BandElem = std::conditional_t<EC, Cx<R>, R>;

template <typename R> JFX_HD bool finite(R x) { return (x - x) == R(0); }   // false for inf and NaN (no fast-math here)

template <typename R> JFX_HD Cx<R> cx_div(const Cx<R>& a, const Cx<R>& b) {
  const R d = b.re * b.re + b.im * b.im;
  return {(a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d};
}

template <typename R, bool EC, bool XC>
struct BA {
  static_assert(!(EC && !XC), "complex matrices need complex right-hand sides");
  using E = std::conditional_t<EC, Cx<R>, R>;
  using X = std::conditional_t<XC, Cx<R>, R>;
  static JFX_HD X xzero() {
    if constexpr (XC) return X{R(0), R(0)};
    else return R(0);
  }
  static JFX_HD E ezero() {
    if constexpr (EC) return E{R(0), R(0)};
    else return R(0);
  }
  // acc -= l * w
  static JFX_HD void msub(X& acc, const E& l, const X& w) {
    if constexpr (!XC) {
      acc -= l * w;
    } else if constexpr (!EC) {
      acc.re -= l * w.re;
      acc.im -= l * w.im;
    } else {
      acc.re -= l.re * w.re - l.im * w.im;
      acc.im -= l.re * w.im + l.im * w.re;
    }
  }
  static JFX_HD X div(const X& a, const E& d) {
    if constexpr (!XC) return a / d;
    else if constexpr (!EC) return X{a.re / d, a.im / d};
    else return cx_div(a, d);
  }
};

// lu[(row * n + j) * n_sys + s] = sum_t W[t, s] * P[t, d, j]   for row = rows[d] (flat entry index idx = (d * n + j) * n_sys + s);
// the other rows were zeroed.  W, P: float64 (EC = false) or complex128 (EC = true) as the host handed them over; the sum is
// formed in double.
template <typename R, bool EC>
JFX_HD void assemble_entry(BandElem<R, EC>* lu, const double* W, const double* P, const int* rows, int n_terms, int n_diags,
                           int64_t n, int64_t n_sys, int64_t idx) {
  const int64_t s = idx % n_sys;
  const int64_t j = (idx / n_sys) % n;
  const int d = (int)(idx / (n_sys * n));
  double re = 0.0, im = 0.0;
  for (int t = 0; t < n_terms; ++t) {
    const int64_t wi = (int64_t)t * n_sys + s, pi = ((int64_t)t * n_diags + d) * n + j;
    if constexpr (EC) {
      const double wr = W[2 * wi], wim = W[2 * wi + 1], pr = P[2 * pi], pim = P[2 * pi + 1];
      re += wr * pr - wim * pim;
      im += wr * pim + wim * pr;
    } else {
      re += W[wi] * P[pi];
    }
  }
  const int64_t dst = ((int64_t)rows[d] * n + j) * n_sys + s;
  if constexpr (EC) lu[dst] = Cx<R>{(R)re, (R)im};
  else lu[dst] = (R)re;
}

// In-place LU without pivoting of system s (diamatrix.py:1937-1973): band[center + off][j] = A[j - off, j], center = p.
// After it: rows 0 .. p-1 hold the multipliers of L (unit diagonal implied), rows p .. p+q hold U.
// Returns true when a pivot is zero or not finite (the reference's DiaMatrix.lu_factor raises for those, diamatrix.py:461-471).
template <typename R, bool EC>
JFX_HD bool factor_system(BandElem<R, EC>* lu, int64_t n, int64_t n_sys, int p, int q, int64_t s) {
  using E = BandElem<R, EC>;
  E* B = lu + s;
  auto at = [&](int row, int64_t j) -> E& { return B[((int64_t)row * n + j) * n_sys]; };
  bool bad = false;
  for (int64_t k = 0; k < n; ++k) {
    const E piv = at(p, k);
    if constexpr (EC) bad |= !(finite(piv.re) && finite(piv.im)) || (piv.re == R(0) && piv.im == R(0));
    else bad |= !finite(piv) || piv == R(0);
    for (int sdiag = 1; sdiag <= p && k + sdiag < n; ++sdiag) {
      E f;
      if constexpr (EC) f = cx_div(at(p - sdiag, k), piv);
      else f = at(p - sdiag, k) / piv;
      at(p - sdiag, k) = f;
      for (int u = 1; u <= q && k + u < n; ++u) {
        const E up = at(p + u, k + u);
        E& tgt = at(p + u - sdiag, k + u);
        if constexpr (EC) {
          tgt.re -= f.re * up.re - f.im * up.im;
          tgt.im -= f.re * up.im + f.im * up.re;
        } else {
          tgt -= f * up;
        }
      }
    }
  }
  return bad;
}

// Both sweeps of  L y = b,  U x = y  for system s (tpmatrix.py:637-680: `_fwd_elim`, `_bwd_sub`); the array is addressed as
// [outer, n, inner] with s = o * inner + i.  W > 0: p, q <= W, window in registers, the loads of U consecutive steps are
// issued before their dependent arithmetic.  W == 0: any bandwidth, earlier unknowns are read back from `out`.
// `rhs` may equal `out` (every step reads its right-hand-side entries before it writes them).
template <typename R, bool EC, bool XC, int W, int U>
JFX_HD void solve_system(const BandElem<R, EC>* lu, const typename BA<R, EC, XC>::X* rhs, typename BA<R, EC, XC>::X* out,
                         int64_t n, int64_t n_sys, int64_t inner, int p, int q, int64_t s) {
  using A = BA<R, EC, XC>;
  using E = typename A::E;
  using X = typename A::X;
  const int64_t o = s / inner, i = s - o * inner;
  const X* b = rhs + o * n * inner + i;
  X* x = out + o * n * inner + i;
  const E* Ls = lu + s;
  auto ld = [&](int row, int64_t j) -> E { return Ls[((int64_t)row * n + j) * n_sys]; };

  if constexpr (W > 0) {
    X win[W];
#pragma unroll
    for (int t = 0; t < W; ++t) win[t] = A::xzero();
    // forward elimination: y_j = b_j - sum_{t=1..p} L[j, j-t] y_{j-t},   L[j, j-t] = band[p - t][j - t]
    for (int64_t j0 = 0; j0 < n; j0 += U) {
      X bv[U];
      E lv[U][W];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t j = j0 + u;
        bv[u] = A::xzero();
#pragma unroll
        for (int t = 1; t <= W; ++t) lv[u][t - 1] = A::ezero();
        if (j < n) {
          bv[u] = b[j * inner];
#pragma unroll
          for (int t = 1; t <= W; ++t)
            if (t <= p && j - t >= 0) lv[u][t - 1] = ld(p - t, j - t);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t j = j0 + u;
        if (j < n) {
          X y = bv[u];
#pragma unroll
          for (int t = 1; t <= W; ++t) A::msub(y, lv[u][t - 1], win[t - 1]);
#pragma unroll
          for (int t = W - 1; t > 0; --t) win[t] = win[t - 1];
          win[0] = y;
          x[j * inner] = y;
        }
      }
    }
    // back substitution: x_j = (y_j - sum_{t=1..q} U[j, j+t] x_{j+t}) / U[j, j],   U[j, j+t] = band[p + t][j + t]
#pragma unroll
    for (int t = 0; t < W; ++t) win[t] = A::xzero();
    for (int64_t j1 = n; j1 > 0; j1 -= U) {
      X yv[U];
      E uv[U][W], dv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t j = j1 - 1 - u;
        yv[u] = A::xzero();
        dv[u] = A::ezero();
#pragma unroll
        for (int t = 1; t <= W; ++t) uv[u][t - 1] = A::ezero();
        if (j >= 0) {
          yv[u] = x[j * inner];
          dv[u] = ld(p, j);
#pragma unroll
          for (int t = 1; t <= W; ++t)
            if (t <= q && j + t < n) uv[u][t - 1] = ld(p + t, j + t);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t j = j1 - 1 - u;
        if (j >= 0) {
          X v = yv[u];
#pragma unroll
          for (int t = 1; t <= W; ++t) A::msub(v, uv[u][t - 1], win[t - 1]);
          v = A::div(v, dv[u]);
#pragma unroll
          for (int t = W - 1; t > 0; --t) win[t] = win[t - 1];
          win[0] = v;
          x[j * inner] = v;
        }
      }
    }
  } else {
    for (int64_t j = 0; j < n; ++j) {
      X y = b[j * inner];
      const int tmax = (int)(j < p ? j : p);
      for (int t = 1; t <= tmax; ++t) A::msub(y, ld(p - t, j - t), x[(j - t) * inner]);
      x[j * inner] = y;
    }
    for (int64_t j = n - 1; j >= 0; --j) {
      X v = x[j * inner];
      const int tmax = (int)(n - 1 - j < q ? n - 1 - j : q);
      for (int t = 1; t <= tmax; ++t) A::msub(v, ld(p + t, j + t), x[(j + t) * inner]);
      x[j * inner] = A::div(v, ld(p, j));
    }
  }
}

// register-window width and chunk length for a bandwidth: (W, U) = (2, 8), (4, 4), (8, 2), else the generic path (0, 1)
template <typename F> inline void dispatch_window(int p, int q, F&& f) {
  const int w = p > q ? p : q;
  if (w <= 2) f(std::integral_constant<int, 2>{}, std::integral_constant<int, 8>{});
  else if (w <= 4) f(std::integral_constant<int, 4>{}, std::integral_constant<int, 4>{});
  else if (w <= 8) f(std::integral_constant<int, 8>{}, std::integral_constant<int, 2>{});
  else f(std::integral_constant<int, 0>{}, std::integral_constant<int, 1>{});
}

}  // namespace banded
}  // namespace jfx
