// Per-system bodies of the wavenumber-batched banded solver (see kernels_banded.cu for the design).  They are
// __host__ __device__ so that tools/banded_emul.cpp can run exactly this code on the CPU (tests/test_banded_emul.py): the
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
  static JFX_HD E recip(const E& d) {
    if constexpr (!EC) return R(1) / d;
    else return cx_div(Cx<R>{R(1), R(0)}, d);
  }
  // a * r
  static JFX_HD X mul(const X& a, const E& r) {
    if constexpr (!XC) return a * r;
    else if constexpr (!EC) return X{a.re * r, a.im * r};
    else return X{a.re * r.re - a.im * r.im, a.re * r.im + a.im * r.re};
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
template <typename R, bool EC, bool XC, int W, int U, bool EXACT = false>
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
    // The loads of a chunk (U steps: right-hand side and matrix entries) do not depend on the recurrence: they are all issued
    // before the chunk's dependent arithmetic.  The reciprocal of the U diagonal is formed in the load phase, after ALL loads
    // of the chunk (the sweep multiplies by it: one rounding more than the reference's division, far inside the 1e-12 bar).
    // With few systems there is one warp per scheduler and the sweep is bound by the INSTRUCTIONS per step (measured: 45 / 83
    // per forward / backward step in the first build = 143 ns per step), so the interior chunks — every index in range — run a
    // lean variant (FULL): no per-step predicates, addresses by pointer increments; EXACT (p == q == W) drops the band guards.
    X win[W];
#pragma unroll
    for (int t = 0; t < W; ++t) win[t] = A::xzero();
    // ---- forward elimination: y_j = b_j - sum_{t=1..p} L[j, j-t] y_{j-t},   L[j, j-t] = band[p - t][j - t]
    auto fwd_chunk = [&](int64_t j0, auto full_tag) {
      constexpr bool FULL = decltype(full_tag)::value;
      X bv[U];
      E lv[U][W];
      if constexpr (FULL) {
        const X* bp = b + j0 * inner;
        const E* lp[W];
#pragma unroll
        for (int t = 1; t <= W; ++t) lp[t - 1] = Ls + ((int64_t)(p - t) * n + (j0 - t)) * n_sys;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          bv[u] = bp[u * inner];
#pragma unroll
          for (int t = 1; t <= W; ++t) lv[u][t - 1] = (EXACT || t <= p) ? lp[t - 1][u * n_sys] : A::ezero();
        }
      } else {
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
      }
      X* xp = x + j0 * inner;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (FULL || j0 + u < n) {
          X y = bv[u];
#pragma unroll
          for (int t = 1; t <= W; ++t) A::msub(y, lv[u][t - 1], win[t - 1]);
#pragma unroll
          for (int t = W - 1; t > 0; --t) win[t] = win[t - 1];
          win[0] = y;
          xp[u * inner] = y;
        }
      }
    };
    {
      int64_t j0 = 0;
      for (; j0 < W && j0 < n; j0 += U) fwd_chunk(j0, std::false_type{});      // head: rows whose band leaves the matrix
      for (; j0 + U <= n; j0 += U) fwd_chunk(j0, std::true_type{});            // interior
      for (; j0 < n; j0 += U) fwd_chunk(j0, std::false_type{});                // ragged tail
    }
    // ---- back substitution: x_j = (y_j - sum_{t=1..q} U[j, j+t] x_{j+t}) / U[j, j],   U[j, j+t] = band[p + t][j + t];
    // a chunk covers j = j1 - 1 down to j1 - U
#pragma unroll
    for (int t = 0; t < W; ++t) win[t] = A::xzero();
    auto bwd_chunk = [&](int64_t j1, auto full_tag) {
      constexpr bool FULL = decltype(full_tag)::value;
      X yv[U];
      E uv[U][W], rd[U];
      if constexpr (FULL) {
        const X* yp = x + (j1 - 1) * inner;
        const E* dp = Ls + ((int64_t)p * n + (j1 - 1)) * n_sys;
        const E* up[W];
#pragma unroll
        for (int t = 1; t <= W; ++t) up[t - 1] = Ls + ((int64_t)(p + t) * n + (j1 - 1 + t)) * n_sys;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          yv[u] = *(yp - u * inner);
          rd[u] = *(dp - u * n_sys);
#pragma unroll
          for (int t = 1; t <= W; ++t) uv[u][t - 1] = (EXACT || t <= q) ? *(up[t - 1] - u * n_sys) : A::ezero();
        }
#pragma unroll
        for (int u = 0; u < U; ++u) rd[u] = A::recip(rd[u]);
      } else {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t j = j1 - 1 - u;
          yv[u] = A::xzero();
          rd[u] = A::ezero();
#pragma unroll
          for (int t = 1; t <= W; ++t) uv[u][t - 1] = A::ezero();
          if (j >= 0) {
            yv[u] = x[j * inner];
            rd[u] = ld(p, j);
#pragma unroll
            for (int t = 1; t <= W; ++t)
              if (t <= q && j + t < n) uv[u][t - 1] = ld(p + t, j + t);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (j1 - 1 - u >= 0) rd[u] = A::recip(rd[u]);
      }
      X* xp = x + (j1 - 1) * inner;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (FULL || j1 - 1 - u >= 0) {
          X v = yv[u];
#pragma unroll
          for (int t = 1; t <= W; ++t) A::msub(v, uv[u][t - 1], win[t - 1]);
          v = A::mul(v, rd[u]);
#pragma unroll
          for (int t = W - 1; t > 0; --t) win[t] = win[t - 1];
          win[0] = v;
          *(xp - u * inner) = v;
        }
      }
    };
    {
      int64_t j1 = n;
      for (; j1 > 0 && j1 + W > n; j1 -= U) bwd_chunk(j1, std::false_type{});  // head: rows whose band leaves the matrix
      for (; j1 - U >= 0; j1 -= U) bwd_chunk(j1, std::true_type{});            // interior
      for (; j1 > 0; j1 -= U) bwd_chunk(j1, std::false_type{});                // ragged tail
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

// ---- polynomial axis LAST (inner == 1): one WARP owns 32 consecutive systems = 32 consecutive rows of the array --------
// With one thread per row every warp load touches 32 different 128-byte lines (measured: 0.35 of HBM against 0.67 for the
// coalesced layouts, the load / store unit is the limit).  Here the warp moves a [32 rows x C steps] tile between global and
// shared memory with row-contiguous accesses (128 bytes per row and chunk for complex128, 64 for the narrower types), each
// lane sweeps its own row inside the tile, and the tile goes back the same way.  The matrix entries are loaded directly (they
// are coalesced across systems already).  The phases are separated by __syncwarp only.
// Host emulation: the lane loop replaces the 32 threads, per-lane state is an array.
#if defined(__CUDA_ARCH__)
#define JFX_BANDED_FOR_LANES(lane) for (int lane = (int)(threadIdx.x & 31u), once_ = 1; once_; once_ = 0)
#define JFX_BANDED_LI(lane) 0
#define JFX_BANDED_NSTATE 1
#define JFX_BANDED_SYNC() __syncwarp()
#else
#define JFX_BANDED_FOR_LANES(lane) for (int lane = 0; lane < 32; ++lane)
#define JFX_BANDED_LI(lane) (lane)
#define JFX_BANDED_NSTATE 32
#define JFX_BANDED_SYNC() ((void)0)
#endif

template <typename R, bool EC, bool XC> constexpr int rows_chunk() {
  return sizeof(typename BA<R, EC, XC>::X) >= 8 ? 8 : 16;   // 128-byte row segments for complex128, 64 bytes otherwise (registers)
}
template <typename R, bool EC, bool XC> constexpr int rows_pitch() { return rows_chunk<R, EC, XC>() + 1; }   // conflict-free

template <typename R, bool EC, bool XC, int W>
JFX_HD void solve_rows_warp(const BandElem<R, EC>* lu, const typename BA<R, EC, XC>::X* rhs, typename BA<R, EC, XC>::X* out,
                            int64_t n, int64_t n_sys, int p, int q, int64_t s0, typename BA<R, EC, XC>::X* tile) {
  using A = BA<R, EC, XC>;
  using E = typename A::E;
  using X = typename A::X;
  constexpr int C = rows_chunk<R, EC, XC>();
  constexpr int LD = rows_pitch<R, EC, XC>();
  const int64_t nchunks = (n + C - 1) / C;
  X win[JFX_BANDED_NSTATE][W];
  JFX_BANDED_FOR_LANES(lane) {
#pragma unroll
    for (int t = 0; t < W; ++t) win[JFX_BANDED_LI(lane)][t] = A::xzero();
  }
  auto tile_in = [&](const X* src, int64_t j0) {
    JFX_BANDED_FOR_LANES(lane) {
#pragma unroll
      for (int it = 0; it < C; ++it) {
        const int e = it * 32 + lane, row = e / C, col = e % C;
        const int64_t s = s0 + row, j = j0 + col;
        if (s < n_sys && j < n) tile[row * LD + col] = src[s * n + j];
      }
    }
    JFX_BANDED_SYNC();
  };
  auto tile_out = [&](int64_t j0) {
    JFX_BANDED_SYNC();
    JFX_BANDED_FOR_LANES(lane) {
#pragma unroll
      for (int it = 0; it < C; ++it) {
        const int e = it * 32 + lane, row = e / C, col = e % C;
        const int64_t s = s0 + row, j = j0 + col;
        if (s < n_sys && j < n) out[s * n + j] = tile[row * LD + col];
      }
    }
    JFX_BANDED_SYNC();
  };
  // forward elimination
  for (int64_t c = 0; c < nchunks; ++c) {
    const int64_t j0 = c * C;
    tile_in(rhs, j0);
    JFX_BANDED_FOR_LANES(lane) {
      const int64_t s = s0 + lane;
      if (s < n_sys) {
        const E* Ls = lu + s;
        X* w = win[JFX_BANDED_LI(lane)];
        E lv[C][W];
#pragma unroll
        for (int u = 0; u < C; ++u) {
          const int64_t j = j0 + u;
#pragma unroll
          for (int t = 1; t <= W; ++t) {
            lv[u][t - 1] = A::ezero();
            if (j < n && t <= p && j - t >= 0) lv[u][t - 1] = Ls[((int64_t)(p - t) * n + (j - t)) * n_sys];
          }
        }
#pragma unroll
        for (int u = 0; u < C; ++u) {
          if (j0 + u < n) {
            X y = tile[lane * LD + u];
#pragma unroll
            for (int t = 1; t <= W; ++t) A::msub(y, lv[u][t - 1], w[t - 1]);
#pragma unroll
            for (int t = W - 1; t > 0; --t) w[t] = w[t - 1];
            w[0] = y;
            tile[lane * LD + u] = y;
          }
        }
      }
    }
    tile_out(j0);
  }
  // back substitution, chunks from the top, steps inside a chunk from the top
  JFX_BANDED_FOR_LANES(lane) {
#pragma unroll
    for (int t = 0; t < W; ++t) win[JFX_BANDED_LI(lane)][t] = A::xzero();
  }
  for (int64_t c = nchunks - 1; c >= 0; --c) {
    const int64_t j0 = c * C;
    tile_in(out, j0);
    JFX_BANDED_FOR_LANES(lane) {
      const int64_t s = s0 + lane;
      if (s < n_sys) {
        const E* Ls = lu + s;
        X* w = win[JFX_BANDED_LI(lane)];
        E uv[C][W], rd[C];
#pragma unroll
        for (int u = 0; u < C; ++u) {
          const int64_t j = j0 + u;
          rd[u] = A::ezero();
          if (j < n) rd[u] = Ls[((int64_t)p * n + j) * n_sys];
#pragma unroll
          for (int t = 1; t <= W; ++t) {
            uv[u][t - 1] = A::ezero();
            if (j < n && t <= q && j + t < n) uv[u][t - 1] = Ls[((int64_t)(p + t) * n + (j + t)) * n_sys];
          }
        }
#pragma unroll
        for (int u = 0; u < C; ++u)
          if (j0 + u < n) rd[u] = A::recip(rd[u]);
#pragma unroll
        for (int u = C - 1; u >= 0; --u) {
          if (j0 + u < n) {
            X v = tile[lane * LD + u];
#pragma unroll
            for (int t = 1; t <= W; ++t) A::msub(v, uv[u][t - 1], w[t - 1]);
            v = A::mul(v, rd[u]);
#pragma unroll
            for (int t = W - 1; t > 0; --t) w[t] = w[t - 1];
            w[0] = v;
            tile[lane * LD + u] = v;
          }
        }
      }
    }
    tile_out(j0);
  }
}

// The row-tile variant serves bandwidths up to 4 on arrays whose polynomial axis is last, when there are enough systems to fill
// the GPU with warps (its chunks are 8 steps, so with few warps it is MORE latency-bound than one thread per system: measured
// 470 vs 292 us for 1024 systems, 456 vs 518 us for 65 536).  Build switches: JFX_BANDED_ROWS=0 off, JFX_BANDED_ROWS_MIN.
#ifndef JFX_BANDED_ROWS
#define JFX_BANDED_ROWS 1
#endif
#ifndef JFX_BANDED_ROWS_MIN
#define JFX_BANDED_ROWS_MIN (148 * 128)
#endif
inline bool rows_variant_applies(int64_t inner, int p, int q, int64_t n_sys) {
  return JFX_BANDED_ROWS && inner == 1 && (p > q ? p : q) <= 4 && n_sys >= JFX_BANDED_ROWS_MIN;
}

// register-window width W and chunk length U for a bandwidth; wider than 8: the generic path (0, 1)
#ifndef JFX_BANDED_U2
#define JFX_BANDED_U2 12   /* measured 8 / 12 / 16: 12 is fastest on every shape but 4096 x 4094 (profiles/r2_banded.txt) */
#endif
#ifndef JFX_BANDED_U4
#define JFX_BANDED_U4 4
#endif
#ifndef JFX_BANDED_U8
#define JFX_BANDED_U8 2
#endif
template <typename F> inline void dispatch_window(int p, int q, F&& f) {
  const int w = p > q ? p : q;
  auto pick = [&](auto wt, auto ut) {
    if (p == decltype(wt)::value && q == decltype(wt)::value) f(wt, ut, std::true_type{});   // EXACT: full band, no guards
    else f(wt, ut, std::false_type{});
  };
  if (w <= 2) pick(std::integral_constant<int, 2>{}, std::integral_constant<int, JFX_BANDED_U2>{});
  else if (w <= 4) pick(std::integral_constant<int, 4>{}, std::integral_constant<int, JFX_BANDED_U4>{});
  else if (w <= 8) pick(std::integral_constant<int, 8>{}, std::integral_constant<int, JFX_BANDED_U8>{});
  else f(std::integral_constant<int, 0>{}, std::integral_constant<int, 1>{}, std::false_type{});
}

}  // namespace banded
}  // namespace jfx
