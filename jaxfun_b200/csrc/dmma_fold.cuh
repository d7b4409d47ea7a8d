// Parity-folded FP64 tensor-core contraction: index math shared by the CUDA kernel (kernels_dense_fold.cu)
// and by the host emulator that checks it on a machine without a GPU (tests/emu/fold_emu.cpp).
//
// Why: every table the polynomial bases produce on symmetric nodes (Legendre, Jacobi alpha = beta, Chebyshev,
// ChebyshevU, Ultraspherical, and their composite / derivative tables) obeys psi_k(-x) = (-1)^k psi_k(x) and
// x_{n-1-j} = -x_j, i.e.
//        backward  T[n-1-j, k] = sigma_k T[j, k]      ("OUT" fold: the mirror pair is a pair of OUTPUT rows)
//        forward   T[k, n-1-j] = sigma_k T[k, j]      ("IN"  fold: the mirror pair is a pair of INPUT rows)
// with sigma_k = +1 for k = par_plus (mod 2) and -1 otherwise.  (Reference: the Vandermonde of
// galerkin/orthogonal.py:131-141 on the nodes of Legendre.py:125-137 / Jacobi.py:112-124.)  Then
//        OUT:  u_j = P_j + Q_j,  u_{n-1-j} = P_j - Q_j,   P_j = sum_{k plus} T[j,k] c_k,  Q_j = sum_{k minus} T[j,k] c_k
//        IN :  c_k = sum_{j < n/2} T[k,j] (u_j + sigma_k u_{n-1-j})
// which is HALF the multiply-adds of the plain contraction.  The kernel keeps the 128 x 128 CTA tile and the
// 64 x 32 warp tile of dgemm_dmma_tma, but every warp tile holds a "plus" half and a "minus" half for the
// SAME output rows / columns, so the butterfly is done in registers:
//   NN order (other axes,  C_o = T * X_o): m-tiles 0..3 = plus rows, 4..7 = minus rows of 32 outputs
//   NT order (last axis,   C   = X * T^T): n-tiles 0..1 = plus cols, 2..3 = minus cols of 16 outputs
// One pipeline stage = three 16 KB tiles:
//   OUT_NN: R0 = folded table [128 r'][16 k'],  R1 = X rows of the plus parity [16][128],  R2 = minus parity
//           (the parity split of the coefficient index is done by the TMA unit: 4-D tensor map (n, parity, k/2, batch))
//   IN_NN : R0 = folded table,                  R1 = X rows j0..j0+15,   R2 = X rows n-16-j0..n-1-j0 (mirror, read reversed)
//   OUT_NT: R0 = X[128 rows][k0..k0+15],        R1 = X[128 rows][k0+16..k0+31],  R2 = folded table [128 c'][16 k']
//           (parity split in the fragment load: one LDS.128 fetches (even k, odd k) for the plus and the minus MMA)
//   IN_NT : R0 = X[128 rows][j0..j0+15],        R1 = mirror columns,    R2 = folded table
//
// Asymmetry correction (OUT fold).  The reference's Gauss-Legendre nodes are mirror images only to an ulp
// (utils/fastgl.py:548-556 takes cos(theta) and cos(pi - theta)), so its Vandermonde is T = A + E with A the
// symmetric part that is folded and E[j,k] = (T[j,k] - sigma_k T[n-1-j,k]) / 2 ~ P_k'(x_j) * 1e-16, which for the
// highest modes reaches 3e-12 of the mode's own magnitude at n = 1024 — above the 1e-12 parity bar for an input that
// excites only such a mode.  With p = sum_{k plus} E c, q = sum_{k minus} E c the exact result is
//        u_j = (P + q) + (Q + p),   u_{n-1-j} = (P + q) - (Q + p)
// i.e. the SAME butterfly on accumulators that additionally received E against the coefficients of the OPPOSITE parity.
// The modes whose asymmetry exceeds tol_fix (a tail k' >= kcorr0 of the folded index) therefore get extra k-tiles: the
// table tile comes from correction columns appended to the folded table, the data tiles are fetched with the parity
// swapped (NN: by the TMA coordinates; NT: by the fragment select).  MMA warps and epilogue do not change.
#pragma once
#include <cstdint>
#include <cmath>
#include <vector>

#if defined(__CUDACC__)
#define JFX_HD __host__ __device__ __forceinline__
#else
#define JFX_HD inline
#endif

namespace jfx {
namespace dmma {
namespace fold {

// CPLX_NT is not a fold: complex interleaved data on a LAST table axis, any real table.  It reuses the OUT_NT main loop —
// the LDS.128 that fetches (even, odd) there fetches (re, im) of one complex coefficient here, the "plus" accumulators
// collect the real part and the "minus" accumulators the imaginary part against the SAME table rows — and the IN_NT
// epilogue, which interleaves the two accumulator groups into consecutive columns.
enum Variant { OUT_NN = 0, IN_NN = 1, OUT_NT = 2, IN_NT = 3, CPLX_NT = 4 };
enum FoldType { FOLD_NONE = 0, FOLD_OUT = 1, FOLD_IN = 2 };

constexpr int BK = 16, BM = 128, BN = 128, WM = 64, WN = 32;
constexpr int MMA_WARPS = 8, WARPS_N = BN / WN;
constexpr int STAGES = 4;
constexpr int TILE = 128 * 16;                       // doubles per 16 KB tile
constexpr unsigned STAGE_BYTES = 3u * TILE * 8u;     // 48 KB
constexpr int HALF_PER_TILE = 64;                    // mirror pairs covered by one CTA tile

struct alignas(16) Double2 { double x, y; };

// K-contiguous tile [128 rows][16 k], 128-byte swizzle (TMA SWIZZLE_128B on 128-byte rows): the 16-byte chunk
// (k >> 1) of row r is stored at chunk (k >> 1) ^ (r & 7)
JFX_HD int kc_off(int row, int k) { return row * 16 + ((((k >> 1) ^ (row & 7)) << 1) | (k & 1)); }
// N-contiguous tile [16 k][128 n] stored as 16 sub-tiles [16 k][8 n] of 64-byte rows, 64-byte swizzle: chunk
// ((n & 7) >> 1) of row k is stored at chunk ((n & 7) >> 1) ^ ((k >> 1) & 3)
JFX_HD int nc_off(int k, int n) {
  return (n >> 3) * 128 + k * 8 + (((((n & 7) >> 1) ^ ((k >> 1) & 3)) << 1) | (n & 1));
}
// row permutation of the OUT_NT variant: MMA row g of an m-tile is fed from tile row rho(g), so that the eight lanes
// of one LDS.128 phase (two rows x four chunks) cover the eight 16-byte bank groups exactly once
JFX_HD int rho(int g) { return ((g & 1) << 2) | (g >> 1); }

struct Args {
  // logical problem
  int variant;
  int par_plus;        // parity of the coefficient index whose sigma is +1
  int n_fold;          // extent of the mirrored axis (nodes), even
  int n_other;         // extent of the parity-split axis (modes)
  int half;            // n_fold / 2
  int kfold;           // reduction length of the folded problem: OUT: n_other / 2, IN: n_fold / 2
  // GEMM geometry
  int M, N;            // NN: M = folded table rows (padded), N = inner; NT: M = data rows, N = folded table rows (padded)
  int tiles_m, tiles_n, batch;
  double* C;
  long long ldc, strideC;
  int vec_ok;          // 16-byte aligned vector stores allowed
  // k-tiles: kts_main tiles of the folded problem, then kts_corr asymmetry-correction tiles covering k' >= kcorr0
  int kts_main, kts_corr, kcorr0;
  int tm_rot;          // the kernel visits tile_m in the order (tm + tm_rot) % tiles_m (scatter launches: see launch_dmma_fold_scatter)
  int xwide;           // NN: the X tensor map splits the contiguous axis as (8, inner / 8): one box per X tile (inner % 8 == 0)
};

// parity argument of ktile() for k-tile kt: correction tiles pair every accumulator group with the opposite parity
JFX_HD int ktile_par(const Args& q, int kt) { return kt >= q.kts_main ? 1 - q.par_plus : q.par_plus; }

// ------------------------------------------------------------------------------------------------
// Fragment addressing.  The swizzled offsets kc_off / nc_off of every fragment a lane loads in a k-tile differ from a
// handful of per-lane constants only by COMPILE-TIME terms (m-tile / n-tile index, k-step), because the XOR of the
// swizzle acts on bit fields that the k-step and the lane index occupy separately.  frag_init() computes those
// constants once per kernel (offsets in doubles from the stage base); ktile() then addresses every fragment as
// constant + immediate, so the k loop carries no address arithmetic and few live registers (the compiler cannot see the
// bit-disjointness by itself: it kept ~60 precomputed addresses live and recomputed the rest with LOP3 / LEA per load).
//   lane (g, t), th = t >> 1, tl = t & 1, k-step kk = 0..3 (k = 4 kk + t inside the 16-wide k-tile)
//   K-contiguous tile, row r = R + 8 m + g:   kc_off(r, 4 kk + t) = 16 (R + g) + 128 m + 4 (kk ^ (g >> 1)) + 2 (th ^ (g & 1)) + tl
//   N-contiguous tile, col n = 32 wn + 8 j + g: nc_off(4 kk + t, n) = 512 wn + 128 j + 32 kk + 8 t + 2 ((g >> 1) ^ (2 (kk & 1) + th)) + (g & 1)
//   mirrored reads use k' = 15 - k = 4 (3 - kk) + (3 - t)
// ------------------------------------------------------------------------------------------------
struct Frag {
  int p[4];   // NN: table fragment (A) per kk                 NT: table fragment (B) per kk
  int q[4];   // NN: X fragment per (kk & 1) [+2: mirrored]    OUT_NT: X double2 per (kk & 1);  IN_NT: X per kk
  int r[4];   //                                               IN_NT: mirrored X per kk
};

template <int V>
JFX_HD Frag frag_init(int wm, int wn, int g, int t) {
  Frag f{};
  const int th = t >> 1, tl = t & 1, g0 = g & 1, g2 = g >> 1;
  if constexpr (V == OUT_NN || V == IN_NN) {
    const int a_lane = (wm * WM + g) * 16 + ((th ^ g0) << 1) + tl;
    for (int kk = 0; kk < 4; ++kk) f.p[kk] = a_lane + ((kk ^ g2) << 2);
    for (int par = 0; par < 2; ++par) {
      f.q[par] = wn * 512 + t * 8 + (((g2 ^ ((par << 1) | th)) << 1) | g0);
      f.q[2 + par] = wn * 512 + (3 - t) * 8 + (((g2 ^ (((1 - par) << 1) | (1 - th))) << 1) | g0);   // IN_NN mirror tile
    }
  } else {
    const int b_lane = (wn * WN + g) * 16 + ((th ^ g0) << 1) + tl;
    for (int kk = 0; kk < 4; ++kk) f.p[kk] = b_lane + ((kk ^ g2) << 2);
    if constexpr (V == OUT_NT || V == CPLX_NT) {
      const int rr = rho(g);
      for (int par = 0; par < 2; ++par) f.q[par] = (wm * WM + rr) * 16 + ((((par ^ g0) << 2) | (t ^ g2)) << 1);
    } else {
      const int u_lane = (wm * WM + g) * 16;
      for (int kk = 0; kk < 4; ++kk) {
        f.q[kk] = u_lane + ((kk ^ g2) << 2) + ((th ^ g0) << 1) + tl;
        f.r[kk] = u_lane + (((3 - kk) ^ g2) << 2) + (((1 - th) ^ g0) << 1) + (1 - tl);
      }
    }
  }
  return f;
}

// ------------------------------------------------------------------------------------------------
// one k-tile of the MMA warps (lane (g, t) of warp (wm, wn)); S = base of the stage (3 tiles), fr = frag_init()
// mma(d0, d1, a, b) is the m8n8k4 step (device: mma.sync; emulator: a recorder)
// ------------------------------------------------------------------------------------------------
template <int V, class MMA>
JFX_HD void ktile(const double* S, const Frag& fr, int par_plus, double (&acc)[8][4][2], MMA&& mma) {
  if constexpr (V == CPLX_NT) {
    ktile<OUT_NT>(S, fr, 0, acc, mma);
    return;
  }
  const double* R0 = S;
  const double* R1 = S + TILE;
  const double* R2 = S + 2 * TILE;
#pragma unroll
  for (int kk = 0; kk < BK / 4; ++kk) {
    if constexpr (V == OUT_NN || V == IN_NN) {
      // the X fragments of both halves first (IN: u +- u' formed ONCE per k-step), then the plus half (m-tiles 0..3)
      // and the minus half (m-tiles 4..7) of the table
      double bp[4], bm[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if constexpr (V == OUT_NN) {
          bp[j] = R1[fr.q[kk & 1] + j * 128 + kk * 32];
          bm[j] = R2[fr.q[kk & 1] + j * 128 + kk * 32];
        } else {
          const double u = R1[fr.q[kk & 1] + j * 128 + kk * 32], v = R2[fr.q[2 + (kk & 1)] + j * 128 + (3 - kk) * 32];
          bp[j] = u + v;
          bm[j] = u - v;
        }
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        double a[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = R0[fr.p[kk] + (h * 4 + i) * 128];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) mma(acc[h * 4 + i][j][0], acc[h * 4 + i][j][1], a[i], h ? bm[j] : bp[j]);
      }
    } else {
      double b[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = R2[fr.p[kk] + j * 128];
      // four m-tiles at a time: plus / minus operand pairs of 4 rows live together
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        double ap[4], aq[4];
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
          const int i = h * 4 + ii;
          if constexpr (V == OUT_NT) {
            const double* sub = (kk >> 1) ? R1 : R0;
            const Double2 v = *reinterpret_cast<const Double2*>(sub + fr.q[kk & 1] + i * 128);
            ap[ii] = par_plus ? v.y : v.x;
            aq[ii] = par_plus ? v.x : v.y;
          } else {
            const double u = R0[fr.q[kk] + i * 128], v = R1[fr.r[kk] + i * 128];
            ap[ii] = u + v;
            aq[ii] = u - v;
          }
        }
#pragma unroll
        for (int ii = 0; ii < 4; ++ii)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            mma(acc[h * 4 + ii][j][0], acc[h * 4 + ii][j][1], j < 2 ? ap[ii] : aq[ii], b[j]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// epilogue of one CTA tile (tile_m, tile_n, z).  st.s2(idx, v0, v1): two consecutive elements at an even, 16-byte
// aligned element index of C; st.s1(idx, v): one element.
// ------------------------------------------------------------------------------------------------
template <class ST>
JFX_HD void put2(ST& st, bool vec, long long idx, double v0, double v1, bool ok0, bool ok1) {
  if (vec && ok0 && ok1) {
    st.s2(idx, v0, v1);
  } else {
    if (ok0) st.s1(idx, v0);
    if (ok1) st.s1(idx + 1, v1);
  }
}

// Fast path of the epilogue: the warp's 64 x 32 share of the tile lies completely inside the array and vector stores
// are allowed (true for every tile of the power-of-two BASELINE shapes).  Two running element indices and compile-time
// offsets replace the per-store index arithmetic and bounds predicates of the general path below (which cost ~10
// instructions per store and 2 us of idle tensor pipe per tile, ncu r2a).  Returns false when the general path must run;
// the condition depends on the tile and warp only, so it is warp-uniform.
template <int V, class ST>
JFX_HD bool epilogue_interior(const Args& q, int tile_m, int tile_n, long long z, int wm, int wn, int g, int t,
                              const double (&acc)[8][4][2], ST&& st) {
  if (!q.vec_ok) return false;
  if constexpr (V == OUT_NN) {
    const int j0 = tile_m * HALF_PER_TILE + wm * 32, c0 = tile_n * BN + wn * WN;
    if (j0 + 32 > q.half || c0 + WN > q.N) return false;
    const long long step = 8ll * q.ldc;
    long long lo = z * q.strideC + (long long)(j0 + g) * q.ldc + c0 + 2 * t;
    long long hi = z * q.strideC + (long long)(q.n_fold - 1 - j0 - g) * q.ldc + c0 + 2 * t;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double p0 = acc[i][j][0], p1 = acc[i][j][1], q0 = acc[i + 4][j][0], q1 = acc[i + 4][j][1];
        st.s2(lo + j * 8, p0 + q0, p1 + q1);
        st.s2(hi + j * 8, p0 - q0, p1 - q1);
      }
      lo += step;
      hi -= step;
    }
    return true;
  } else if constexpr (V == IN_NN) {
    const int k0 = tile_m * HALF_PER_TILE + wm * 32, c0 = tile_n * BN + wn * WN;
    if (2 * (k0 + 31) + 1 >= q.n_other || c0 + WN > q.N) return false;
    const long long step = 16ll * q.ldc;
    const long long base = z * q.strideC + (long long)(2 * (k0 + g)) * q.ldc + c0 + 2 * t;
    long long rp = base + (long long)q.par_plus * q.ldc, rq = base + (long long)(1 - q.par_plus) * q.ldc;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        st.s2(rp + j * 8, acc[i][j][0], acc[i][j][1]);
        st.s2(rq + j * 8, acc[i + 4][j][0], acc[i + 4][j][1]);
      }
      rp += step;
      rq += step;
    }
    return true;
  } else {
    const int r0 = tile_m * BM + wm * WM, c0 = tile_n * HALF_PER_TILE + wn * 16;
    if (r0 + WM > q.M) return false;
    const long long step = 8ll * q.ldc;
    if constexpr (V == OUT_NT) {
      if (c0 + 16 > q.half) return false;
      long long lo = (long long)(r0 + rho(g)) * q.ldc + c0 + 2 * t;
      long long mi = (long long)(r0 + rho(g)) * q.ldc + (q.n_fold - 2 - c0 - 2 * t);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const double p0 = acc[i][j][0], p1 = acc[i][j][1], q0 = acc[i][j + 2][0], q1 = acc[i][j + 2][1];
          st.s2(lo + j * 8, p0 + q0, p1 + q1);
          st.s2(mi - j * 8, p1 - q1, p0 - q0);   // mirror columns n-2-c, n-1-c
        }
        lo += step;
        mi += step;
      }
    } else {
      // IN_NT and CPLX_NT: four consecutive output columns 2c .. 2c+3 per (j, t)
      if (2 * (c0 + 16) > q.n_other) return false;
      const int gg = V == CPLX_NT ? rho(g) : g;
      const bool swap = V == IN_NT && q.par_plus != 0;
      long long rb = (long long)(r0 + gg) * q.ldc + 2 * c0 + 4 * t;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const double p0 = acc[i][j][0], p1 = acc[i][j][1], q0 = acc[i][j + 2][0], q1 = acc[i][j + 2][1];
          st.s2(rb + j * 16, swap ? q0 : p0, swap ? p0 : q0);
          st.s2(rb + j * 16 + 2, swap ? q1 : p1, swap ? p1 : q1);
        }
        rb += step;
      }
    }
    return true;
  }
}

template <int V, class ST>
JFX_HD void epilogue(const Args& q, int tile_m, int tile_n, long long z, int wm, int wn, int g, int t,
                     const double (&acc)[8][4][2], ST&& st) {
#ifndef JFX_FOLD_NO_INTERIOR
  if (epilogue_interior<V>(q, tile_m, tile_n, z, wm, wn, g, t, acc, st)) return;
#endif
  if constexpr (V == CPLX_NT) {
    // rows follow the OUT_NT main loop (permuted by rho); columns: (re, im) pairs = the IN_NT interleave with par_plus = 0
    const bool vec = q.vec_ok != 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = tile_m * BM + wm * WM + i * 8 + rho(g);
      if (row >= q.M) continue;
      const long long rb = (long long)row * q.ldc;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int k0 = 2 * (tile_n * HALF_PER_TILE + wn * 16 + j * 8 + 2 * t);
        put2(st, vec, rb + k0, acc[i][j][0], acc[i][j + 2][0], k0 < q.n_other, k0 + 1 < q.n_other);
        put2(st, vec, rb + k0 + 2, acc[i][j][1], acc[i][j + 2][1], k0 + 2 < q.n_other, k0 + 3 < q.n_other);
      }
    }
    return;
  }
  const bool vec = q.vec_ok != 0;
  const int pp = q.par_plus, pq = 1 - q.par_plus;
  if constexpr (V == OUT_NN) {
    // rows: mirror pairs of the node index; cols: inner
    const long long base = z * q.strideC;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int jidx = tile_m * HALF_PER_TILE + wm * 32 + i * 8 + g;
      if (jidx >= q.half) continue;
      const long long lo = base + (long long)jidx * q.ldc, hi = base + (long long)(q.n_fold - 1 - jidx) * q.ldc;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = tile_n * BN + wn * WN + j * 8 + 2 * t;
        const bool ok0 = col < q.N, ok1 = col + 1 < q.N;
        const double p0 = acc[i][j][0], p1 = acc[i][j][1], q0 = acc[i + 4][j][0], q1 = acc[i + 4][j][1];
        put2(st, vec, lo + col, p0 + q0, p1 + q1, ok0, ok1);
        put2(st, vec, hi + col, p0 - q0, p1 - q1, ok0, ok1);
      }
    }
  } else if constexpr (V == IN_NN) {
    // rows: mode index k = 2 * kidx + parity; cols: inner
    const long long base = z * q.strideC;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int kidx = tile_m * HALF_PER_TILE + wm * 32 + (i & 3) * 8 + g;
      const int k = 2 * kidx + ((i >> 2) ? pq : pp);
      if (k >= q.n_other) continue;
      const long long row = base + (long long)k * q.ldc;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = tile_n * BN + wn * WN + j * 8 + 2 * t;
        put2(st, vec, row + col, acc[i][j][0], acc[i][j][1], col < q.N, col + 1 < q.N);
      }
    }
  } else if constexpr (V == OUT_NT) {
    // rows: data lines (permuted by rho inside an m-tile); cols: mirror pairs of the node index
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = tile_m * BM + wm * WM + i * 8 + rho(g);
      if (row >= q.M) continue;
      const long long rb = (long long)row * q.ldc;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int c = tile_n * HALF_PER_TILE + wn * 16 + j * 8 + 2 * t;
        const bool ok0 = c < q.half, ok1 = c + 1 < q.half;
        const double p0 = acc[i][j][0], p1 = acc[i][j][1], q0 = acc[i][j + 2][0], q1 = acc[i][j + 2][1];
        put2(st, vec, rb + c, p0 + q0, p1 + q1, ok0, ok1);
        // mirror columns n-1-c (value 0) and n-2-c (value 1): stored ascending as (n-2-c, n-1-c)
        const long long m = rb + (q.n_fold - 2 - c);
        put2(st, vec, m, p1 - q1, p0 - q0, ok1, ok0);
      }
    }
  } else {
    // rows: data lines; cols: mode index k = 2 * kidx + parity -> four consecutive columns 2c .. 2c+3
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = tile_m * BM + wm * WM + i * 8 + g;
      if (row >= q.M) continue;
      const long long rb = (long long)row * q.ldc;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int c = tile_n * HALF_PER_TILE + wn * 16 + j * 8 + 2 * t;
        const double p0 = acc[i][j][0], p1 = acc[i][j][1], q0 = acc[i][j + 2][0], q1 = acc[i][j + 2][1];
        const double e0 = pp ? q0 : p0, o0 = pp ? p0 : q0, e1 = pp ? q1 : p1, o1 = pp ? p1 : q1;
        const int k0 = 2 * c;
        put2(st, vec, rb + k0, e0, o0, k0 < q.n_other, k0 + 1 < q.n_other);
        put2(st, vec, rb + k0 + 2, e1, o1, k0 + 2 < q.n_other, k0 + 3 < q.n_other);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Scatter epilogue of the NT variants: the contraction writes its result straight into the receive buffers of the slab
// exchange on every GPU (peer-mapped pointers over NVLink), so that pack + all-to-all (+ unpack) of sharding.py:83-89 of
// the reference become the stores of this kernel.  The output rows of the pass are (a, b) = (row / B, row % B):
//   mode 1 (spectral -> physical, split b): rank p = b / (B/P) receives the row at  [src * A + a][b % (B/P)]   of [P*A][B/P][cols]
//   mode 2 (physical -> spectral, split a): rank p = a / (A/P) receives the row at  [a % (A/P)][src * B + b]    of [A/P][P*B][cols]
// i.e. exactly what lax.all_to_all(split_axis, concat_axis, tiled=True) leaves on rank p (mode 2 includes the unpack).
// ------------------------------------------------------------------------------------------------
constexpr int MAX_PEERS = 8;
struct Scatter {
  int mode, parts, src, A, B;
  int pad_;
  double* peer[MAX_PEERS];
};

JFX_HD void scatter_row(const Scatter& s, int row, int* dest, long long* drow) {
  const int a = row / s.B, b = row - a * s.B;
  if (s.mode == 1) {
    const int bp = s.B / s.parts;
    *dest = b / bp;
    *drow = ((long long)s.src * s.A + a) * bp + (b - *dest * bp);
  } else {
    const int ap = s.A / s.parts;
    *dest = a / ap;
    *drow = (long long)(a - *dest * ap) * ((long long)s.B * s.parts) + (long long)s.src * s.B + b;
  }
}

// mk(dest) returns the store object (s1 / s2 as in epilogue) of rank `dest`'s buffer
template <int V, class MK>
JFX_HD void epilogue_scatter(const Args& q, const Scatter& sc, int tile_m, int tile_n, int wm, int wn, int g, int t,
                             const double (&acc)[8][4][2], MK&& mk) {
  static_assert(V == OUT_NT || V == IN_NT, "scatter epilogue: NT variants only");
  const bool vec = q.vec_ok != 0;
  const int pp = q.par_plus;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = tile_m * BM + wm * WM + i * 8 + (V == OUT_NT ? rho(g) : g);
    if (row >= q.M) continue;
    int dest;
    long long drow;
    scatter_row(sc, row, &dest, &drow);
    auto st = mk(dest);
    const long long rb = drow * q.ldc;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int c = tile_n * HALF_PER_TILE + wn * 16 + j * 8 + 2 * t;
      const double p0 = acc[i][j][0], p1 = acc[i][j][1], q0 = acc[i][j + 2][0], q1 = acc[i][j + 2][1];
      if constexpr (V == OUT_NT) {
        const bool ok0 = c < q.half, ok1 = c + 1 < q.half;
        put2(st, vec, rb + c, p0 + q0, p1 + q1, ok0, ok1);
        put2(st, vec, rb + (q.n_fold - 2 - c), p1 - q1, p0 - q0, ok1, ok0);
      } else {
        const double e0 = pp ? q0 : p0, o0 = pp ? p0 : q0, e1 = pp ? q1 : p1, o1 = pp ? p1 : q1;
        const int k0 = 2 * c;
        put2(st, vec, rb + k0, e0, o0, k0 < q.n_other, k0 + 1 < q.n_other);
        put2(st, vec, rb + k0 + 2, e1, o1, k0 + 2 < q.n_other, k0 + 3 < q.n_other);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// TMA copies of one pipeline stage.  issue(map, dst_offset_in_doubles, rank, c0, ..., c4); map 0 = "A" tensor map
// (K-contiguous operand of the GEMM), 1 = "B" tensor map.  kt = k-tile of the folded problem.
// ------------------------------------------------------------------------------------------------
// issue(map, dst_offset_in_doubles, rank, c0, c1, c2, c3, c4)
template <int V, class ISSUE>
JFX_HD void stage_copies(const Args& q, int kt, int tile_m, int tile_n, int z, ISSUE&& issue) {
  const int m0 = tile_m * BM, n0 = tile_n * BN, k0 = kt * BK;
  // correction tiles (kt >= kts_main): table columns simply continue after the padded main part; the data are the
  // coefficients k' = kcorr0 + ... again, with the parity roles swapped
  const bool corr = kt >= q.kts_main;
  const int kx = corr ? q.kcorr0 + (kt - q.kts_main) * BK : k0;
  if constexpr (V == OUT_NN) {
    const int pp = corr ? 1 - q.par_plus : q.par_plus;
    issue(0, 0, 2, k0, m0, 0, 0, 0);
    if (q.xwide) {
      // dims (n_lo = 8, k', n_hi, parity, batch): one box = the 16 sub-tiles [16 k][8 n] of an X tile, in that order
      issue(1, TILE, 5, 0, kx, n0 >> 3, pp, z);
      issue(1, 2 * TILE, 5, 0, kx, n0 >> 3, 1 - pp, z);
    } else {
#pragma unroll
      for (int sub = 0; sub < BN / 8; ++sub) {
        issue(1, TILE + sub * 128, 4, n0 + 8 * sub, pp, kx, z, 0);
        issue(1, 2 * TILE + sub * 128, 4, n0 + 8 * sub, 1 - pp, kx, z, 0);
      }
    }
  } else if constexpr (V == IN_NN) {
    issue(0, 0, 2, k0, m0, 0, 0, 0);
    if (q.xwide) {
      // dims (n_lo = 8, k, n_hi, batch)
      issue(1, TILE, 4, 0, k0, n0 >> 3, z, 0);
      issue(1, 2 * TILE, 4, 0, q.n_fold - 16 - k0, n0 >> 3, z, 0);
    } else {
#pragma unroll
      for (int sub = 0; sub < BN / 8; ++sub) {
        issue(1, TILE + sub * 128, 3, n0 + 8 * sub, k0, z, 0, 0);
        issue(1, 2 * TILE + sub * 128, 3, n0 + 8 * sub, q.n_fold - 16 - k0, z, 0, 0);
      }
    }
  } else if constexpr (V == OUT_NT || V == CPLX_NT) {
    issue(0, 0, 2, 2 * kx, m0, 0, 0, 0);
    issue(0, TILE, 2, 2 * kx + 16, m0, 0, 0, 0);
    issue(1, 2 * TILE, 2, k0, n0, 0, 0, 0);
  } else {
    issue(0, 0, 2, k0, m0, 0, 0, 0);
    issue(0, TILE, 2, q.n_fold - 16 - k0, m0, 0, 0, 0);
    issue(1, 2 * TILE, 2, k0, n0, 0, 0, 0);
  }
}

// ------------------------------------------------------------------------------------------------
// host side: symmetry analysis, folded tables, tensor-map descriptions
// ------------------------------------------------------------------------------------------------
struct MapDesc {
  const void* base;
  int rank;
  unsigned long long dims[5];
  unsigned long long strides_bytes[4];   // strides of dims 1..rank-1
  unsigned box[5];
  int swizzle_bytes;                     // 64 or 128
};

struct FoldInfo {
  int type;       // FoldType
  int par_plus;
  int kcorr0;     // OUT fold: folded mode index k' from which the asymmetry correction applies (-1: none needed)
};

// Relative asymmetry a mode may keep WITHOUT correction: the folded result of an input that excites only this mode
// differs from the reference's by at most this fraction of the mode's own magnitude (the parity bar is 1e-12).
constexpr double TOL_FIX = 3e-13;

// T: [rows][cols] row-major.  Detects the OUT fold (rows mirrored, both extents even) or the IN fold (columns mirrored,
// column count even), with sigma_k strictly alternating.  tol is relative to max |T|.  The reference's nodes are mirror
// images only up to an ulp (fastgl.py:512-544 takes cos(theta) and cos(pi - theta)), which shows in its Vandermonde as an
// asymmetry of 1e-14 (n = 64) to 4e-13 (n = 1024) of max |T|; a table without the symmetry is off by O(1).  build() folds
// the mean of the two mirror images A; what is lost is E = (T - sigma * mirror T) / 2.  Per mode (column of an OUT table, row
// of an IN table) max |E| / max |T| is the error of a single-mode input relative to its own result:
//   <= tol_fix                    : folded as is;
//   OUT, larger, in the tail      : folded, and the modes k' >= kcorr0 get correction k-tiles (exact again, see the header);
//   OUT, more than 1/4 of the modes, or IN : not folded (plain kernel).
// Measured on the host tables: only the Legendre backward Vandermonde needs it (n = 320: 2 modes, 512: 15, 1024: 101 of 1024).
inline FoldInfo analyze(const double* T, int rows, int cols, double tol = 1e-12, double tol_local = 1e-10,
                        double tol_fix = TOL_FIX) {
  FoldInfo none{FOLD_NONE, 0, -1};
  if (rows < 2 || cols < 2) return none;
  double tmax = 0;
  for (long long i = 0; i < (long long)rows * cols; ++i) {
    const double a = std::fabs(T[i]);
    if (!(a <= 1.79e308)) return none;   // NaN / Inf
    if (a > tmax) tmax = a;
  }
  if (tmax == 0) return none;
  const double eps = tol * tmax;
  // Outer bounds first: the asymmetry is below tol of max |T| everywhere, and below tol_local of the scale of the coefficient
  // it belongs to (a table that merely looks symmetric at the scale of its largest mode is refused).
  // OUT: T[rows-1-j][k] = sigma_k T[j][k]
  if (rows % 2 == 0 && cols % 2 == 0) {
    for (int pp = 0; pp < 2; ++pp) {
      bool ok = true;
      int first_bad = cols;
      for (int k = 0; k < cols && ok; ++k) {
        const double s = ((k & 1) == pp) ? 1.0 : -1.0;
        double cmax = 0, dmax = 0;
        for (int j = 0; j < rows; ++j) cmax = std::fmax(cmax, std::fabs(T[(long long)j * cols + k]));
        for (int j = 0; j < rows / 2; ++j)
          dmax = std::fmax(dmax, std::fabs(T[(long long)(rows - 1 - j) * cols + k] - s * T[(long long)j * cols + k]));
        if (dmax > eps || dmax > tol_local * cmax) ok = false;
        if (0.5 * dmax > tol_fix * cmax && k < first_bad) first_bad = k;
      }
      if (!ok) continue;
      const int kfold = cols / 2;
      int kc = first_bad < cols ? first_bad / 2 : -1;
      if (kc >= 0 && (kfold - kc) * 4 > kfold) continue;   // too many modes to correct: the plain kernel is the better choice
      return FoldInfo{FOLD_OUT, pp, kc};
    }
  }
  // IN: T[k][cols-1-j] = sigma_k T[k][j]
  if (cols % 2 == 0) {
    for (int pp = 0; pp < 2; ++pp) {
      bool ok = true;
      for (int k = 0; k < rows && ok; ++k) {
        const double s = ((k & 1) == pp) ? 1.0 : -1.0;
        double rmax = 0, dmax = 0;
        for (int j = 0; j < cols; ++j) rmax = std::fmax(rmax, std::fabs(T[(long long)k * cols + j]));
        for (int j = 0; j < cols / 2; ++j)
          dmax = std::fmax(dmax, std::fabs(T[(long long)k * cols + (cols - 1 - j)] - s * T[(long long)k * cols + j]));
        if (dmax > eps || 0.5 * dmax > tol_fix * rmax) ok = false;
      }
      if (ok) return FoldInfo{FOLD_IN, pp, -1};
    }
  }
  return none;
}

struct FoldedTable {
  int type = FOLD_NONE, par_plus = 0;
  int n_fold = 0, n_other = 0, half = 0, kfold = 0;
  int rows_nn = 0, rows_nt = 0, ld = 0;       // padded row counts of the two layouts, leading dimension (even)
  int kts_main = 0, kts_corr = 0, kcorr0 = 0;  // k-tiles of the folded problem / of the asymmetry correction (columns kts_main * BK ...)
  std::vector<double> nn, nt;                  // folded tables for the NN and NT warp layouts, [rows][ld]
};

// row r' of the NN layout: tile = r' / 128, wm = (r' / 64) & 1, i = (r' / 8) & 7, g = r' & 7
//   -> pair index = tile * 64 + wm * 32 + (i & 3) * 8 + g, group = i >> 2 (0 = plus, 1 = minus)
JFX_HD void nn_row(int r, int* idx, int* grp) {
  const int tile = r >> 7, wm = (r >> 6) & 1, i = (r >> 3) & 7, g = r & 7;
  *idx = tile * HALF_PER_TILE + wm * 32 + (i & 3) * 8 + g;
  *grp = i >> 2;
}
// row c' of the NT layout: tile = c' / 128, wn = (c' / 32) & 3, j = (c' / 8) & 3, g = c' & 7
//   -> pair index = tile * 64 + wn * 16 + (j & 1) * 8 + g, group = j >> 1
JFX_HD void nt_row(int c, int* idx, int* grp) {
  const int tile = c >> 7, wn = (c >> 5) & 3, j = (c >> 3) & 3, g = c & 7;
  *idx = tile * HALF_PER_TILE + wn * 16 + (j & 1) * 8 + g;
  *grp = j >> 1;
}

// T: [rows][cols] as given to the plan (OUT: rows = nodes, cols = modes; IN: rows = modes, cols = nodes)
inline FoldedTable build(const double* T, int rows, int cols, const FoldInfo& fi) {
  FoldedTable f;
  f.type = fi.type;
  f.par_plus = fi.par_plus;
  if (fi.type == FOLD_NONE) return f;
  const bool out = fi.type == FOLD_OUT;
  f.n_fold = out ? rows : cols;
  f.n_other = out ? cols : rows;
  f.half = f.n_fold / 2;
  f.kfold = out ? f.n_other / 2 : f.half;
  const int pairs = out ? f.half : (f.n_other + 1) / 2;     // mirror pairs (OUT) / mode pairs (IN) along the table rows
  const int tiles = (pairs + HALF_PER_TILE - 1) / HALF_PER_TILE;
  f.rows_nn = f.rows_nt = tiles * 128;
  f.kts_main = (f.kfold + BK - 1) / BK;
  const bool corr = out && fi.kcorr0 >= 0 && fi.kcorr0 < f.kfold;
  f.kcorr0 = corr ? fi.kcorr0 : f.kfold;
  f.kts_corr = corr ? (f.kfold - f.kcorr0 + BK - 1) / BK : 0;
  // without correction columns the row is just the folded modes (the TMA unit zero-fills the k tail); with them the main part
  // is padded to whole k-tiles so that the correction columns start at kts_main * BK
  f.ld = corr ? (f.kts_main + f.kts_corr) * BK : ((f.kfold + 1) & ~1);
  if (f.ld < 2) f.ld = 2;
  f.nn.assign((size_t)f.rows_nn * f.ld, 0.0);
  f.nt.assign((size_t)f.rows_nt * f.ld, 0.0);
  for (int layout = 0; layout < 2; ++layout) {
    std::vector<double>& dst = layout == 0 ? f.nn : f.nt;
    for (int r = 0; r < tiles * 128; ++r) {
      int idx, grp;
      if (layout == 0) nn_row(r, &idx, &grp); else nt_row(r, &idx, &grp);
      const int par = grp ? 1 - fi.par_plus : fi.par_plus;
      if (out) {
        if (idx >= f.half) continue;
        // sigma = +1 for the plus group, -1 for the minus group
        const double sg = grp ? -1.0 : 1.0;
        for (int kp = 0; kp < f.kfold; ++kp) {
          const int k = 2 * kp + par;
          dst[(size_t)r * f.ld + kp] = 0.5 * (T[(long long)idx * cols + k] + sg * T[(long long)(rows - 1 - idx) * cols + k]);
        }
        // correction columns: the plus group accumulates E against the MINUS-parity coefficients and vice versa,
        // E[j, k] = (T[j, k] - sigma_k T[n-1-j, k]) / 2
        for (int c = 0; c < (corr ? f.kfold - f.kcorr0 : 0); ++c) {
          const int k = 2 * (f.kcorr0 + c) + (1 - par);
          dst[(size_t)r * f.ld + (size_t)f.kts_main * BK + c] =
              0.5 * (T[(long long)idx * cols + k] + sg * T[(long long)(rows - 1 - idx) * cols + k]);
        }
      } else {
        const int k = 2 * idx + par;
        if (k >= f.n_other) continue;
        const double sg = grp ? -1.0 : 1.0;
        for (int j = 0; j < f.kfold; ++j)
          dst[(size_t)r * f.ld + j] = 0.5 * (T[(long long)k * cols + j] + sg * T[(long long)k * cols + (cols - 1 - j)]);
      }
    }
  }
  return f;
}

// Geometry of one launch.  table = device (or emulated) pointer of the NN / NT folded table, X = input, C = output.
// nn: the array is viewed as [outer][n_in][inner] (inner = real columns, even, > 1); !nn: [outer rows][n_in].
inline bool make_launch(const FoldedTable& f, bool nn, long long outer, long long inner, const double* table,
                        const double* X, double* C, Args* q, MapDesc* mA, MapDesc* mB) {
  if (f.type == FOLD_NONE) return false;
  const bool out = f.type == FOLD_OUT;
  const int n_in = out ? f.n_other : f.n_fold, n_out = out ? f.n_fold : f.n_other;
  Args a{};
  a.variant = nn ? (out ? OUT_NN : IN_NN) : (out ? OUT_NT : IN_NT);
  a.par_plus = f.par_plus;
  a.n_fold = f.n_fold; a.n_other = f.n_other; a.half = f.half; a.kfold = f.kfold;
  a.kts_main = f.kts_main; a.kts_corr = f.kts_corr; a.kcorr0 = f.kcorr0;
  // table columns the tensor map exposes: the folded modes (k tail zero-filled by the TMA unit), or — with correction
  // columns — the whole padded row
  const unsigned long long tcols = f.kts_corr ? (unsigned long long)f.ld : (unsigned long long)f.kfold;
  a.C = C;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al16(table) || !al16(X)) return false;
  if (n_in % 2) return false;                       // TMA row strides must be multiples of 16 bytes
  if (outer <= 0 || outer >= (1ll << 31) || inner >= (1ll << 31)) return false;
  if (nn) {
    if (inner < 2 || (inner & 1)) return false;
    a.M = f.rows_nn; a.N = (int)inner;
    a.tiles_m = f.rows_nn / BM; a.tiles_n = (int)((inner + BN - 1) / BN); a.batch = (int)outer;
    a.ldc = inner; a.strideC = (long long)n_out * inner;
    a.vec_ok = al16(C) ? 1 : 0;                     // inner even -> every row / batch offset is even
    *mA = MapDesc{table, 2, {tcols, (unsigned long long)f.rows_nn, 1, 1},
                  {(unsigned long long)f.ld * 8, 0, 0}, {BK, BM, 1, 1}, 128};
    // Wide X map (inner % 8 == 0): the contiguous axis is split as (n_lo = 8, n_hi = inner / 8) and the box order is
    // (n_lo, k, n_hi), so ONE box copy lands as the 16 sub-tiles [16 k][8 n] the fragment loads expect — 3 copies per
    // stage instead of 33.  Otherwise one box per sub-tile (the k tail / n tail are zero-filled by the TMA unit either way).
    a.xwide = (inner % 8 == 0) ? 1 : 0;
    const unsigned long long I = (unsigned long long)inner;
    if (out) {
      // X_o[k][n] with k = 2 kk + parity
      if (a.xwide)
        *mB = MapDesc{X, 5, {8, (unsigned long long)(n_in / 2), I / 8, 2, (unsigned long long)outer},
                      {I * 16, 64, I * 8, (unsigned long long)n_in * I * 8}, {8, BK, BN / 8, 1, 1}, 64};
      else   // dims (n, parity, kk, batch)
        *mB = MapDesc{X, 4, {I, 2, (unsigned long long)(n_in / 2), (unsigned long long)outer, 1},
                      {I * 8, I * 16, (unsigned long long)n_in * I * 8, 0}, {8, 1, BK, 1, 1}, 64};
    } else {
      if (a.xwide)
        *mB = MapDesc{X, 4, {8, (unsigned long long)n_in, I / 8, (unsigned long long)outer, 1},
                      {I * 8, 64, (unsigned long long)n_in * I * 8, 0}, {8, BK, BN / 8, 1, 1}, 64};
      else
        *mB = MapDesc{X, 3, {I, (unsigned long long)n_in, (unsigned long long)outer, 1, 1},
                      {I * 8, (unsigned long long)n_in * I * 8, 0, 0}, {8, BK, 1, 1, 1}, 64};
    }
  } else {
    a.M = (int)outer; a.N = f.rows_nt;
    a.tiles_m = (int)((outer + BM - 1) / BM); a.tiles_n = f.rows_nt / BN; a.batch = 1;
    a.ldc = n_out; a.strideC = 0;
    a.vec_ok = (al16(C) && n_out % 2 == 0) ? 1 : 0;
    *mA = MapDesc{X, 2, {(unsigned long long)n_in, (unsigned long long)outer, 1, 1},
                  {(unsigned long long)n_in * 8, 0, 0}, {BK, BM, 1, 1}, 128};
    *mB = MapDesc{table, 2, {tcols, (unsigned long long)f.rows_nt, 1, 1},
                  {(unsigned long long)f.ld * 8, 0, 0}, {BK, BN, 1, 1}, 128};
  }
  if ((long long)a.tiles_m * a.tiles_n * a.batch >= (1ll << 31)) return false;   // the kernel counts tiles in 32 bits
  *q = a;
  return true;
}

JFX_HD int ktiles(const Args& q) { return q.kts_main + q.kts_corr; }

// ---- CPLX_NT: complex interleaved rows [outer][n_in] (as 2 n_in doubles) times a real table [n_out][n_in] ------------
struct CplxTable {
  int n_in = 0, n_out = 0, rows_nt = 0, ld = 0;
  std::vector<double> nt;   // [rows_nt][ld]: row c' holds T[idx(c')][:] for both accumulator groups
};

inline CplxTable build_cplx(const double* T, int n_out, int n_in) {
  CplxTable c;
  c.n_in = n_in; c.n_out = n_out;
  c.rows_nt = (n_out + HALF_PER_TILE - 1) / HALF_PER_TILE * 128;
  c.ld = (n_in + 1) & ~1;
  if (c.ld < 2) c.ld = 2;
  c.nt.assign((size_t)c.rows_nt * c.ld, 0.0);
  for (int r = 0; r < c.rows_nt; ++r) {
    int idx, grp;
    nt_row(r, &idx, &grp);
    if (idx >= n_out) continue;
    for (int k = 0; k < n_in; ++k) c.nt[(size_t)r * c.ld + k] = T[(long long)idx * n_in + k];
  }
  return c;
}

inline bool make_launch_cplx(const CplxTable& c, long long outer, const double* table, const double* X, double* C, Args* q,
                             MapDesc* mA, MapDesc* mB) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al16(table) || !al16(X) || !al16(C)) return false;
  if (outer <= 0 || outer >= (1ll << 31) || c.n_in < 1 || c.n_out < 1) return false;
  Args a{};
  a.variant = CPLX_NT;
  a.par_plus = 0;
  a.n_fold = 0; a.half = 0;
  a.n_other = 2 * c.n_out;            // output columns (doubles) per row
  a.kfold = c.n_in;                   // complex coefficients per row = reduction length per accumulator group
  a.kts_main = (c.n_in + BK - 1) / BK; a.kts_corr = 0; a.kcorr0 = c.n_in;
  a.C = C;
  a.M = (int)outer; a.N = c.rows_nt;
  a.tiles_m = (int)((outer + BM - 1) / BM); a.tiles_n = c.rows_nt / BN; a.batch = 1;
  a.ldc = 2ll * c.n_out; a.strideC = 0;
  a.vec_ok = 1;                       // complex elements are 16 bytes: every (re, im) pair is an aligned double2
  *mA = MapDesc{X, 2, {(unsigned long long)(2 * c.n_in), (unsigned long long)outer, 1, 1},
                {(unsigned long long)c.n_in * 16, 0, 0}, {BK, BM, 1, 1}, 128};
  *mB = MapDesc{table, 2, {(unsigned long long)c.n_in, (unsigned long long)c.rows_nt, 1, 1},
                {(unsigned long long)c.ld * 8, 0, 0}, {BK, BN, 1, 1}, 128};
  *q = a;
  return true;
}

}  // namespace fold
}  // namespace dmma
}  // namespace jfx
