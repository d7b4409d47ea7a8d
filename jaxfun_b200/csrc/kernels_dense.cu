// Dense per-axis table contraction: out[o, r, i] = sum_c T[r, c] * in[o, c, i].
//
// This is the reference's `(u * wj) @ conj(Pi)` (galerkin/orthogonal.py:277) and, with a
// precomputed Vandermonde, its recurrence `backward` (galerkin/Jacobi.py:65-110), applied along an
// arbitrary tensor axis without materialising a transpose (what `_build_local_apply_fn`,
// sharding.py:24-40, gets from jit(vmap)).
//
// Two kernels:
//   * table_apply_generic  — any dtype / any shape / real or complex table.  Correctness net for
//     odd sizes, fp32 and complex tables; one thread per output element.
//   * dgemm_dmma           — fp64 real table, FP64 tensor-core (mma.sync m8n8k4 -> SASS DMMA.8x8x4,
//     the only f64 MMA shape sm_100a has), cp.async multi-stage shared-memory pipeline,
//     128x128x16 CTA tile, 64x32 warp tile.  Handles both operand orders:
//        last axis   (inner == 1): C[M=outer, n_out]  = X[M, K] * T^T      ("NT": both K-contiguous)
//        other axes  (inner  > 1): C_o[n_out, inner]  = T[n_out, K] * X_o[K, inner]   ("NN")
//     Complex interleaved data on a non-last axis is the NN case with inner' = 2*inner.
#include <cuda_runtime.h>
#include <cuComplex.h>
#include <stdlib.h>

#include "jfx_common.h"
#include "dmma_params.h"

namespace jfx {

// ------------------------------------------------------------------------------------------
// generic kernel
// ------------------------------------------------------------------------------------------
template <typename T> struct Cx { T re, im; };

template <typename T> __device__ __forceinline__ T zero_of(T) { return T(0); }
template <typename T> __device__ __forceinline__ Cx<T> zero_of(Cx<T>) { return Cx<T>{T(0), T(0)}; }

template <typename T> __device__ __forceinline__ void fma_acc(T& acc, T t, T x) { acc = fma(t, x, acc); }
template <typename T> __device__ __forceinline__ void fma_acc(Cx<T>& acc, T t, Cx<T> x) {
  acc.re = fma(t, x.re, acc.re);
  acc.im = fma(t, x.im, acc.im);
}
template <typename T> __device__ __forceinline__ void fma_acc(Cx<T>& acc, Cx<T> t, Cx<T> x) {
  acc.re = fma(t.re, x.re, acc.re);
  acc.re = fma(-t.im, x.im, acc.re);
  acc.im = fma(t.re, x.im, acc.im);
  acc.im = fma(t.im, x.re, acc.im);
}

// D = data element type, TT = table element type
template <typename D, typename TT>
__global__ void __launch_bounds__(256)
table_apply_generic(const TT* __restrict__ table, const D* __restrict__ in, D* __restrict__ out,
                    int64_t outer, int n_in, int n_out, int64_t inner) {
  const int64_t total = outer * n_out * inner;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx % inner;
    const int64_t r = (idx / inner) % n_out;
    const int64_t o = idx / (inner * n_out);
    const D* x = in + o * n_in * inner + i;
    const TT* t = table + r * n_in;
    D acc = zero_of(D{});
    for (int c = 0; c < n_in; ++c) fma_acc(acc, t[c], x[(int64_t)c * inner]);
    out[idx] = acc;
  }
}

template <typename D, typename TT>
static int run_generic(cudaStream_t s, const AxisGeom& g, const void* table, const void* in,
                       void* out) {
  const int64_t total = g.outer * g.n_out * g.inner;
  if (total == 0) return JFX_OK;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  table_apply_generic<D, TT><<<(unsigned)blocks, 256, 0, s>>>(
      (const TT*)table, (const D*)in, (D*)out, g.outer, g.n_in, g.n_out, g.inner);
  JFX_CUDA_OK(cudaGetLastError());
  return JFX_OK;
}

// ------------------------------------------------------------------------------------------
// FP64 tensor-core kernel
// ------------------------------------------------------------------------------------------
namespace dmma {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int STAGES = 4;
constexpr int THREADS = 256;
constexpr int WM = 64, WN = 32;          // warp tile
constexpr int LDA = BK + 4;              // As[m][k]   pitch (doubles), == 4 mod 16 -> conflict free
constexpr int LDB_NT = BK + 4;           // Bs[n][k]
constexpr int LDB_NN = BN + 4;           // Bs[k][n]
constexpr int A_STAGE = BM * LDA;        // doubles
constexpr int B_STAGE_NT = BN * LDB_NT;
constexpr int B_STAGE_NN = BK * LDB_NN;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool pred) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  int bytes = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void mma_884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}


// WM x WN = warp tile; the CTA tile is always 128 x 128, so THREADS = (128/WM) * (128/WN) * 32:
//   64 x 32 -> 8 warps (2 per SM sub-partition, 64 accumulators per thread),
//   32 x 32 -> 16 warps (4 per sub-partition: more latency tolerance, 0.5 instead of 0.375 LDS per DMMA)
template <bool NN, int WM, int WN>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32, 1) dgemm_dmma(const Params p) {
  constexpr int THREADS = (BM / WM) * (BN / WN) * 32;
  constexpr int WARPS_N = BN / WN;
  extern __shared__ __align__(16) double smem[];
  double* As = smem;
  double* Bs = smem + STAGES * A_STAGE;
  constexpr int B_STAGE = NN ? B_STAGE_NN : B_STAGE_NT;

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int wm = warp / WARPS_N, wn = warp % WARPS_N;
  const int g = lane >> 2, q = lane & 3;

  // blockIdx.x walks N tiles fastest so CTAs sharing an A panel run together (L2 reuse)
  const int bn0 = blockIdx.x * BN;
  const int bm0 = blockIdx.y * BM;
  const double* A = p.A + (int64_t)blockIdx.z * p.strideA;
  const double* B = p.B + (int64_t)blockIdx.z * p.strideB;
  double* C = p.C + (int64_t)blockIdx.z * p.strideC;

  const int ktiles = (p.K + BK - 1) / BK;

  // ---- loader state, hoisted out of the k loop -------------------------------------------------
  // Every thread copies 4 A chunks and 4 B chunks (16 B each) per k-tile.  Their global pointers only
  // advance by a constant per k-tile and their shared-memory offsets never change, so nothing but a
  // pointer bump and the k-bound test is left inside the loop; part `kk` of the copies is issued in
  // front of butterfly step kk so the DMMA stream of a warp is never interrupted by a long burst of
  // address arithmetic right after the barrier (which is when the other warp of the SMSP does the same).
  constexpr int NCH = (BM * BK / 2) / THREADS;           // chunks per thread and operand (4 or 2)
  constexpr int KSTEPS = BK / 4;                         // butterfly steps per k-tile
  static_assert(KSTEPS % NCH == 0, "chunks are spread evenly over the butterfly steps");
  constexpr int RSTEP = THREADS / 8;                     // A / NT-B: 8 chunks per row -> rows between a thread's chunks
  constexpr int NN_RSTEP = THREADS / 64;                 // NN-B: 64 chunks per row
  const int a_row = tid >> 3, a_kc = (tid & 7) * 2;
  const int b_row = NN ? (tid >> 6) : a_row;
  const int b_col = NN ? (tid & 63) * 2 : a_kc;
  const double* a_ptr = A + (int64_t)(bm0 + a_row) * p.lda + a_kc;
  const double* b_ptr = NN ? B + (int64_t)b_row * p.ldb + bn0 + b_col : B + (int64_t)(bn0 + b_row) * p.ldb + b_col;
  const int64_t a_step = (int64_t)RSTEP * p.lda;         // between a thread's chunks
  const int64_t b_step = NN ? (int64_t)NN_RSTEP * p.ldb : (int64_t)RSTEP * p.ldb;
  const int64_t b_adv = NN ? (int64_t)BK * p.ldb : BK;   // per k-tile (A advances by BK)
  const int a_soff = a_row * LDA + a_kc;
  const int b_soff = NN ? b_row * LDB_NN + b_col : b_row * LDB_NT + b_col;
  constexpr int A_SSTEP = RSTEP * LDA, B_SSTEP = NN ? NN_RSTEP * LDB_NN : RSTEP * LDB_NT;
  bool a_ok[NCH], b_ok[NCH];
#pragma unroll
  for (int it = 0; it < NCH; ++it) {
    a_ok[it] = bm0 + a_row + RSTEP * it < p.M;
    b_ok[it] = NN ? (bn0 + b_col < p.N) : (bn0 + b_row + RSTEP * it < p.N);
  }
  int l_k0 = 0;   // k origin of the next tile to load

  auto load_part = [&](int stage, int it) {
    // A chunk `it`
    {
      const bool ok = a_ok[it] && (l_k0 + a_kc < p.K);
      cp_async16(As + stage * A_STAGE + a_soff + it * A_SSTEP, ok ? a_ptr + it * a_step : A, ok);
    }
    {
      const bool kin = NN ? (l_k0 + b_row + NN_RSTEP * it < p.K) : (l_k0 + b_col < p.K);
      const bool ok = b_ok[it] && kin;
      cp_async16(Bs + stage * B_STAGE + b_soff + it * B_SSTEP, ok ? b_ptr + it * b_step : B, ok);
    }
  };
  auto advance = [&]() { a_ptr += BK; b_ptr += b_adv; l_k0 += BK; };

  double acc[WM / 8][WN / 8][2];
#pragma unroll
  for (int i = 0; i < WM / 8; ++i)
#pragma unroll
    for (int j = 0; j < WN / 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  // prologue
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < ktiles) {
#pragma unroll
      for (int it = 0; it < NCH; ++it) load_part(s, it);
      advance();
    }
    cp_async_commit();
  }

  for (int kt = 0; kt < ktiles; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    // tile kt + STAGES - 1 goes into the slot freed at iteration kt - 1, one part per butterfly step
    const int nk = kt + STAGES - 1;
    const bool more = nk < ktiles;
    const int lstage = nk % STAGES;
    const double* as = As + (kt % STAGES) * A_STAGE + (wm * WM + g) * LDA + q;
    const double* bs = Bs + (kt % STAGES) * B_STAGE;
#pragma unroll
    for (int kk = 0; kk < BK / 4; ++kk) {
      if (more && kk % (KSTEPS / NCH) == 0) load_part(lstage, kk / (KSTEPS / NCH));
      double a[WM / 8], b[WN / 8];
#pragma unroll
      for (int i = 0; i < WM / 8; ++i) a[i] = as[i * 8 * LDA + kk * 4];
#pragma unroll
      for (int j = 0; j < WN / 8; ++j) {
        if (!NN) b[j] = bs[(wn * WN + j * 8 + g) * LDB_NT + kk * 4 + q];
        else     b[j] = bs[(kk * 4 + q) * LDB_NN + wn * WN + j * 8 + g];
      }
#pragma unroll
      for (int i = 0; i < WM / 8; ++i)
#pragma unroll
        for (int j = 0; j < WN / 8; ++j) mma_884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    if (more) advance();
    cp_async_commit();
  }
  cp_async_wait<0>();

  // epilogue: each lane owns C[row g][cols 2q, 2q+1] of every 8x8 fragment -> 16 B stores
  const bool vec_ok = ((p.ldc & 1) == 0) && ((((uintptr_t)C) & 15) == 0);
#pragma unroll
  for (int i = 0; i < WM / 8; ++i) {
    const int row = bm0 + wm * WM + i * 8 + g;
    if (row >= p.M) continue;
#pragma unroll
    for (int j = 0; j < WN / 8; ++j) {
      const int col = bn0 + wn * WN + j * 8 + 2 * q;
      double* dst = C + (int64_t)row * p.ldc + col;
      if (vec_ok && col + 1 < p.N) {
        *reinterpret_cast<double2*>(dst) = make_double2(acc[i][j][0], acc[i][j][1]);
      } else {
        if (col < p.N) dst[0] = acc[i][j][0];
        if (col + 1 < p.N) dst[1] = acc[i][j][1];
      }
    }
  }
}

// ---- persistent variant -------------------------------------------------------------------------
// One CTA per SM walks its share of the (batch, M-tile, N-tile) tiles.  The cp.async pipeline runs
// STAGES-1 k-tiles ahead ACROSS tile boundaries, so the loads of the next tile are in flight while the
// epilogue of the current one stores C: the per-tile pipeline fill / drain of the one-tile-per-CTA kernel
// (~20 % of its time at K = 256) disappears.  BK = 32 halves the number of CTA-wide barriers.
namespace pers {
constexpr int BK = 32, STAGES = 3;
constexpr int LDA = BK + 4, LDB_NT = BK + 4, LDB_NN = BN + 4;
constexpr int A_STAGE = BM * LDA, B_STAGE_NT = BN * LDB_NT, B_STAGE_NN = BK * LDB_NN;
constexpr size_t smem_bytes(bool nn) {
  return (size_t)STAGES * (A_STAGE + (nn ? B_STAGE_NN : B_STAGE_NT)) * sizeof(double);
}
}  // namespace pers

struct PParams {
  Params p;
  int tiles_n, tiles_m, batch;   // tile grid; n fastest
};

template <bool NN>
__global__ void __launch_bounds__(THREADS, 1) dgemm_dmma_persistent(const PParams pp) {
  constexpr int BK = pers::BK, STAGES = pers::STAGES, LDA = pers::LDA, LDB_NT = pers::LDB_NT, LDB_NN = pers::LDB_NN;
  constexpr int A_STAGE = pers::A_STAGE, B_STAGE_NT = pers::B_STAGE_NT, B_STAGE_NN = pers::B_STAGE_NN;
  extern __shared__ __align__(16) double smem[];
  double* As = smem;
  double* Bs = smem + STAGES * A_STAGE;
  constexpr int B_STAGE = NN ? B_STAGE_NN : B_STAGE_NT;
  const Params& p = pp.p;

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;  // 2 x 4 warps
  const int g = lane >> 2, q = lane & 3;

  const int ktiles = (p.K + BK - 1) / BK;
  const long long total_tiles = (long long)pp.tiles_n * pp.tiles_m * pp.batch;
  const long long my_tiles = total_tiles > blockIdx.x ? (total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const long long total_iters = my_tiles * ktiles;

  // tile t -> (batch z, m-tile, n-tile); n fastest so CTAs that share an A panel run together
  auto tile_origin = [&](long long t, int& bm0, int& bn0, long long& z) {
    const int tn = (int)(t % pp.tiles_n);
    const long long r = t / pp.tiles_n;
    bm0 = (int)(r % pp.tiles_m) * BM;
    bn0 = tn * BN;
    z = r / pp.tiles_m;
  };

  // ---- loader: per-thread chunk geometry is fixed, pointers / predicates are refreshed once per tile ---
  constexpr int NCH = (BM * BK / 2) / THREADS;           // chunks per thread and operand per k-tile (8)
  constexpr int CPR = BK / 2;                            // 16-byte chunks per A / NT-B row
  constexpr int RSTEP = THREADS / CPR;                   // rows between a thread's chunks (16)
  constexpr int NN_RSTEP = THREADS / (BN / 2);           // NN-B: rows between a thread's chunks (4)
  static_assert(NCH == BK / 4, "one A and one B chunk per butterfly step");
  const int a_row = tid / CPR, a_kc = (tid % CPR) * 2;
  const int b_row = NN ? (tid / (BN / 2)) : a_row;
  const int b_col = NN ? (tid % (BN / 2)) * 2 : a_kc;
  const int64_t a_step = (int64_t)RSTEP * p.lda;
  const int64_t b_step = NN ? (int64_t)NN_RSTEP * p.ldb : (int64_t)RSTEP * p.ldb;
  const int64_t b_adv = NN ? (int64_t)BK * p.ldb : BK;
  const int a_soff = a_row * LDA + a_kc;
  const int b_soff = NN ? b_row * LDB_NN + b_col : b_row * LDB_NT + b_col;
  constexpr int A_SSTEP = RSTEP * LDA, B_SSTEP = NN ? NN_RSTEP * LDB_NN : RSTEP * LDB_NT;

  long long l_it = 0, l_tile = blockIdx.x;
  int l_kt = 0, l_k0 = 0;
  const double *a_ptr = p.A, *b_ptr = p.B;
  unsigned a_okm = 0, b_okm = 0;                         // bit it: chunk `it` is inside the matrix
  auto refresh = [&]() {
    int bm0, bn0;
    long long z;
    tile_origin(l_tile, bm0, bn0, z);
    a_ptr = p.A + z * p.strideA + (int64_t)(bm0 + a_row) * p.lda + a_kc;
    b_ptr = NN ? p.B + z * p.strideB + (int64_t)b_row * p.ldb + bn0 + b_col
               : p.B + z * p.strideB + (int64_t)(bn0 + b_row) * p.ldb + b_col;
    a_okm = b_okm = 0;
#pragma unroll
    for (int it = 0; it < NCH; ++it) {
      if (bm0 + a_row + RSTEP * it < p.M) a_okm |= 1u << it;
      if (NN ? (bn0 + b_col < p.N) : (bn0 + b_row + RSTEP * it < p.N)) b_okm |= 1u << it;
    }
    l_k0 = 0;
  };
  if (my_tiles > 0) refresh();

  auto load_part = [&](int it) {
    if (l_it >= total_iters) return;
    const int stage = (int)(l_it % STAGES);
    {
      const bool ok = ((a_okm >> it) & 1u) && (l_k0 + a_kc < p.K);
      cp_async16(As + stage * A_STAGE + a_soff + it * A_SSTEP, ok ? a_ptr + it * a_step : p.A, ok);
    }
    {
      const bool kin = NN ? (l_k0 + b_row + NN_RSTEP * it < p.K) : (l_k0 + b_col < p.K);
      const bool ok = ((b_okm >> it) & 1u) && kin;
      cp_async16(Bs + stage * B_STAGE + b_soff + it * B_SSTEP, ok ? b_ptr + it * b_step : p.B, ok);
    }
  };
  auto advance = [&]() {   // after the last part of a k-tile
    if (l_it >= total_iters) return;
    ++l_it;
    if (++l_kt == ktiles) {
      l_kt = 0;
      l_tile += gridDim.x;
      if (l_it < total_iters) refresh();
    } else {
      a_ptr += BK; b_ptr += b_adv; l_k0 += BK;
    }
  };

  double acc[WM / 8][WN / 8][2];
#pragma unroll
  for (int i = 0; i < WM / 8; ++i)
#pragma unroll
    for (int j = 0; j < WN / 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
#pragma unroll
    for (int it = 0; it < NCH; ++it) load_part(it);
    advance();
    cp_async_commit();
  }

  // ---- compute cursor ----------------------------------------------------------------------------
  long long c_tile = blockIdx.x;
  int c_kt = 0;
  for (long long it = 0; it < total_iters; ++it) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    const int stage = (int)(it % STAGES);
    const double* as = As + stage * A_STAGE + (wm * WM + g) * LDA + q;
    const double* bs = Bs + stage * B_STAGE;
#pragma unroll
    for (int kk = 0; kk < BK / 4; ++kk) {
      load_part(kk);   // refills the stage consumed at iteration it - 1, one part per butterfly step
      double a[WM / 8], b[WN / 8];
#pragma unroll
      for (int i = 0; i < WM / 8; ++i) a[i] = as[i * 8 * LDA + kk * 4];
#pragma unroll
      for (int j = 0; j < WN / 8; ++j) {
        if (!NN) b[j] = bs[(wn * WN + j * 8 + g) * LDB_NT + kk * 4 + q];
        else     b[j] = bs[(kk * 4 + q) * LDB_NN + wn * WN + j * 8 + g];
      }
#pragma unroll
      for (int i = 0; i < WM / 8; ++i)
#pragma unroll
        for (int j = 0; j < WN / 8; ++j) mma_884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    advance();
    cp_async_commit();
    if (++c_kt == ktiles) {
      // epilogue of this tile (the next tile's loads are already in flight)
      int bm0, bn0;
      long long z;
      tile_origin(c_tile, bm0, bn0, z);
      double* C = p.C + z * p.strideC;
      const bool vec_ok = ((p.ldc & 1) == 0) && ((((uintptr_t)C) & 15) == 0);
#pragma unroll
      for (int i = 0; i < WM / 8; ++i) {
        const int row = bm0 + wm * WM + i * 8 + g;
#pragma unroll
        for (int j = 0; j < WN / 8; ++j) {
          const int col = bn0 + wn * WN + j * 8 + 2 * q;
          double* dst = C + (int64_t)row * p.ldc + col;
          if (row < p.M) {
            if (vec_ok && col + 1 < p.N) {
              *reinterpret_cast<double2*>(dst) = make_double2(acc[i][j][0], acc[i][j][1]);
            } else {
              if (col < p.N) dst[0] = acc[i][j][0];
              if (col + 1 < p.N) dst[1] = acc[i][j][1];
            }
          }
          acc[i][j][0] = acc[i][j][1] = 0.0;
        }
      }
      c_kt = 0;
      c_tile += gridDim.x;
    }
  }
  cp_async_wait<0>();
}

constexpr size_t smem_bytes(bool nn) {
  return (size_t)STAGES * (A_STAGE + (nn ? B_STAGE_NN : B_STAGE_NT)) * sizeof(double);
}

}  // namespace dmma

bool table_apply_uses_dmma(const AxisGeom& g, int dtype, bool table_complex) {
  if (table_complex) return false;
  if (dtype != JFX_F64 && dtype != JFX_C128) return false;
  const int64_t inner = g.inner * (dtype == JFX_C128 ? 2 : 1);
  if (inner == 1) {
    // NT: rows of X and of T must be 16-byte aligned for cp.async
    return (g.n_in % 2 == 0) && g.n_in >= 8;
  }
  // NN: rows of X_o (length inner) must be 16-byte aligned; batch stride n_in*inner is then even too
  return (inner % 2 == 0) && (g.n_in % 2 == 0) && g.n_in >= 8;
}

static int run_dmma(cudaStream_t s, const AxisGeom& g, int dtype, const void* table, const void* in,
                    void* out) {
  using namespace dmma;
  static bool attr_set = false;
  static int sms = 148;
  if (!attr_set) {
    JFX_CUDA_OK(cudaFuncSetAttribute(dgemm_dmma<false, 64, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem_bytes(false)));
    JFX_CUDA_OK(cudaFuncSetAttribute(dgemm_dmma<true, 64, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem_bytes(true)));
    JFX_CUDA_OK(cudaFuncSetAttribute(dgemm_dmma<false, 32, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem_bytes(false)));
    JFX_CUDA_OK(cudaFuncSetAttribute(dgemm_dmma<true, 32, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem_bytes(true)));
    JFX_CUDA_OK(cudaFuncSetAttribute(dgemm_dmma_persistent<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)pers::smem_bytes(false)));
    JFX_CUDA_OK(cudaFuncSetAttribute(dgemm_dmma_persistent<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)pers::smem_bytes(true)));
    int dev = 0;
    JFX_CUDA_OK(cudaGetDevice(&dev));
    JFX_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    attr_set = true;
  }
  static const bool force_pers = [] { const char* e = getenv("JFX_DMMA_PERSISTENT"); return e && e[0] == '1'; }();
  const int64_t inner = g.inner * (dtype == JFX_C128 ? 2 : 1);
  Params p{};
  dim3 grid;
  const bool nn = inner != 1;
  if (!nn) {
    p.A = (const double*)in; p.B = (const double*)table; p.C = (double*)out;
    p.M = (int)g.outer; p.N = g.n_out; p.K = g.n_in;
    JFX_REQUIRE(g.outer < (1ll << 31), JFX_ERR_UNSUPPORTED, "outer extent too large");
    p.lda = g.n_in; p.ldb = g.n_in; p.ldc = g.n_out;
    grid = dim3((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, 1);
  } else {
    JFX_REQUIRE(inner < (1ll << 31), JFX_ERR_UNSUPPORTED, "inner extent too large");
    p.A = (const double*)table; p.B = (const double*)in; p.C = (double*)out;
    p.M = g.n_out; p.N = (int)inner; p.K = g.n_in;
    p.lda = g.n_in; p.ldb = inner; p.ldc = inner;
    p.strideA = 0; p.strideB = (int64_t)g.n_in * inner; p.strideC = (int64_t)g.n_out * inner;
    JFX_REQUIRE(g.outer < (1ll << 31), JFX_ERR_UNSUPPORTED, "outer extent too large");
    grid = dim3((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, (unsigned)g.outer);
  }
  static const bool no_tma = [] { const char* e = getenv("JFX_DMMA_TMA"); return e && e[0] == '0'; }();
  if (!no_tma && !force_pers) {
    const int rc = launch_dmma_tma(s, p, nn, (int)grid.z, sms);
    if (rc < 0) return rc;
    if (rc == 1) return JFX_OK;
  }
  if (force_pers || grid.y > 65535 || grid.z > 65535) {
    PParams pp{p, (int)grid.x, (int)grid.y, (int)grid.z};
    const long long tiles = (long long)grid.x * grid.y * grid.z;
    const unsigned ctas = (unsigned)(tiles < sms ? tiles : sms);
    if (nn) dgemm_dmma_persistent<true><<<ctas, THREADS, pers::smem_bytes(true), s>>>(pp);
    else dgemm_dmma_persistent<false><<<ctas, THREADS, pers::smem_bytes(false), s>>>(pp);
  } else {
    static const bool w16 = [] { const char* e = getenv("JFX_DMMA_WARPS"); return e && atoi(e) == 16; }();
    if (w16) {
      if (nn) dgemm_dmma<true, 32, 32><<<grid, 512, smem_bytes(true), s>>>(p);
      else dgemm_dmma<false, 32, 32><<<grid, 512, smem_bytes(false), s>>>(p);
    } else {
      if (nn) dgemm_dmma<true, 64, 32><<<grid, THREADS, smem_bytes(true), s>>>(p);
      else dgemm_dmma<false, 64, 32><<<grid, THREADS, smem_bytes(false), s>>>(p);
    }
  }
  JFX_CUDA_OK(cudaGetLastError());
  return JFX_OK;
}

int launch_table_apply(cudaStream_t s, const AxisGeom& g, int dtype, const void* table,
                       bool table_complex, const void* in, void* out, int* used_dmma) {
  if (used_dmma) *used_dmma = 0;
  if (g.outer * g.inner * g.n_out == 0) return JFX_OK;
  if (table_apply_uses_dmma(g, dtype, table_complex) &&
      g.outer < (1ll << 31)) {
    if (used_dmma) *used_dmma = 1;
    return run_dmma(s, g, dtype, table, in, out);
  }
  switch (dtype) {
    case JFX_F32:
      JFX_REQUIRE(!table_complex, JFX_ERR_INVALID, "complex table on real data");
      return run_generic<float, float>(s, g, table, in, out);
    case JFX_F64:
      JFX_REQUIRE(!table_complex, JFX_ERR_INVALID, "complex table on real data");
      return run_generic<double, double>(s, g, table, in, out);
    case JFX_C64:
      return table_complex ? run_generic<Cx<float>, Cx<float>>(s, g, table, in, out)
                           : run_generic<Cx<float>, float>(s, g, table, in, out);
    case JFX_C128:
      return table_complex ? run_generic<Cx<double>, Cx<double>>(s, g, table, in, out)
                           : run_generic<Cx<double>, double>(s, g, table, in, out);
  }
  set_error("bad dtype %d", dtype);
  return JFX_ERR_INVALID;
}

// ------------------------------------------------------------------------------------------
// calibration: register-resident FP64 tensor / vector throughput (bench.py roofline denominators)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters) {
  double d[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) { d[i][0] = threadIdx.x * 1e-9; d[i][1] = i * 1e-9; }
  double a = 1.0 + threadIdx.x * 1e-12, b = 1.0 - threadIdx.x * 1e-12;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) dmma::mma_884(d[i][0], d[i][1], a, b);
  }
  double sacc = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) sacc += d[i][0] + d[i][1];
  if (sacc == 123.456) out[0] = sacc;
}

// same, but with the operand pattern of the GEMM inner loop: acc[i][j] += a[i] * b[j], 8 x 4 fragments;
// MODE 1 additionally reloads the fragments from shared memory every step (LDS.64, conflict-free layout),
// MODE 2 also puts a CTA barrier every 8 steps (one k-tile of BK = 32)
template <int MODE>
__global__ void __launch_bounds__(256) dmma_peak_kernel_frag(double* out, int iters) {
  __shared__ double sm[2][128 * 20];
  extern __shared__ __align__(16) double sm2[];
  double d[8][4][2];
  double a[8], b[4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, q = lane & 3, wm = warp >> 2, wn = warp & 3;
  for (int i = threadIdx.x; i < 2 * 128 * 20; i += 256) (&sm[0][0])[i] = 1.0 + i * 1e-12;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = 1.0 + (threadIdx.x + i) * 1e-12;
#pragma unroll
    for (int j = 0; j < 4; ++j) { d[i][j][0] = threadIdx.x * 1e-9; d[i][j][1] = (i + j) * 1e-9; }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) b[j] = 1.0 - (threadIdx.x + j) * 1e-12;
  for (int it = 0; it < iters; ++it) {
    if (MODE >= 1) {
      const double* as = &sm[0][(wm * 64 + g) * 20 + q + (it & 3) * 4];
      const double* bs = &sm[1][(wn * 32 + g) * 20 + q + (it & 3) * 4];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = as[i * 8 * 20];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = bs[j * 8 * 20];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) dmma::mma_884(d[i][j][0], d[i][j][1], a[i], b[j]);
    if (MODE == 2 && (it & 7) == 7) __syncthreads();
    if (MODE >= 3 && (it & 3) == 3) {
      // one k-tile of BK = 16: 8 x 16-byte cp.async per thread (the GEMM's A + B stage), then wait + barrier
      const double* src = out + 1024 + ((size_t)blockIdx.x * 8192 + ((it >> 2) & 3) * 2048 + threadIdx.x * 2) ;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        dmma::cp_async16(&sm2[(((it >> 2) % 3) * 4096) + c * 512 + threadIdx.x * 2], src + c * 512, true);
      dmma::cp_async_commit();
      dmma::cp_async_wait<1>();
      __syncthreads();
    }
  }
  double sacc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) sacc += d[i][j][0] + d[i][j][1];
  if (sacc == 123.456) out[0] = sacc;
}

__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters) {
  double d[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) d[i] = threadIdx.x * 1e-9 + i;
  const double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) d[i] = fma(d[i], a, b);
  }
  double sacc = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) sacc += d[i];
  if (sacc == 123.456) out[0] = sacc;
}

template <typename K>
static int time_peak(cudaStream_t s, K kernel, int iters, double flops_per_thread_iter, double* tflops, int dyn_smem = 0) {
  double* dummy = nullptr;
  JFX_CUDA_OK(cudaMalloc(&dummy, (size_t)(1024 + 148 * 8 * 8192 + 8192) * 8));
  if (dyn_smem) JFX_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_smem));
  cudaEvent_t e0, e1;
  JFX_CUDA_OK(cudaEventCreate(&e0));
  JFX_CUDA_OK(cudaEventCreate(&e1));
  // best over 1, 2 and 8 resident CTAs per SM (8 warps each): the pipes saturate at 2 warps per SM
  // sub-partition, and grids that are not a whole number of waves lose throughput to the tail
  const char* e_ = getenv("JFX_CAL_BLOCKS_PER_SM");
  const int fixed = e_ ? atoi(e_) : 0;
  const int tries[3] = {1, 2, 8};
  double best = 0;
  for (int t = 0; t < 3; ++t) {
    const int per_sm = fixed ? fixed : tries[t];
    const int blocks = 148 * per_sm, threads = 256;
    kernel<<<blocks, threads, dyn_smem, s>>>(dummy, iters);  // warm-up
    JFX_CUDA_OK(cudaEventRecord(e0, s));
    kernel<<<blocks, threads, dyn_smem, s>>>(dummy, iters);
    JFX_CUDA_OK(cudaEventRecord(e1, s));
    JFX_CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0;
    JFX_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    const double tf = flops_per_thread_iter * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
    if (fixed) break;
  }
  *tflops = best;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(dummy);
  return JFX_OK;
}

int calibrate_dmma(cudaStream_t s, int iters, double* tflops) {
  // per warp-instruction 8*8*4 FMAs = 512 flops -> 16 flops per thread per mma; 16 mma per iter
  const char* e_ = getenv("JFX_CAL_FRAG");
  if (e_ && e_[0] == '1') return time_peak(s, dmma_peak_kernel_frag<0>, iters, 16.0 * 32.0, tflops);
  if (e_ && e_[0] == '2') return time_peak(s, dmma_peak_kernel_frag<1>, iters, 16.0 * 32.0, tflops);
  if (e_ && e_[0] == '3') return time_peak(s, dmma_peak_kernel_frag<2>, iters, 16.0 * 32.0, tflops);
  if (e_ && e_[0] == '4') return time_peak(s, dmma_peak_kernel_frag<3>, iters, 16.0 * 32.0, tflops, 3 * 4096 * 8);
  return time_peak(s, dmma_peak_kernel, iters, 16.0 * 16.0, tflops);
}
int calibrate_dfma(cudaStream_t s, int iters, double* tflops) {
  return time_peak(s, dfma_peak_kernel, iters, 16.0 * 2.0, tflops);
}

}  // namespace jfx
