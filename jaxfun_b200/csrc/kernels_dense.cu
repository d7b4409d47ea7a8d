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

#include "jfx_common.h"

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
constexpr int STAGES = 3;
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

struct Params {
  const double* A;  // [M, K] row-major (lda)
  const double* B;  // NT: [N, K] row-major (ldb); NN: [K, N] row-major (ldb)
  double* C;        // [M, N] row-major (ldc)
  int M, N, K;
  int64_t lda, ldb, ldc;
  int64_t strideA, strideB, strideC;  // per batch (blockIdx.z)
};

template <bool NN>
__global__ void __launch_bounds__(THREADS, 1) dgemm_dmma(const Params p) {
  extern __shared__ __align__(16) double smem[];
  double* As = smem;
  double* Bs = smem + STAGES * A_STAGE;
  constexpr int B_STAGE = NN ? B_STAGE_NN : B_STAGE_NT;

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;  // 2 x 4 warps
  const int g = lane >> 2, q = lane & 3;

  // blockIdx.x walks N tiles fastest so CTAs sharing an A panel run together (L2 reuse)
  const int bn0 = blockIdx.x * BN;
  const int bm0 = blockIdx.y * BM;
  const double* A = p.A + (int64_t)blockIdx.z * p.strideA;
  const double* B = p.B + (int64_t)blockIdx.z * p.strideB;
  double* C = p.C + (int64_t)blockIdx.z * p.strideC;

  const int ktiles = (p.K + BK - 1) / BK;

  auto load_stage = [&](int stage, int kt) {
    const int k0 = kt * BK;
    // A tile: BM rows x 16 doubles = 8 chunks of 16 B per row
    double* as = As + stage * A_STAGE;
#pragma unroll
    for (int it = 0; it < (BM * BK / 2) / THREADS; ++it) {
      const int c = tid + it * THREADS;
      const int row = c >> 3, kc = (c & 7) * 2;
      const bool ok = (bm0 + row < p.M) && (k0 + kc < p.K);
      const double* src = ok ? A + (int64_t)(bm0 + row) * p.lda + k0 + kc : A;
      cp_async16(as + row * LDA + kc, src, ok);
    }
    double* bs = Bs + stage * B_STAGE;
    if (!NN) {
#pragma unroll
      for (int it = 0; it < (BN * BK / 2) / THREADS; ++it) {
        const int c = tid + it * THREADS;
        const int row = c >> 3, kc = (c & 7) * 2;
        const bool ok = (bn0 + row < p.N) && (k0 + kc < p.K);
        const double* src = ok ? B + (int64_t)(bn0 + row) * p.ldb + k0 + kc : B;
        cp_async16(bs + row * LDB_NT + kc, src, ok);
      }
    } else {
      // B tile: BK rows x 128 doubles = 64 chunks per row
#pragma unroll
      for (int it = 0; it < (BK * BN / 2) / THREADS; ++it) {
        const int c = tid + it * THREADS;
        const int row = c >> 6, nc = (c & 63) * 2;
        const bool ok = (k0 + row < p.K) && (bn0 + nc < p.N);
        const double* src = ok ? B + (int64_t)(k0 + row) * p.ldb + bn0 + nc : B;
        cp_async16(bs + row * LDB_NN + nc, src, ok);
      }
    }
  };

  double acc[WM / 8][WN / 8][2];
#pragma unroll
  for (int i = 0; i < WM / 8; ++i)
#pragma unroll
    for (int j = 0; j < WN / 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  // prologue
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < ktiles) load_stage(s, s);
    cp_async_commit();
  }

  for (int kt = 0; kt < ktiles; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    // prefetch tile kt + STAGES - 1 into the slot freed at iteration kt - 1
    {
      const int nk = kt + STAGES - 1;
      if (nk < ktiles) load_stage(nk % STAGES, nk);
      cp_async_commit();
    }
    const double* as = As + (kt % STAGES) * A_STAGE + (wm * WM + g) * LDA + q;
    const double* bs = Bs + (kt % STAGES) * B_STAGE;
#pragma unroll
    for (int kk = 0; kk < BK / 4; ++kk) {
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
  if (!attr_set) {
    JFX_CUDA_OK(cudaFuncSetAttribute(dgemm_dmma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem_bytes(false)));
    JFX_CUDA_OK(cudaFuncSetAttribute(dgemm_dmma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem_bytes(true)));
    attr_set = true;
  }
  const int64_t inner = g.inner * (dtype == JFX_C128 ? 2 : 1);
  Params p{};
  if (inner == 1) {
    p.A = (const double*)in; p.B = (const double*)table; p.C = (double*)out;
    p.M = (int)g.outer; p.N = g.n_out; p.K = g.n_in;
    JFX_REQUIRE(g.outer < (1ll << 31), JFX_ERR_UNSUPPORTED, "outer extent too large");
    p.lda = g.n_in; p.ldb = g.n_in; p.ldc = g.n_out;
    dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, 1);
    JFX_REQUIRE(grid.y <= 65535, JFX_ERR_UNSUPPORTED, "too many row tiles (%u)", grid.y);
    dgemm_dmma<false><<<grid, THREADS, smem_bytes(false), s>>>(p);
  } else {
    JFX_REQUIRE(inner < (1ll << 31), JFX_ERR_UNSUPPORTED, "inner extent too large");
    p.A = (const double*)table; p.B = (const double*)in; p.C = (double*)out;
    p.M = g.n_out; p.N = (int)inner; p.K = g.n_in;
    p.lda = g.n_in; p.ldb = inner; p.ldc = inner;
    p.strideA = 0; p.strideB = (int64_t)g.n_in * inner; p.strideC = (int64_t)g.n_out * inner;
    JFX_REQUIRE(g.outer <= 65535, JFX_ERR_UNSUPPORTED, "outer extent %lld > 65535 on a non-last axis",
                (long long)g.outer);
    dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, (unsigned)g.outer);
    dgemm_dmma<true><<<grid, THREADS, smem_bytes(true), s>>>(p);
  }
  JFX_CUDA_OK(cudaGetLastError());
  return JFX_OK;
}

int launch_table_apply(cudaStream_t s, const AxisGeom& g, int dtype, const void* table,
                       bool table_complex, const void* in, void* out, int* used_dmma) {
  if (used_dmma) *used_dmma = 0;
  if (g.outer * g.inner * g.n_out == 0) return JFX_OK;
  if (table_apply_uses_dmma(g, dtype, table_complex) &&
      (g.inner * (dtype == JFX_C128 ? 2 : 1) == 1 ? g.outer < (1ll << 31) : g.outer <= 65535)) {
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
static int time_peak(cudaStream_t s, K kernel, int iters, double flops_per_thread_iter, double* tflops) {
  double* dummy = nullptr;
  JFX_CUDA_OK(cudaMalloc(&dummy, 8));
  cudaEvent_t e0, e1;
  JFX_CUDA_OK(cudaEventCreate(&e0));
  JFX_CUDA_OK(cudaEventCreate(&e1));
  const int blocks = 148 * 8, threads = 256;
  kernel<<<blocks, threads, 0, s>>>(dummy, iters);  // warm-up
  JFX_CUDA_OK(cudaEventRecord(e0, s));
  kernel<<<blocks, threads, 0, s>>>(dummy, iters);
  JFX_CUDA_OK(cudaEventRecord(e1, s));
  JFX_CUDA_OK(cudaEventSynchronize(e1));
  float ms = 0;
  JFX_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
  *tflops = flops_per_thread_iter * iters * (double)blocks * threads / (ms * 1e-3) / 1e12;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(dummy);
  return JFX_OK;
}

int calibrate_dmma(cudaStream_t s, int iters, double* tflops) {
  // per warp-instruction 8*8*4 FMAs = 512 flops -> 16 flops per thread per mma; 16 mma per iter
  return time_peak(s, dmma_peak_kernel, iters, 16.0 * 16.0, tflops);
}
int calibrate_dfma(cudaStream_t s, int iters, double* tflops) {
  return time_peak(s, dfma_peak_kernel, iters, 16.0 * 2.0, tflops);
}

}  // namespace jfx
