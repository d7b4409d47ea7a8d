// FP64 tensor-core contraction fed by TMA.
//
// Same math and tiling as dgemm_dmma (kernels_dense.cu): C[M,N] (+batch) = A[M,K] * B, 128x128x16 CTA tile,
// 8 MMA warps with 64x32 warp tiles of mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4).  What changes is the operand
// pipeline:
//   * one elected lane issues cp.async.bulk.tensor (TMA) box copies into a 6-stage ring of shared-memory
//     tiles; completion is signalled on per-stage "full" mbarriers (transaction bytes), consumption on
//     per-stage "empty" mbarriers (one arrival per MMA warp);
//   * no CTA-wide barrier and no address arithmetic in the MMA warps' k loop: a warp only waits for the
//     stage it needs, so the eight warps drift instead of stopping together once per k-tile;
//   * tiles land in the 128-byte-swizzled layout (16 doubles per row): fragment loads are conflict-free
//     LDS.64 for both operand orders (128 B swizzle for K-contiguous tiles, 64 B swizzle on 64-byte rows
//     for the N-contiguous operand of the "NN" order);
//   * the grid is persistent (one CTA per SM) and the producer runs ahead across tile boundaries, so the
//     epilogue stores of one tile overlap the loads of the next; out-of-range rows / columns / k are
//     zero-filled by the TMA unit (no predicates anywhere).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdlib.h>

#include "jfx_common.h"
#include "dmma_params.h"

namespace jfx {
namespace dmma {

namespace tma {
constexpr int BK = 16;
constexpr int MMA_WARPS = 8, THREADS = MMA_WARPS * 32;
constexpr int WM = 64, WN = 32;
// CTA tile shapes (always 8 warps of 64 x 32): 128 x 128, and 64 x 256 / 256 x 64 for table extents that are odd
// multiples of 64 (96 -> no gain, 192, 320, ...), where a 128-wide tile would be a quarter padding
template <int SHAPE> struct Tile {
  static constexpr int BM = SHAPE == 0 ? 128 : (SHAPE == 1 ? 64 : 256);
  static constexpr int BN = SHAPE == 0 ? 128 : (SHAPE == 1 ? 256 : 64);
  static constexpr int WARPS_N = BN / WN;
  static constexpr int A_TILE = BM * BK, B_TILE = BN * BK;             // doubles
  static constexpr unsigned STAGE_BYTES = (A_TILE + B_TILE) * sizeof(double);
  static constexpr int STAGES = SHAPE == 0 ? 6 : 5;
  static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 2 * STAGES * 8;
  static_assert((BM / WM) * (BN / WN) == MMA_WARPS, "8 MMA warps");
};
constexpr unsigned SPIN_LIMIT = 1u << 27;

__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n.reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned spins = 0;
  while (!mbar_try(bar, parity)) {
    if (++spins > SPIN_LIMIT) __trap();   // never hang the device: a lost signal becomes a launch failure
  }
}
__device__ __forceinline__ void tma_2d(unsigned dst, const CUtensorMap* tm, int c0, int c1, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_3d(unsigned dst, const CUtensorMap* tm, int c0, int c1, int c2, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void mma_884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
}  // namespace tma

struct TmaArgs {
  int M, N, K;
  double* C;
  int64_t ldc, strideC;
  int tiles_n, tiles_m, batch;
};

// 8 MMA warps, no separate producer warp: registers are allocated in units of 4 warps, so a 9th warp would
// cost a 384-thread budget (170 registers per thread, spills).  Instead lane 0 of warp 0 refills, before
// each k-tile it computes, the stage that was consumed LAG = 2 k-tiles earlier — by then every warp has
// normally released it, so the "empty" wait returns at once and warp 0 is not held up.
template <bool NN, bool PW, int SHAPE>
__global__ void __launch_bounds__(tma::THREADS + (PW ? 32 : 0), 1)
dgemm_dmma_tma(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TmaArgs q) {
  using namespace tma;
  using TL = Tile<SHAPE>;
  constexpr int BM = TL::BM, BN = TL::BN, STAGES = TL::STAGES, A_TILE = TL::A_TILE;
  constexpr unsigned STAGE_BYTES = TL::STAGE_BYTES;
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte aligned tile ring (the 128 B swizzle is a function of address bits 4..9)
  const unsigned base = (s32(smem_raw) + 1023u) & ~1023u;
  unsigned char* gbase = smem_raw + (base - s32(smem_raw));
  const unsigned bar_full = base + STAGES * STAGE_BYTES;       // STAGES x 8 B
  const unsigned bar_empty = bar_full + STAGES * 8;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, MMA_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int ktiles = (q.K + BK - 1) / BK;
  const long long total_tiles = (long long)q.tiles_n * q.tiles_m * q.batch;
  const long long my_tiles = total_tiles > blockIdx.x ? (total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const long long total_iters = my_tiles * ktiles;

  // ---- load cursor (warp 0 only; lane 0 issues) ---------------------------------------------------
  long long l_it = 0, l_tile = blockIdx.x;
  int l_kt = 0, l_m0 = 0, l_n0 = 0, l_z = 0;
  auto l_decode = [&]() {
    const int tn = (int)(l_tile % q.tiles_n);
    const long long r = l_tile / q.tiles_n;
    l_m0 = (int)(r % q.tiles_m) * BM;
    l_n0 = tn * BN;
    l_z = (int)(r / q.tiles_m);
  };
  const bool loader = PW ? (warp == MMA_WARPS) : (warp == 0);
  if (loader && my_tiles > 0) l_decode();
  auto load_next = [&]() {   // loader warp, uniform
    if (l_it >= total_iters) return;
    const int s = (int)(l_it % STAGES);
    if (lane == 0) {
      const unsigned ph = (unsigned)((l_it / STAGES) & 1);
      mbar_wait(bar_empty + 8 * s, ph ^ 1u);
      const unsigned full = bar_full + 8 * s;
      mbar_expect_tx(full, STAGE_BYTES);
      const unsigned a_dst = base + s * STAGE_BYTES, b_dst = a_dst + A_TILE * sizeof(double);
      const int k0 = l_kt * BK;
      tma_2d(a_dst, &tmA, k0, l_m0, full);                   // K-contiguous operand: [128 rows][16 k]
      if (!NN) {
        tma_2d(b_dst, &tmB, k0, l_n0, full);
      } else {
#pragma unroll
        for (int sub = 0; sub < BN / 8; ++sub)               // N-contiguous operand: 16 boxes [16 k][8 n]
          tma_3d(b_dst + sub * 16 * 8 * sizeof(double), &tmB, l_n0 + 8 * sub, k0, l_z, full);
      }
    }
    ++l_it;
    if (++l_kt == ktiles) {
      l_kt = 0;
      l_tile += gridDim.x;
      if (l_it < total_iters) l_decode();
    }
  };
  constexpr int AHEAD = STAGES - 2;   // loads in flight ahead of the compute cursor
  if (PW) {
    if (warp == MMA_WARPS) {   // dedicated producer warp: runs the whole load sequence, paced by the empty barriers
      while (l_it < total_iters) load_next();
      return;
    }
  } else if (warp == 0) {
#pragma unroll 1
    for (int i = 0; i < AHEAD; ++i) load_next();
  }

  // ================================ MMA warps ================================
  const int wm = warp / TL::WARPS_N, wn = warp % TL::WARPS_N;
  const int g = lane >> 2, qd = lane & 3;
  const int nn_x = (g >> 1) ^ (qd >> 1);
  const int nn_base0 = wn * 512 + qd * 8 + (nn_x << 1) + (g & 1);
  const int nn_base1 = wn * 512 + qd * 8 + ((nn_x ^ 2) << 1) + (g & 1);
  double acc[WM / 8][WN / 8][2];
#pragma unroll
  for (int i = 0; i < WM / 8; ++i)
#pragma unroll
    for (int j = 0; j < WN / 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  long long it = 0;
  int prev = -1;
  for (long long t = blockIdx.x; t < total_tiles; t += gridDim.x) {
    for (int kt = 0; kt < ktiles; ++kt, ++it) {
      if (!PW && warp == 0) load_next();   // refills the stage consumed at iteration it - 2
      const int s = (int)(it % STAGES);
      const unsigned ph = (unsigned)((it / STAGES) & 1);
      mbar_wait(bar_full + 8 * s, ph);
      // the previous stage is released here, after the spin loop: every MMA of the previous k-tile has issued by now, so
      // every fragment load of it has completed (an arrive placed at the end of the k-tile is scheduled behind the last
      // LDS *issue* and can take effect before they read; see kernels_dense_fold.cu)
      if (prev >= 0) {
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8 * prev);
      }
      prev = s;
      const double* As = reinterpret_cast<const double*>(gbase + (size_t)s * STAGE_BYTES);
      const double* Bs = As + A_TILE;
#pragma unroll
      for (int kk = 0; kk < BK / 4; ++kk) {
        double a[WM / 8], b[WN / 8];
        // swizzled element (row r, k): r*16 + (((k >> 1) ^ (r & 7)) << 1) + (k & 1); here r & 7 == g
        const int kx = kk * 4 + qd;
        const int col = (((kx >> 1) ^ g) << 1) + (kx & 1);
#pragma unroll
        for (int i = 0; i < WM / 8; ++i) a[i] = As[(wm * WM + i * 8 + g) * BK + col];
#pragma unroll
        for (int j = 0; j < WN / 8; ++j) {
          if (!NN) {
            b[j] = Bs[(wn * WN + j * 8 + g) * BK + col];
          } else {
            // sub-tile of 8 n: [16 k][8 n] with 64-byte rows and the 64 B swizzle (16-byte chunk c>>1 of row k
            // is stored at chunk (c >> 1) ^ ((k >> 1) & 3)): the four k rows of a step then cover all 16
            // banks twice -> 2 wavefronts per LDS.64, the minimum.  (k >> 1) & 3 = ((kk & 1) << 1) | (qd >> 1):
            // two per-thread bases (kk even / odd) plus compile-time offsets.
            b[j] = Bs[((kk & 1) ? nn_base1 : nn_base0) + j * 128 + kk * 32];
          }
        }
#pragma unroll
        for (int i = 0; i < WM / 8; ++i)
#pragma unroll
          for (int j = 0; j < WN / 8; ++j) mma_884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
    }
    // epilogue of this tile: the ring is already being filled for the next one
    const int tn = (int)(t % q.tiles_n);
    const long long r = t / q.tiles_n;
    const int tm_ = (int)(r % q.tiles_m);
    const long long z = r / q.tiles_m;
    const int bm0 = tm_ * BM, bn0 = tn * BN;
    double* C = q.C + z * q.strideC;
    const bool vec_ok = ((q.ldc & 1) == 0) && ((((uintptr_t)C) & 15) == 0);
#pragma unroll
    for (int i = 0; i < WM / 8; ++i) {
      const int row = bm0 + wm * WM + i * 8 + g;
#pragma unroll
      for (int j = 0; j < WN / 8; ++j) {
        const int col = bn0 + wn * WN + j * 8 + 2 * qd;
        double* dst = C + (int64_t)row * q.ldc + col;
        if (row < q.M) {
          if (vec_ok && col + 1 < q.N) {
            *reinterpret_cast<double2*>(dst) = make_double2(acc[i][j][0], acc[i][j][1]);
          } else {
            if (col < q.N) dst[0] = acc[i][j][0];
            if (col + 1 < q.N) dst[1] = acc[i][j][1];
          }
        }
        acc[i][j][0] = acc[i][j][1] = 0.0;
      }
    }
  }
}

// ---- host ----------------------------------------------------------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn encode_fn() {
  static EncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (EncodeFn)p;
  }();
  return fn;
}

static bool encode(CUtensorMap* tm, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                   const cuuint32_t* box, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeFn fn = encode_fn();
  if (!fn) return false;
  const cuuint32_t estr[3] = {1, 1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides_bytes, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <bool NN, int SHAPE>
static int launch_shape(cudaStream_t s, const Params& p, int batch, int sms) {
  using namespace tma;
  using TL = Tile<SHAPE>;
  static bool attr = false;
  if (!attr) {
    JFX_CUDA_OK(cudaFuncSetAttribute(dgemm_dmma_tma<NN, true, SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)TL::SMEM_BYTES));
    JFX_CUDA_OK(cudaFuncSetAttribute(dgemm_dmma_tma<NN, false, SHAPE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)TL::SMEM_BYTES));
    attr = true;
  }
  CUtensorMap tmA, tmB;
  {
    // A: [M, K] row-major, K contiguous -> box {16 k, BM m}
    const cuuint64_t dims[2] = {(cuuint64_t)p.K, (cuuint64_t)p.M};
    const cuuint64_t str[1] = {(cuuint64_t)p.lda * 8};
    const cuuint32_t box[2] = {BK, (cuuint32_t)TL::BM};
    if (!encode(&tmA, p.A, 2, dims, str, box)) return 0;
  }
  if (!NN) {
    const cuuint64_t dims[2] = {(cuuint64_t)p.K, (cuuint64_t)p.N};
    const cuuint64_t str[1] = {(cuuint64_t)p.ldb * 8};
    const cuuint32_t box[2] = {BK, (cuuint32_t)TL::BN};
    if (!encode(&tmB, p.B, 2, dims, str, box)) return 0;
  } else {
    // B: [batch][K][N] with N contiguous -> boxes {8 n, 16 k, 1}, 64 B swizzle
    const cuuint64_t dims[3] = {(cuuint64_t)p.N, (cuuint64_t)p.K, (cuuint64_t)batch};
    const cuuint64_t sb = batch > 1 ? (cuuint64_t)p.strideB * 8 : (cuuint64_t)p.ldb * 8 * (cuuint64_t)p.K;
    const cuuint64_t str[2] = {(cuuint64_t)p.ldb * 8, sb};
    const cuuint32_t box[3] = {8, BK, 1};
    if (!encode(&tmB, p.B, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B)) return 0;
  }
  const int tiles_m = (p.M + TL::BM - 1) / TL::BM, tiles_n = (p.N + TL::BN - 1) / TL::BN;
  TmaArgs q{p.M, p.N, p.K, p.C, p.ldc, p.strideC, tiles_n, tiles_m, batch};
  const long long tiles = (long long)tiles_n * tiles_m * batch;
  const unsigned ctas = (unsigned)(tiles < sms ? tiles : sms);
  // dedicated producer warp (default) or lane 0 of MMA warp 0 (JFX_DMMA_TMA_PW=0)
  static const bool pw = [] { const char* e = getenv("JFX_DMMA_TMA_PW"); return !(e && e[0] == '0'); }();
  if (pw) dgemm_dmma_tma<NN, true, SHAPE><<<ctas, THREADS + 32, TL::SMEM_BYTES, s>>>(tmA, tmB, q);
  else dgemm_dmma_tma<NN, false, SHAPE><<<ctas, THREADS, TL::SMEM_BYTES, s>>>(tmA, tmB, q);
  JFX_CUDA_OK(cudaGetLastError());
  return 1;
}

int launch_dmma_tma(cudaStream_t s, const Params& p, bool nn, int batch, int sms) {
  // envelope: 16-byte aligned bases and strides (cuTensorMapEncodeTiled), no batch stride on the table
  auto al16 = [](const void* x) { return (reinterpret_cast<uintptr_t>(x) & 15) == 0; };
  if (!al16(p.A) || !al16(p.B) || (p.lda & 1) || (p.ldb & 1)) return 0;
  if (p.strideA != 0 && (p.strideA & 1)) return 0;
  if (nn && batch > 1 && (p.strideB & 1)) return 0;
  if (!nn && batch != 1) return 0;
  if (nn && p.strideA != 0) return 0;
  // tile shape: the table extent (M in the NN order, N in the NT order) decides; a 64-wide tile pays when it
  // removes padding (192 -> 3 x 64 instead of 2 x 128) and the other extent is long enough for a 256-wide tile
  const int table_extent = nn ? p.M : p.N, other = nn ? p.N : p.M;
  const int pad128 = (table_extent + 127) / 128 * 128, pad64 = (table_extent + 63) / 64 * 64;
  const bool narrow = pad64 < pad128 && other >= 512;
  if (nn) return narrow ? launch_shape<true, 1>(s, p, batch, sms) : launch_shape<true, 0>(s, p, batch, sms);
  return narrow ? launch_shape<false, 2>(s, p, batch, sms) : launch_shape<false, 0>(s, p, batch, sms);
}

}  // namespace dmma
}  // namespace jfx
