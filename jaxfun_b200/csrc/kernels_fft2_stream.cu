// Streaming variant of the FFT / DCT axis pass: persistent CTAs whose NEXT tile is fetched by bulk
// asynchronous copies (cp.async.bulk -> shared memory, completion on an mbarrier) while the butterflies,
// exchanges and stores of the CURRENT tile run.  The plain kernel (kernels_fft2.cu) alternates between a
// load phase and a compute phase inside every CTA and relies on 4 resident CTAs per SM being out of phase;
// here the copy engine keeps ~one tile per CTA in flight all the time, independent of what the warps do.
//
//   tile of a strided axis   : n rows of LPB*16 B (>= 128 B) each        -> n bulk copies, staging [n][LPB]
//   tile of a contiguous axis: LPB lines of n*16 B (or 2*LPB real rows)  -> LPB (2*LPB) bulk copies
//
// Pass 0 reads its points from the staging tile (conflict-free: neighbouring lanes read neighbouring
// 16-byte words), everything after that is the code of kernels_fft2.cu (fft2_tile).  Only full tiles
// without padding / truncation take this path; everything else uses the plain kernel.
#include <cuda_runtime.h>

#include "fft2_tile.cuh"

namespace jfx {
namespace f2 {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

// Same schedule with per-thread 16-byte cp.async (LDGSTS) copies instead of bulk copies: every thread of the CTA moves
// TILE_BYTES / THREADS bytes, fully coalesced in both layouts (a strided tile row is LPB * 16 B >= 128 B), no mbarrier:
// the tile is complete after cp.async.wait_group + one CTA barrier.  JFX_FFT_STREAM=2.
template <typename T, int N, int KIND, int LAY>
__global__ void __launch_bounds__(Cta<N, LAY>::THREADS, Cta<N, LAY>::MINB)
fft2_prefetch_kernel(const __grid_constant__ FftArgs a, long long ntiles) {
  constexpr int THREADS = Cta<N, LAY>::THREADS;
  constexpr int TN = N / Geo<N>::RMAX, LPB = THREADS / TN;
  constexpr unsigned TILE_BYTES = (unsigned)((size_t)LPB * N * sizeof(Cpx<T>));
  constexpr int CHUNKS = TILE_BYTES / 16, CPR = (int)(LPB * sizeof(Cpx<T>) / 16);   // 16-byte chunks per tile / per strided row
  static_assert(CHUNKS % THREADS == 0, "tile splits evenly over the CTA");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* stage = smem_raw;
  Cpx<T>* S = reinterpret_cast<Cpx<T>*>(smem_raw + TILE_BYTES);
  const int tid = threadIdx.x;
  const unsigned inner = (unsigned)a.inner;

  auto issue = [&](long long tile) {
    if (LAY == LAY_STRIDED) {
      const unsigned l0 = (unsigned)tile * LPB, o = l0 / inner, b = l0 - o * inner;
      const unsigned char* src = reinterpret_cast<const unsigned char*>(
          reinterpret_cast<const Cpx<T>*>(a.in) + (size_t)o * N * inner + b);
      const size_t row_bytes = (size_t)inner * sizeof(Cpx<T>);
#pragma unroll
      for (int i = 0; i < CHUNKS / THREADS; ++i) {
        const int c = tid + i * THREADS, row = c / CPR, col = c % CPR;
        cp_async16(stage + (size_t)c * 16, src + row * row_bytes + (size_t)col * 16);
      }
    } else {
      const unsigned char* src = reinterpret_cast<const unsigned char*>(a.in) + (size_t)tile * TILE_BYTES;
#pragma unroll
      for (int i = 0; i < CHUNKS / THREADS; ++i) {
        const int c = tid + i * THREADS;
        cp_async16(stage + (size_t)c * 16, src + (size_t)c * 16);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  long long tile = blockIdx.x;
  if (tile < ntiles) issue(tile);
  for (; tile < ntiles; tile += gridDim.x) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const long long next = tile + gridDim.x;
    auto after_load = [&]() {
      __syncthreads();                       // every thread has its points in registers
      if (next < ntiles) issue(next);
    };
    fft2_tile<T, N, KIND, LAY, false, SRC_STAGED>(a, tile, S, stage, after_load);
    __syncthreads();
  }
}

template <typename T, int N, int KIND, int LAY>
__global__ void __launch_bounds__(Cta<N, LAY>::THREADS, Cta<N, LAY>::MINB)
fft2_stream_kernel(const __grid_constant__ FftArgs a, long long ntiles) {
  constexpr int THREADS = Cta<N, LAY>::THREADS;
  constexpr int TN = N / Geo<N>::RMAX, LPB = THREADS / TN;
  constexpr unsigned TILE_BYTES = (unsigned)((size_t)LPB * N * sizeof(Cpx<T>));
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Cpx<T>* stage = reinterpret_cast<Cpx<T>*>(smem_raw);                     // raw tile as copied
  Cpx<T>* S = reinterpret_cast<Cpx<T>*>(smem_raw + TILE_BYTES);            // exchange buffer
  __shared__ __align__(8) unsigned long long bar;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // warp 0 issues the copies of one tile
  auto issue = [&](long long tile) {
    if (warp != 0) return;
    if (lane == 0) mbar_expect_tx(&bar, TILE_BYTES);
    __syncwarp();
    const long long l0 = tile * LPB;
    if (LAY == LAY_STRIDED) {
      const long long o = l0 / a.inner, b = l0 - o * a.inner;
      const Cpx<T>* src = reinterpret_cast<const Cpx<T>*>(a.in) + (size_t)o * N * a.inner + b;
      for (int i = lane; i < N; i += 32)
        bulk_g2s(stage + (size_t)i * LPB, src + (size_t)i * a.inner, LPB * sizeof(Cpx<T>), &bar);
    } else if (LAY == LAY_CONTIG) {
      const Cpx<T>* src = reinterpret_cast<const Cpx<T>*>(a.in) + (size_t)l0 * N;
      for (int i = lane; i < LPB; i += 32)
        bulk_g2s(stage + (size_t)i * N, src + (size_t)i * N, N * sizeof(Cpx<T>), &bar);
    } else {
      const T* src = reinterpret_cast<const T*>(a.in) + (size_t)(2 * l0) * N;
      T* dst = reinterpret_cast<T*>(stage);
      for (int i = lane; i < 2 * LPB; i += 32)
        bulk_g2s(dst + (size_t)i * N, src + (size_t)i * N, N * sizeof(T), &bar);
    }
  };

  long long tile = blockIdx.x;
  if (tile < ntiles) issue(tile);
  unsigned parity = 0;
  for (; tile < ntiles; tile += gridDim.x) {
    mbar_wait(&bar, parity);
    parity ^= 1u;
    const long long next = tile + gridDim.x;
    auto after_load = [&]() {
      // every thread has its points in registers: the staging tile can take the next copy
      __syncthreads();
      if (next < ntiles) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(next);
      }
    };
    fft2_tile<T, N, KIND, LAY, false, SRC_STAGED>(a, tile, S, stage, after_load);
    __syncthreads();   // exchange buffer and (Chebyshev) output staging are reused by the next tile
  }
}

template <typename T, int N, int KIND, int LAY>
static int launch_prefetch_variant(cudaStream_t s, const FftArgs& a) {
  constexpr int THREADS = Cta<N, LAY>::THREADS;
  constexpr int TN = N / Geo<N>::RMAX, LPB = THREADS / TN;
  const size_t smem = (size_t)LPB * N * sizeof(Cpx<T>) + (size_t)LPB * Geo<N>::PITCH * sizeof(Cpx<T>);
  static int grid = -1;
  if (grid < 0) {
    if (smem > 113 * 1024) { grid = 0; }
    else {
      JFX_CUDA_OK(cudaFuncSetAttribute(fft2_prefetch_kernel<T, N, KIND, LAY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int dev = 0, sms = 0, nb = 0;
      JFX_CUDA_OK(cudaGetDevice(&dev));
      JFX_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      JFX_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fft2_prefetch_kernel<T, N, KIND, LAY>, THREADS, smem));
      grid = nb >= 2 ? sms * nb : 0;
    }
  }
  if (grid == 0) return 0;
  const long long ntiles = a.lines / LPB;
  if (ntiles < 2LL * grid) return 0;
  fft2_prefetch_kernel<T, N, KIND, LAY><<<grid, THREADS, smem, s>>>(a, ntiles);
  JFX_CUDA_OK(cudaGetLastError());
  return 1;
}

template <typename T, int N, int KIND, int LAY>
static int launch_stream_variant(cudaStream_t s, const FftArgs& a) {
  constexpr int THREADS = Cta<N, LAY>::THREADS;
  constexpr int TN = N / Geo<N>::RMAX, LPB = THREADS / TN;
  const size_t smem = (size_t)LPB * N * sizeof(Cpx<T>) + (size_t)LPB * Geo<N>::PITCH * sizeof(Cpx<T>);
  static int grid = -1;
  if (grid < 0) {
    if (smem > 113 * 1024) { grid = 0; }   // fewer than two CTAs per SM: the plain kernel overlaps better
    else {
      JFX_CUDA_OK(cudaFuncSetAttribute(fft2_stream_kernel<T, N, KIND, LAY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int dev = 0, sms = 0, nb = 0;
      JFX_CUDA_OK(cudaGetDevice(&dev));
      JFX_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      JFX_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fft2_stream_kernel<T, N, KIND, LAY>, THREADS, smem));
      grid = nb >= 2 ? sms * nb : 0;
    }
  }
  if (grid == 0) return 0;
  const long long ntiles = a.lines / LPB;
  if (ntiles < 2LL * grid) return 0;       // too small to stream
  fft2_stream_kernel<T, N, KIND, LAY><<<grid, THREADS, smem, s>>>(a, ntiles);
  JFX_CUDA_OK(cudaGetLastError());
  return 1;
}

template <typename T, int N>
static int launch_stream_n(cudaStream_t s, const FftArgs& a, int mode) {
  constexpr int TN = N / Geo<N>::RMAX;
  const int lay = a.real_pair ? LAY_REALPAIR : (a.inner > 1 ? LAY_STRIDED : LAY_CONTIG);
  // envelope: full tiles, no padding / truncation, 16-byte aligned rows
  if (a.n_in != N || a.n_out != N || a.pre) return 0;
  int k4;
  switch (a.kind) {
    case FAST_CHEB_BACKWARD: k4 = K_CHEB_BWD; break;
    case FAST_CHEB_FORWARD: case FAST_CHEB_SCALAR: k4 = K_CHEB_FWD; break;
    case FAST_FOURIER_BACKWARD: k4 = K_FOUR_BWD; break;
    default: k4 = K_FOUR_FWD;
  }
  if ((reinterpret_cast<uintptr_t>(a.in) & 15) != 0) return 0;
#define JFX_CASE(K, L)                                                              \
  if (k4 == K && lay == L) {                                                        \
    constexpr int LPB = Cta<N, L>::THREADS / TN;                                    \
    if (a.lines % LPB) return 0;                                                    \
    if (L == LAY_STRIDED && (a.inner % LPB)) return 0;                              \
    if (L == LAY_REALPAIR && ((a.real_lines & 1) || (N * sizeof(T)) % 16)) return 0; \
    return mode == 2 ? launch_prefetch_variant<T, N, K, L>(s, a) : launch_stream_variant<T, N, K, L>(s, a); \
  }
  JFX_CASE(K_CHEB_BWD, LAY_CONTIG) JFX_CASE(K_CHEB_BWD, LAY_STRIDED) JFX_CASE(K_CHEB_BWD, LAY_REALPAIR)
  JFX_CASE(K_CHEB_FWD, LAY_CONTIG) JFX_CASE(K_CHEB_FWD, LAY_STRIDED) JFX_CASE(K_CHEB_FWD, LAY_REALPAIR)
  JFX_CASE(K_FOUR_BWD, LAY_CONTIG) JFX_CASE(K_FOUR_BWD, LAY_STRIDED)
  JFX_CASE(K_FOUR_FWD, LAY_CONTIG) JFX_CASE(K_FOUR_FWD, LAY_STRIDED)
#undef JFX_CASE
  return 0;
}

}  // namespace f2

// 1 = launched, 0 = outside the envelope (use the plain kernel), < 0 = error.  fp64 only for now.
int launch_fast_axis_stream(cudaStream_t s, const FftArgs& a, int n, bool dbl, int mode) {
  if (!dbl || a.lines <= 0) return 0;
  if (a.lines >= (1ll << 31) || a.inner >= (1ll << 31)) return 0;
#ifdef JFX_FFT2_ONLY
  if (n == JFX_FFT2_ONLY) return f2::launch_stream_n<double, JFX_FFT2_ONLY>(s, a, mode);
  return 0;
#endif
  switch (n) {
    case 64: return f2::launch_stream_n<double, 64>(s, a, mode);
    case 128: return f2::launch_stream_n<double, 128>(s, a, mode);
    case 256: return f2::launch_stream_n<double, 256>(s, a, mode);
    case 512: return f2::launch_stream_n<double, 512>(s, a, mode);
    case 1024: return f2::launch_stream_n<double, 1024>(s, a, mode);
    case 2048: return f2::launch_stream_n<double, 2048>(s, a, mode);
  }
  return 0;
}

}  // namespace jfx
