// Second-generation fast transform kernel: Fourier c2c FFT and Chebyshev DCT-II / DCT-III along one
// tensor axis with the FIRST butterfly pass fed straight from global memory into registers and the
// LAST pass stored straight from registers to global memory.
//
// Reference semantics are those of kernels_fft.cu (galerkin/Fourier.py:126-180,
// galerkin/Chebyshev.py:225-279); this file only changes how the work is mapped onto the SM:
//
//   * a thread owns E = 4..16 points of one line; pass 0 loads them with the Stockham input stride
//     N/R0 — for a contiguous axis the TN = N/E threads of a line read TN consecutive elements per
//     request (>= 128 B), for a strided axis LPB neighbouring lines are read together (>= 128 B);
//   * the register butterflies (radix 4/8/16) exchange data through padded shared memory ONCE per
//     extra pass (2 transits for n <= 256, 4 for n <= 4096) instead of staging the tile in and out;
//   * all index arithmetic is hoisted: per-thread base pointers, compile-time element offsets, and
//     a GEN template flag that compiles the padding / truncation / ragged-tile predicates only
//     into the variant that needs them;
//   * the DCT's spectral-side pairing (k, n-k) reads the partner coefficient directly from global
//     (backward) or through one extra shared-memory exchange (forward); its even/odd sample
//     permutation is folded into the global addresses (free on a strided axis; on a contiguous
//     axis the backward result is permuted through shared memory so stores stay coalesced).
#include <cuda_runtime.h>

#include "fft2_tile.cuh"

namespace jfx {
namespace f2 {

// ---------------------------------------------------------------------------------------------------
template <typename T, int N, int KIND, int LAY, bool GEN>
static int launch_variant(cudaStream_t s, const FftArgs& a) {
  constexpr int THREADS = Cta<N, LAY>::THREADS;
  constexpr int E = Geo<N>::RMAX, TN = N / E, LPB = THREADS / TN;
  const size_t smem = (size_t)LPB * Geo<N>::PITCH * sizeof(Cpx<T>);
  static bool attr_done = false;
  if (!attr_done) {
    JFX_CUDA_OK(cudaFuncSetAttribute(fft2_kernel<T, N, KIND, LAY, GEN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
    attr_done = true;
  }
  const long long blocks = (a.lines + LPB - 1) / LPB;
  fft2_kernel<T, N, KIND, LAY, GEN><<<(unsigned)blocks, THREADS, smem, s>>>(a);
  JFX_CUDA_OK(cudaGetLastError());
  return 1;
}

template <typename T, int N, int KIND, int LAY>
static int launch_gen(cudaStream_t s, const FftArgs& a) {
  constexpr int THREADS = Cta<N, LAY>::THREADS;
  constexpr int E = Geo<N>::RMAX, TN = N / E, LPB = THREADS / TN;
  bool gen = (a.lines % LPB) != 0 || a.n_in != N || a.n_out != N;
  if (LAY == LAY_REALPAIR && (a.real_lines & 1)) gen = true;
  return gen ? launch_variant<T, N, KIND, LAY, true>(s, a) : launch_variant<T, N, KIND, LAY, false>(s, a);
}

template <typename T, int N>
static int launch_n(cudaStream_t s, const FftArgs& a) {
  const int lay = a.real_pair ? LAY_REALPAIR : (a.inner > 1 ? LAY_STRIDED : LAY_CONTIG);
  int k4;
  switch (a.kind) {
    case FAST_CHEB_BACKWARD: k4 = K_CHEB_BWD; break;
    case FAST_CHEB_FORWARD: case FAST_CHEB_SCALAR: k4 = K_CHEB_FWD; break;
    case FAST_FOURIER_BACKWARD: k4 = K_FOUR_BWD; break;
    default: k4 = K_FOUR_FWD;
  }
#define JFX_CASE(K, L) if (k4 == K && lay == L) return launch_gen<T, N, K, L>(s, a);
  JFX_CASE(K_CHEB_BWD, LAY_CONTIG) JFX_CASE(K_CHEB_BWD, LAY_STRIDED) JFX_CASE(K_CHEB_BWD, LAY_REALPAIR)
  JFX_CASE(K_CHEB_FWD, LAY_CONTIG) JFX_CASE(K_CHEB_FWD, LAY_STRIDED) JFX_CASE(K_CHEB_FWD, LAY_REALPAIR)
  JFX_CASE(K_FOUR_BWD, LAY_CONTIG) JFX_CASE(K_FOUR_BWD, LAY_STRIDED)
  JFX_CASE(K_FOUR_FWD, LAY_CONTIG) JFX_CASE(K_FOUR_FWD, LAY_STRIDED)
#undef JFX_CASE
  return 0;
}

template <typename T>
static int launch_t(cudaStream_t s, const FftArgs& a, int n) {
#ifdef JFX_FFT2_ONLY   /* A/B variant builds (tools/build_variant.py): instantiate one length, everything else falls back */
  if (n == JFX_FFT2_ONLY) return launch_n<T, JFX_FFT2_ONLY>(s, a);
  return 0;
#endif
  switch (n) {
    case 16: return launch_n<T, 16>(s, a);
    case 48: return launch_n<T, 48>(s, a);
    case 96: return launch_n<T, 96>(s, a);
    case 192: return launch_n<T, 192>(s, a);
    case 384: return launch_n<T, 384>(s, a);
    case 80: return launch_n<T, 80>(s, a);
    case 160: return launch_n<T, 160>(s, a);
    case 320: return launch_n<T, 320>(s, a);
    case 32: return launch_n<T, 32>(s, a);
    case 64: return launch_n<T, 64>(s, a);
    case 128: return launch_n<T, 128>(s, a);
    case 256: return launch_n<T, 256>(s, a);
    case 512: return launch_n<T, 512>(s, a);
    case 1024: return launch_n<T, 1024>(s, a);
    case 2048: return launch_n<T, 2048>(s, a);
    case 4096: return launch_n<T, 4096>(s, a);
  }
  return 0;
}

}  // namespace f2

int launch_fast_axis_v2(cudaStream_t s, const FftArgs& a, int n, bool dbl) {
  // envelope: 32-bit line arithmetic on strided axes
  if (a.lines <= 0) return 1;
  if (a.lines >= (1ll << 31) || a.inner >= (1ll << 31)) return 0;
  return dbl ? f2::launch_t<double>(s, a, n) : f2::launch_t<float>(s, a, n);
}

}  // namespace jfx
