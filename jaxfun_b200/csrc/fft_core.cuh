// Register/shared-memory core of the Stockham FFT used by kernels_fft2.cu and kernels_fused.cu.
#pragma once
#include "fft_common.cuh"

namespace jfx {

// Forward DFT of one line of length N spread over TN = N/E threads (E = Geo<N>::RMAX points each).
//   in : v[bf * R0 + r] = x[jj + r * (N / R0)],  jj = j + bf * TN          (pass-0 input stride)
//   out: v[bf * RL + r] = X[jj + r * (N / RL)],  RL = radix of the last pass
// tw is the table of fast_tables_create: [0, N) the base twiddles W^m, then the per-pass tables of TwLayout<N>, in which the
// twiddles of one butterfly leg are stored by butterfly index: the lanes of a warp (consecutive k) read consecutive 16-byte
// words, where indexing the base table with r * k * TS is a gather with a stride of up to r * TS words per lane (one L1
// wavefront per lane instead of four per warp).
// Sl is the line's exchange buffer (Geo<N>::PITCH elements, skew-padded).  The caller must put a
// barrier between the return of this function and its next write to Sl.
template <typename T, int N, bool WARP_SYNC>
__device__ __forceinline__ void fft_core(Cpx<T>* v, Cpx<T>* __restrict__ Sl, int j, const Cpx<T>* __restrict__ tw) {
  using P = Plan<N>;
  constexpr int R0 = P::R0, R1 = P::R1, R2 = P::R2;
  constexpr int E = Geo<N>::RMAX, TN = N / E, LOGSK = Geo<N>::LOGSK;
  constexpr bool THREE = (R2 > 1);
  constexpr int RL = THREE ? R2 : R1, NSL = N / RL;
  auto sync = [&]() { if (WARP_SYNC) __syncwarp(); else __syncthreads(); };
#ifdef JFX_FFT_SKIP_CORE
  return;   // timing diagnostic only (wrong results): what the loads, stores and addressing cost without the butterflies
#endif
  {
    constexpr int BPT = E / R0;
#pragma unroll
    for (int bf = 0; bf < BPT; ++bf) Dft<T, R0>::run(&v[bf * R0]);
#pragma unroll
    for (int bf = 0; bf < BPT; ++bf) {
      const int jj = j + bf * TN;
      Cpx<T>* dst = Sl + jj * (R0 + 1);                  // sk(jj*R0 + r) = jj*(R0+1) + r
#pragma unroll
      for (int r = 0; r < R0; ++r) dst[r] = v[bf * R0 + r];
    }
  }
  sync();
  if constexpr (THREE) {
    constexpr int R = R1, NS = R0, STR = N / R, BPT = E / R, TS = N / (NS * R);
    constexpr int IN_OFF = STR + (STR >> LOGSK), OUT_OFF = NS + (NS >> LOGSK);
#pragma unroll
    for (int bf = 0; bf < BPT; ++bf) {
      const Cpx<T>* src = Sl + sk<LOGSK>(j + bf * TN);
#pragma unroll
      for (int r = 0; r < R; ++r) v[bf * R + r] = src[r * IN_OFF];
    }
    sync();
#pragma unroll
    for (int bf = 0; bf < BPT; ++bf) {
      const int jj = j + bf * TN, k = jj % NS;
      (void)TS;
      const Cpx<T>* w = tw + TwLayout<N>::OFF_MID + k;     // Wm[(r-1) * NS + k] = W^(r k TS): neighbouring lanes, neighbouring words
#pragma unroll
      for (int r = 1; r < R; ++r) v[bf * R + r] = cmul(v[bf * R + r], w[(r - 1) * NS]);
      Dft<T, R>::run(&v[bf * R]);
      Cpx<T>* dst = Sl + sk<LOGSK>((jj / NS) * NS * R + k);
#pragma unroll
      for (int r = 0; r < R; ++r) dst[r * OUT_OFF] = v[bf * R + r];
    }
    sync();
  }
  {
    constexpr int R = RL, NS = NSL, STR = N / R, BPT = E / R, TS = N / (NS * R);
    constexpr int IN_OFF = STR + (STR >> LOGSK);
    static_assert(TS == 1, "last pass spans the whole transform");
#pragma unroll
    for (int bf = 0; bf < BPT; ++bf) {
      const Cpx<T>* src = Sl + sk<LOGSK>(j + bf * TN);
#pragma unroll
      for (int r = 0; r < R; ++r) v[bf * R + r] = src[r * IN_OFF];
    }
#pragma unroll
    for (int bf = 0; bf < BPT; ++bf) {
      const int k = j + bf * TN;                         // jj < NS: k = jj
      const Cpx<T>* w = tw + TwLayout<N>::OFF_LAST + k;    // Wl[(r-1) * NS + k] = W^(r k)
#pragma unroll
      for (int r = 1; r < R; ++r) v[bf * R + r] = cmul(v[bf * R + r], w[(r - 1) * NS]);
      Dft<T, R>::run(&v[bf * R]);
    }
  }
}

}  // namespace jfx
