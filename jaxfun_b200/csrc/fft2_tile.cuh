// Tile body of the second-generation FFT / DCT axis pass (see kernels_fft2.cu for the design notes).
// Shared by the one-pass kernel (kernels_fft2.cu) and the plane-fused two-pass kernel (kernels_fft2_pair.cu).
#pragma once
#include <cuda_runtime.h>

#include "fft_common.cuh"
#include "fft_core.cuh"

namespace jfx {
namespace f2 {

// threads per CTA: small CTAs (several per SM, each in a different load / butterfly / store phase)
// overlap DRAM time with FP64 time; a line needs TN = N/E of them
#ifndef JFX_FFT_REGS16
#define JFX_FFT_REGS16 128
#endif
#ifndef JFX_FFT_REGS8
#define JFX_FFT_REGS8 80
#endif
template <int N, int LAY> struct Cta {
  static constexpr int TN = N / Geo<N>::RMAX;
  // strided axes read LPB neighbouring lines per request: keep LPB * 16 B >= 128 B while the tile fits
#ifndef JFX_LPB4096
#define JFX_LPB4096 2
#endif
  static constexpr int LPB_STRIDED = N <= 512 ? 8 : (N <= 2048 ? 4 : JFX_LPB4096);
  static constexpr int T0 = (LAY == LAY_STRIDED) ? TN * LPB_STRIDED : TN;
#ifdef JFX_PLAN256_884
  static constexpr int TMIN = (N == 256) ? 256 : 128;     // the plane-fused kernel wants one CTA size for both layouts
#else
  static constexpr int TMIN = 128;
#endif
  static constexpr int THREADS = T0 > TMIN ? T0 : TMIN;
  static constexpr int MINB = (65536 / THREADS) / (Geo<N>::RMAX >= 24 ? 168 : Geo<N>::RMAX >= 16 ? JFX_FFT_REGS16 : JFX_FFT_REGS8);   // register budget per thread
};

template <typename T> __device__ __forceinline__ Cpx<T> ldg(const Cpx<T>* p) { return *p; }

// ---- per-thread view of one line ----------------------------------------------------------------
enum { SRC_GLOBAL = 0, SRC_GLOBAL_CG = 1, SRC_STAGED = 2 };
struct NoHook { __device__ __forceinline__ void operator()() const {} };

// SRC: where pass 0 reads from — global memory, global memory through L2 only (ld.global.cg), or a
// shared-memory staging tile that a bulk async copy filled (kernels_fft2_stream.cu); SLPB = lines per
// staging row of a strided tile ([n][SLPB] layout), unused otherwise
template <typename T, int LAY, int SRC, int SLPB> struct LineIO {
  const Cpx<T>* __restrict__ st;    // SRC_STAGED: element 0 of this line in the staging tile
  const T* __restrict__ sr0;        // SRC_STAGED, REALPAIR: the two staged real rows
  const T* __restrict__ sr1;
  const Cpx<T>* __restrict__ cin;   // CONTIG / STRIDED: element 0 of the line
  Cpx<T>* __restrict__ cout;
  const T* __restrict__ r0i;        // REALPAIR: the two real rows packed as (re, im)
  const T* __restrict__ r1i;
  T* __restrict__ r0o;
  T* __restrict__ r1o;
  long long is;                     // STRIDED: element stride along the axis (= inner)
  bool valid, valid1;               // line exists / second real row exists

  __device__ __forceinline__ Cpx<T> load(int idx) const {
    if (SRC == SRC_STAGED) {
      if (LAY == LAY_REALPAIR) return Cpx<T>{sr0[idx], sr1[idx]};
      if (LAY == LAY_STRIDED) return st[idx * SLPB];
      return st[idx];
    }
    if (SRC == SRC_GLOBAL_CG) {
      if (LAY == LAY_REALPAIR) return Cpx<T>{__ldcg(r0i + idx), __ldcg(r1i + idx)};
      const Cpx<T>* p = (LAY == LAY_STRIDED) ? cin + (long long)idx * is : cin + idx;
      if (sizeof(T) == 8) {
        const double2 v = __ldcg(reinterpret_cast<const double2*>(p));
        return Cpx<T>{(T)v.x, (T)v.y};
      } else {
        const float2 v = __ldcg(reinterpret_cast<const float2*>(p));
        return Cpx<T>{(T)v.x, (T)v.y};
      }
    }
    if (LAY == LAY_REALPAIR) return Cpx<T>{r0i[idx], r1i[idx]};
    if (LAY == LAY_STRIDED) return cin[(long long)idx * is];
    return cin[idx];
  }
  template <bool GEN> __device__ __forceinline__ void store(int idx, Cpx<T> v) const {
    if (LAY == LAY_REALPAIR) {
      if (!GEN || valid) r0o[idx] = v.x;
      if (!GEN || valid1) r1o[idx] = v.y;
    } else if (LAY == LAY_STRIDED) {
      if (!GEN || valid) cout[(long long)idx * is] = v;
    } else {
      if (!GEN || valid) cout[idx] = v;
    }
  }
};

// One tile (LPB lines) of an axis pass.  `tile` is the global tile index over all lines of the array,
// S the CTA's exchange buffer (LPB * Geo<N>::PITCH elements).  CG: read the input with ld.global.cg
// (L2 only) — for data produced earlier in the SAME launch by other SMs (fft2_pair.cu).  `after_load`
// runs once all of the tile's input has been consumed into registers (before any other barrier).
template <typename T, int N, int KIND, int LAY, bool GEN, int SRC = SRC_GLOBAL, typename Hook = NoHook>
__device__ __forceinline__ void fft2_tile(const FftArgs& a, long long tile, Cpx<T>* __restrict__ S,
                                          const void* stage = nullptr, Hook after_load = Hook()) {
  constexpr int THREADS = Cta<N, LAY>::THREADS;
  using P = Plan<N>;
  constexpr int R0 = P::R0, R1 = P::R1, R2 = P::R2;
  constexpr int E = Geo<N>::RMAX;            // points per thread
  constexpr int TN = N / E;                  // threads per line
  constexpr int LPB = THREADS / TN;          // lines per block
  constexpr int LOGSK = Geo<N>::LOGSK, PITCH = Geo<N>::PITCH;
  constexpr bool THREE = (R2 > 1);
  constexpr int RL = THREE ? R2 : R1;        // radix of the last pass
  constexpr int NSL = N / RL;                // its output stride; bins of a thread: jj + r * NSL
  constexpr bool CHEB = (KIND == K_CHEB_BWD || KIND == K_CHEB_FWD);
  // lines whose TN threads sit in one warp only need warp-level syncs
  constexpr bool WARP_SYNC = (LAY != LAY_STRIDED) && (TN <= 32);
  static_assert(TN >= 2 && LPB >= 1, "unsupported geometry");

  const int tid = threadIdx.x;
  int ll, j;
  if (LAY == LAY_STRIDED) { ll = tid % LPB; j = tid / LPB; }
  else { ll = tid / TN; j = tid % TN; }
  Cpx<T>* __restrict__ Sl = S + ll * PITCH;

  auto sync = [&]() { if (WARP_SYNC) __syncwarp(); else __syncthreads(); };

  // ---- line addressing -----------------------------------------------------------------------
  LineIO<T, LAY, SRC, LPB> io;
  {
    long long l = tile * LPB + ll;
    io.valid = true; io.valid1 = true;
    if (GEN && l >= a.lines) { l = a.lines - 1; io.valid = false; }
    if (LAY == LAY_REALPAIR) {
      const long long row1 = 2 * l + 1;
      io.valid1 = io.valid;
      const T* in = reinterpret_cast<const T*>(a.in);
      T* out = reinterpret_cast<T*>(a.out);
      io.r0i = in + (size_t)(2 * l) * a.n_in;
      io.r0o = out + (size_t)(2 * l) * a.n_out;
      if (GEN && row1 >= a.real_lines) { io.valid1 = false; io.r1i = io.r0i; io.r1o = io.r0o; }
      else { io.r1i = io.r0i + a.n_in; io.r1o = io.r0o + a.n_out; }
    } else if (LAY == LAY_STRIDED) {
      const unsigned inner = (unsigned)a.inner;
      const unsigned o = (unsigned)l / inner, b = (unsigned)l - o * inner;
      io.is = a.inner;
      io.cin = reinterpret_cast<const Cpx<T>*>(a.in) + (size_t)o * a.n_in * inner + b;
      io.cout = reinterpret_cast<Cpx<T>*>(a.out) + (size_t)o * a.n_out * inner + b;
    } else {
      io.cin = reinterpret_cast<const Cpx<T>*>(a.in) + (size_t)l * a.n_in;
      io.cout = reinterpret_cast<Cpx<T>*>(a.out) + (size_t)l * a.n_out;
    }
  }
  if (SRC == SRC_STAGED) {
    const Cpx<T>* sg = reinterpret_cast<const Cpx<T>*>(stage);
    if (LAY == LAY_REALPAIR) {
      io.sr0 = reinterpret_cast<const T*>(stage) + (size_t)(2 * ll) * N;
      io.sr1 = io.sr0 + N;
    } else if (LAY == LAY_STRIDED) {
      io.st = sg + ll;                  // [n][LPB]
    } else {
      io.st = sg + (size_t)ll * N;      // [LPB][n]
    }
  }
  const Cpx<T>* __restrict__ tw = reinterpret_cast<const Cpx<T>*>(a.tw);
  const Cpx<T>* __restrict__ half = reinterpret_cast<const Cpx<T>*>(a.half);
  const Cpx<T>* __restrict__ pre = reinterpret_cast<const Cpx<T>*>(a.pre);
  const int n_in = a.n_in;

  Cpx<T> v[E];

  // ================================ pass 0: global -> registers ================================
  {
    constexpr int STR = N / R0, BPT = E / R0;
#pragma unroll
    for (int bf = 0; bf < BPT; ++bf) {
      const int jj = j + bf * TN;
#pragma unroll
      for (int r = 0; r < R0; ++r) {
        const int m = jj + r * STR;                     // FFT input index
        Cpx<T> z;
        if (KIND == K_FOUR_FWD) {
          z = io.load(m);
        } else if (KIND == K_FOUR_BWD) {
          // mid-spectrum zero padding (Fourier.py:139-147): index m of the padded spectrum <- coefficient p
          int p = m;
          bool ok = true;
          if (GEN && n_in != N) {
            const int hlf = n_in >> 1;
            if (m < hlf) p = m;
            else if (m >= N - (n_in - hlf)) p = m - (N - n_in);
            else { p = 0; ok = false; }
          }
          z = io.load(p);
          if (pre) z = cmul(z, pre[p]);
          if (GEN && !ok) z = Cpx<T>{T(0), T(0)};
          z.y = -z.y;                                    // inverse DFT = conj(FFT(conj(.)))
        } else if (KIND == K_CHEB_FWD) {
          // DCT-II input permutation: v[m] = x[2m] (m < n/2), x[2n-1-2m] otherwise
          const int src = (r < R0 / 2) ? 2 * m : 2 * N - 1 - 2 * m;
          z = io.load(src);
        } else {  // K_CHEB_BWD
          // z_m = e^{+i pi m/(2n)}/2 (A_m - i A_{n-m}), z_0 = A_0, A_k = c_k (-1)^k (zero beyond n_in)
          Cpx<T> cm{T(0), T(0)}, cp{T(0), T(0)};
          const int mp = N - m;
          if (!GEN || m < n_in) cm = io.load(m);
          const bool has_p = !(r == 0 && bf == 0) || j != 0;   // m != 0
          if (has_p && (!GEN || mp < n_in)) cp = io.load((r == 0 && bf == 0) ? (j != 0 ? mp : 0) : mp);
          const Cpx<T> t = half[m];                      // e^{-i pi m/(2n)}
          const Cpx<T> w{cm.x + cp.y, cm.y - cp.x};      // c_m - i c_{n-m}
          T hs = (jj & 1) ? T(-0.5) : T(0.5);            // (-1)^m / 2 (n even: m and n-m share parity)
          if (r == 0 && bf == 0 && j == 0) hs = T(1);
          z.x = hs * (t.x * w.x + t.y * w.y);            // conj(t) * w
          z.y = -hs * (t.x * w.y - t.y * w.x);           // ... conjugated for the inverse DFT
        }
        v[bf * R0 + r] = z;
      }
    }
  }
  after_load();
  fft_core<T, N, WARP_SYNC>(v, Sl, j, tw);
  {
    constexpr int R = RL, NS = NSL, BPT = E / R;
    // thread now holds FFT bins b = jj + r * NS

    if (KIND == K_FOUR_BWD) {
#pragma unroll
      for (int bf = 0; bf < BPT; ++bf)
#pragma unroll
        for (int r = 0; r < R; ++r) {
          Cpx<T> z = v[bf * R + r];
          z.y = -z.y;
          io.template store<GEN>(j + bf * TN + r * NS, z);
        }
    } else if (KIND == K_FOUR_FWD) {
      const T scale = (T)a.scale;
      const int nm = a.n_modes;
#pragma unroll
      for (int bf = 0; bf < BPT; ++bf)
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int b = j + bf * TN + r * NS;
          Cpx<T> z = v[bf * R + r];
          z.x *= scale; z.y *= scale;
          if (GEN && nm != N) {
            // wavenumber gather (Fourier.py:177-179): keep bins [0, ceil(nm/2)) and [N - floor(nm/2), N)
            if (b < ((nm + 1) >> 1)) io.template store<GEN>(b, z);
            else if (b >= N - (nm >> 1)) io.template store<GEN>(b - (N - nm), z);
          } else {
            io.template store<GEN>(b, z);
          }
        }
    } else if (KIND == K_CHEB_BWD) {
      if (LAY == LAY_STRIDED) {
        // u[2b] = V[b] (b < n/2), u[2(n-1-b)+1] = V[b] otherwise; rows of a strided axis: no cost
#pragma unroll
        for (int bf = 0; bf < BPT; ++bf)
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const int b = j + bf * TN + r * NS;
            Cpx<T> z = v[bf * R + r];
            z.y = -z.y;
            io.template store<GEN>((r < R / 2) ? 2 * b : 2 * N - 1 - 2 * b, z);
          }
      } else {
        // contiguous axis: permute through shared memory so that global stores stay coalesced
        sync();                                          // all reads of the previous exchange done
#pragma unroll
        for (int bf = 0; bf < BPT; ++bf)
#pragma unroll
          for (int r = 0; r < R; ++r) {
            Cpx<T> z = v[bf * R + r];
            z.y = -z.y;
            Sl[sk<LOGSK>(j + bf * TN + r * NS)] = z;
          }
        sync();
        const bool odd = j & 1;                          // TN even: position parity = parity of j
#pragma unroll
        for (int e = 0; e < E; ++e) {
          const int pos = j + e * TN;
          const int b = odd ? (N - 1 - (pos >> 1)) : (pos >> 1);
          io.template store<GEN>(pos, Sl[sk<LOGSK>(b)]);
        }
      }
    } else {  // K_CHEB_FWD: C_b = t_b W[b] + conj(t_b) W[n-b], t_b = e^{-i pi b/(2n)}
      sync();
#pragma unroll
      for (int bf = 0; bf < BPT; ++bf)
#pragma unroll
        for (int r = 0; r < R; ++r) Sl[sk<LOGSK>(j + bf * TN + r * NS)] = v[bf * R + r];
      sync();
      const T scale = (T)a.scale;
      const int nout = a.n_out;
#pragma unroll
      for (int bf = 0; bf < BPT; ++bf) {
        const int jj = j + bf * TN;
        const T s = (jj & 1) ? -scale : scale;           // NS even: parity of b = parity of jj
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int b = jj + r * NS;
          if (GEN && b >= nout) continue;
          const Cpx<T> wk = v[bf * R + r];
          const Cpx<T> wm = Sl[sk<LOGSK>(b == 0 ? 0 : N - b)];
          const Cpx<T> t = half[b];
          Cpx<T> c;
          c.x = t.x * (wk.x + wm.x) - t.y * (wk.y - wm.y);   // t wk + conj(t) wm
          c.y = t.x * (wk.y + wm.y) + t.y * (wk.x - wm.x);
          T sc = s;
          if (r == 0 && bf == 0 && j == 0 && a.kind == FAST_CHEB_FORWARD) sc *= T(0.5);
          c.x *= sc; c.y *= sc;
          io.template store<GEN>(b, c);
        }
      }
    }
  }
  (void)CHEB;
}

template <typename T, int N, int KIND, int LAY, bool GEN>
__global__ void __launch_bounds__(Cta<N, LAY>::THREADS, Cta<N, LAY>::MINB)
fft2_kernel(const FftArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  fft2_tile<T, N, KIND, LAY, GEN>(a, a.reverse ? gridDim.x - 1 - blockIdx.x : blockIdx.x, reinterpret_cast<Cpx<T>*>(smem_raw));
}

}  // namespace f2
}  // namespace jfx
