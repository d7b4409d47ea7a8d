// Shared pieces of the fast-transform kernels (kernels_fft.cu: first-generation smem-staged kernel,
// kernels_fft2.cu: direct global<->register kernel).  Internal header.
#pragma once
#include <cuda_runtime.h>

#include "jfx_common.h"

namespace jfx {

template <typename T> struct alignas(2 * sizeof(T)) Cpx { T x, y; };
template <typename T> __device__ __forceinline__ Cpx<T> operator+(Cpx<T> a, Cpx<T> b) { return {a.x + b.x, a.y + b.y}; }
template <typename T> __device__ __forceinline__ Cpx<T> operator-(Cpx<T> a, Cpx<T> b) { return {a.x - b.x, a.y - b.y}; }
template <typename T> __device__ __forceinline__ Cpx<T> cmul(Cpx<T> a, Cpx<T> b) {
  return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
template <typename T> __device__ __forceinline__ Cpx<T> conj_(Cpx<T> a) { return {a.x, -a.y}; }

// cos(2 pi k/16), sin(2 pi k/16), k = 0..7
__device__ constexpr double kCos16[8] = {1.0, 0.92387953251128673848, 0.70710678118654752440, 0.38268343236508977173,
                                         0.0, -0.38268343236508977173, -0.70710678118654752440, -0.92387953251128673848};
__device__ constexpr double kSin16[8] = {0.0, 0.38268343236508977173, 0.70710678118654752440, 0.92387953251128673848,
                                         1.0, 0.92387953251128673848, 0.70710678118654752440, 0.38268343236508977173};

// cos(2 pi k/12), sin(2 pi k/12), k = 0..5 (radices 6 and 12 of the 3 * 2^m lengths 48, 96, 192)
__device__ constexpr double kCos12[6] = {1.0, 0.86602540378443864676, 0.5, 0.0, -0.5, -0.86602540378443864676};
__device__ constexpr double kSin12[6] = {0.0, 0.5, 0.86602540378443864676, 1.0, 0.86602540378443864676, 0.5};

// cos(2 pi k/20), sin(2 pi k/20), k = 0..9 (radices 10 and 20 of the 5 * 2^m lengths 80, 160, 320)
__device__ constexpr double kCos20[10] = {1.0, 0.9510565162951535, 0.8090169943749475, 0.5877852522924731, 0.30901699437494745, 0.0, -0.30901699437494734, -0.587785252292473, -0.8090169943749473, -0.9510565162951535};
__device__ constexpr double kSin20[10] = {0.0, 0.3090169943749474, 0.5877852522924731, 0.8090169943749475, 0.9510565162951535, 1.0, 0.9510565162951536, 0.8090169943749475, 0.5877852522924732, 0.3090169943749475};

// In-register forward DFT of R points, natural order in and out (decimation in time).
template <typename T, int R> struct Dft {
  static __device__ __forceinline__ void run(Cpx<T>* v) {
    Cpx<T> e[R / 2], o[R / 2];
#pragma unroll
    for (int i = 0; i < R / 2; ++i) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
    Dft<T, R / 2>::run(e);
    Dft<T, R / 2>::run(o);
#pragma unroll
    for (int k = 0; k < R / 2; ++k) {
      Cpx<T> t;
      if (k == 0) t = o[k];
      else if (4 * k == R) t = Cpx<T>{o[k].y, -o[k].x};   // * (-i)
      else {
        const T c = (R % 5 == 0) ? (T)kCos20[k * (20 / R)] : (R % 3 == 0) ? (T)kCos12[k * (12 / R)] : (T)kCos16[k * (16 / R)];   // W = c - i s
        const T s = (R % 5 == 0) ? (T)kSin20[k * (20 / R)] : (R % 3 == 0) ? (T)kSin12[k * (12 / R)] : (T)kSin16[k * (16 / R)];
        t = Cpx<T>{o[k].x * c + o[k].y * s, o[k].y * c - o[k].x * s};
      }
      v[k] = e[k] + t;
      v[k + R / 2] = e[k] - t;
    }
  }
};
template <typename T> struct Dft<T, 2> {
  static __device__ __forceinline__ void run(Cpx<T>* v) {
    Cpx<T> a = v[0] + v[1], b = v[0] - v[1];
    v[0] = a; v[1] = b;
  }
};
template <typename T> struct Dft<T, 1> { static __device__ __forceinline__ void run(Cpx<T>*) {} };
template <typename T> struct Dft<T, 3> {
  static __device__ __forceinline__ void run(Cpx<T>* v) {
    // X1 = v0 - (v1+v2)/2 - i (sqrt3/2)(v1 - v2),  X2 = conj-mirror
    const Cpx<T> s = v[1] + v[2], d = v[1] - v[2];
    const Cpx<T> t{v[0].x - T(0.5) * s.x, v[0].y - T(0.5) * s.y};
    const T h = (T)0.86602540378443864676;
    const Cpx<T> u{h * d.y, -h * d.x};     // -i * h * d
    v[0] = v[0] + s;
    v[1] = t + u;
    v[2] = t - u;
  }
};

template <typename T> struct Dft<T, 5> {
  static __device__ __forceinline__ void run(Cpx<T>* v) {
    // Winograd-style radix 5: X_k = v0 + sum_m v_m W5^{mk}, with the (1,4) and (2,3) pairs combined
    const T c1 = (T)0.30901699437494742410, c2 = (T)-0.80901699437494742410;   // cos(2 pi/5), cos(4 pi/5)
    const T s1 = (T)0.95105651629515357212, s2 = (T)0.58778525229247312917;    // sin(2 pi/5), sin(4 pi/5)
    const Cpx<T> a1 = v[1] + v[4], b1 = v[1] - v[4], a2 = v[2] + v[3], b2 = v[2] - v[3];
    const Cpx<T> t1{v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y};
    const Cpx<T> t2{v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y};
    // -i (s1 b1 + s2 b2)  and  -i (s2 b1 - s1 b2)
    const Cpx<T> u1{s1 * b1.y + s2 * b2.y, -(s1 * b1.x + s2 * b2.x)};
    const Cpx<T> u2{s2 * b1.y - s1 * b2.y, -(s2 * b1.x - s1 * b2.x)};
    v[0] = v[0] + a1 + a2;
    v[1] = t1 + u1;
    v[4] = t1 - u1;
    v[2] = t2 + u2;
    v[3] = t2 - u2;
  }
};

// radix plans
template <int N> struct Plan;
template <> struct Plan<16>   { static constexpr int R0 = 4,  R1 = 4,  R2 = 1;  };
template <> struct Plan<32>   { static constexpr int R0 = 4,  R1 = 8,  R2 = 1;  };
template <> struct Plan<64>   { static constexpr int R0 = 8,  R1 = 8,  R2 = 1;  };
template <> struct Plan<128>  { static constexpr int R0 = 8,  R1 = 16, R2 = 1;  };
#ifdef JFX_PLAN256_884
template <> struct Plan<256>  { static constexpr int R0 = 8,  R1 = 8,  R2 = 4;  };   // 8 points per thread: half the registers, twice the warps
#else
template <> struct Plan<256>  { static constexpr int R0 = 16, R1 = 16, R2 = 1;  };
#endif
template <> struct Plan<512>  { static constexpr int R0 = 8,  R1 = 8,  R2 = 8;  };
template <> struct Plan<1024> { static constexpr int R0 = 16, R1 = 8,  R2 = 8;  };
template <> struct Plan<2048> { static constexpr int R0 = 16, R1 = 16, R2 = 8;  };
template <> struct Plan<4096> { static constexpr int R0 = 16, R1 = 16, R2 = 16; };
// 3 * 2^m: the first radix stays a power of two (it defines the shared-memory skew), the factor 3 sits in the
// last pass; a thread owns 12 points so that every pass holds a whole number of butterflies
template <> struct Plan<48>   { static constexpr int R0 = 4,  R1 = 12, R2 = 1;  };
template <> struct Plan<96>   { static constexpr int R0 = 4,  R1 = 4,  R2 = 6;  };
template <> struct Plan<192>  { static constexpr int R0 = 4,  R1 = 4,  R2 = 12; };
template <> struct Plan<384>  { static constexpr int R0 = 4,  R1 = 8,  R2 = 12; };   // 3/2 x 256: a thread owns 24 points
// 5 * 2^m: the factor 5 sits in the last pass (radix 20 / 10); a thread owns 20 points
template <> struct Plan<80>   { static constexpr int R0 = 4,  R1 = 20, R2 = 1;  };
template <> struct Plan<160>  { static constexpr int R0 = 4,  R1 = 4,  R2 = 10; };
template <> struct Plan<320>  { static constexpr int R0 = 4,  R1 = 4,  R2 = 20; };

constexpr int cmax(int a, int b) { return a > b ? a : b; }
constexpr int ilog2(int v) { return v <= 1 ? 0 : 1 + ilog2(v / 2); }
template <int N> struct PointsPerThread { static constexpr int value = cmax(Plan<N>::R0, cmax(Plan<N>::R1, Plan<N>::R2)); };
template <> struct PointsPerThread<48> { static constexpr int value = 12; };
template <> struct PointsPerThread<96> { static constexpr int value = 12; };
template <> struct PointsPerThread<192> { static constexpr int value = 12; };
template <> struct PointsPerThread<384> { static constexpr int value = 24; };
template <> struct PointsPerThread<80> { static constexpr int value = 20; };
template <> struct PointsPerThread<160> { static constexpr int value = 20; };
template <> struct PointsPerThread<320> { static constexpr int value = 20; };
template <int N> struct Geo {
  static constexpr int RMAX = PointsPerThread<N>::value;   // points a thread owns (= largest radix for 2^m)
  static constexpr int TN = N / RMAX;                       // threads per line
  static constexpr int LOGSK = ilog2(Plan<N>::R0);          // smem skew: i + (i >> LOGSK)
  static constexpr int PITCH = N + (N >> LOGSK) + 1;        // odd -> conflict-free across lines
};

// Layout of the twiddle table behind the N base entries (host: fast_tables_create, device: fft_core)
//   middle pass of a three-pass plan (radix R1, butterfly index k < R0, twiddle stride TS = N / (R0 R1)):
//       tw[OFF_MID  + (r - 1) * R0      + k] = W_N^(r k TS)      r = 1 .. R1 - 1
//   last pass (radix RL, k < N / RL):
//       tw[OFF_LAST + (r - 1) * (N/RL)  + k] = W_N^(r k)         r = 1 .. RL - 1
template <int N> struct TwLayout {
  using P = Plan<N>;
  static constexpr bool THREE = P::R2 > 1;
  static constexpr int RL = THREE ? P::R2 : P::R1;
  static constexpr int OFF_MID = N;
  static constexpr int MID = THREE ? (P::R1 - 1) * P::R0 : 0;
  static constexpr int OFF_LAST = N + MID;
  static constexpr int LAST = (RL - 1) * (N / RL);
  static constexpr int TOTAL = N + MID + LAST;
};

template <int LOGSK> __device__ __forceinline__ int sk(int i) { return i + (i >> LOGSK); }


struct FftArgs {
  const void* in;
  void* out;
  const void* tw;     // W_n^m = exp(-2 pi i m / n), m < n
  const void* half;   // exp(-i pi k / (2n)), k < n          (Chebyshev)
  const void* pre;    // per-coefficient complex prescale    (Fourier backward derivative), or null
  long long lines;    // complex lines (real-pair mode: pairs of real lines)
  long long inner;    // complex inner extent; 1 = contiguous axis
  long long real_lines;
  int n_in, n_out;    // axis extents of the input / output arrays
  int n_modes;        // N
  int kind;
  int real_pair;
  int lpb;
  int log_lpb;
  int reverse;        // tile order last-to-first (L2 reuse between consecutive passes)
  double scale;
};


struct FastTables {
  int n = 0;
  bool dbl = true;
  void* d_tw = nullptr;
  void* d_half = nullptr;
  void* d_pre = nullptr;
  ~FastTables() {
    if (d_tw) cudaFree(d_tw);
    if (d_half) cudaFree(d_half);
    if (d_pre) cudaFree(d_pre);
  }
};


enum { LAY_CONTIG = 0, LAY_STRIDED = 1, LAY_REALPAIR = 2 };
enum { K_CHEB_BWD = 0, K_CHEB_FWD = 1, K_FOUR_BWD = 2, K_FOUR_FWD = 3 };

// second-generation kernel (kernels_fft2.cu): returns 1 when it handled the launch, 0 when the
// configuration is outside its envelope (caller falls back to the staged kernel), < 0 on error
int launch_fast_axis_v2(cudaStream_t s, const FftArgs& a, int n, bool dbl);
// streaming variant (kernels_fft2_stream.cu): same return convention
int launch_fast_axis_stream(cudaStream_t s, const FftArgs& a, int n, bool dbl, int mode);   // mode 1: bulk copies, 2: cp.async

int make_fft_args(const AxisGeom& g, int dtype, const FastParams& p, const FastTables* t, const void* in, void* out,
                  FftArgs* out_args, bool* empty);

// plane-fused two-axis pass (kernels_fft2_pair.cu)
size_t fast_pair_counter_bytes(long long planes);
int launch_fast_pair(cudaStream_t s, FftArgs a, FftArgs b, int n, bool dbl, long long planes, void* ring_buf,
                     size_t ring_bytes, void* counters, bool query);

}  // namespace jfx
