// Fast transform kernels: Fourier c2c FFT and Chebyshev DCT-II / DCT-III along one tensor axis.
//
// Reference semantics (all formulas restated from the Python reference):
//   Fourier.backward        galerkin/Fourier.py:126-148   mid-spectrum zero pad + ifft(norm="forward")
//   Fourier.forward         galerkin/Fourier.py:165-180   fft(norm="forward") + wavenumber gather
//   Fourier.scalar_product  galerkin/Fourier.py:150-163   ... * 2 pi / domain_factor
//   Fourier.derivative      galerkin/Fourier.py:206-219   (i m)^k prescale (host table `pre`)
//   Chebyshev.backward      galerkin/Chebyshev.py:225-241 0.5*uh[0] + n*idct(uh), uh = c (-1)^k
//   Chebyshev.forward       galerkin/Chebyshev.py:243-260 dct(u), uh[0]/2, *(-1)^k/n, truncate
//   Chebyshev.scalar_product galerkin/Chebyshev.py:262-279 dct(u) * pi (-1)^k / (2 n df), truncate
//
// Design: one CTA owns a tile of LPB lines of length n (n = power of two, 16..4096) in shared
// memory.  Global traffic is one coalesced read and one coalesced write of the tile (16-byte
// elements; for a strided axis the tile is [n, LPB] with the LPB neighbouring lines contiguous in
// memory).  The FFT itself is a Stockham autosort with radix-4/8/16 register butterflies, twiddles
// from a host-computed fp64 table, 2-3 passes through padded (bank-conflict-skewed) shared memory.
// DCTs use the n-point complex FFT of TWO real sequences at once: the real and imaginary parts of a
// complex line, or two neighbouring real lines ("real pair" mode) — so real data costs half a
// complex transform per line.  The cosine transforms' half-sample twiddles, the (-1)^k signs, the
// padding / truncation / wavenumber gather and all scalings are fused into the tile load / store.
#include <cuda_runtime.h>
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "jfx_common.h"
#include "fft_common.cuh"

namespace jfx {

template <int TN> __device__ __forceinline__ void line_sync();

// one Stockham pass over a line held in shared memory (in place: load all, sync, store all)
template <typename T, int N, int R, int NS>
__device__ __forceinline__ void fft_pass(Cpx<T>* __restrict__ S, int j, const Cpx<T>* __restrict__ tw) {
  constexpr int TN = Geo<N>::TN, LOGSK = Geo<N>::LOGSK, G = 1 << LOGSK;
  constexpr int BPT = (N / R) / TN;
  constexpr int STR = N / R;                          // input stride between a butterfly's points
  static_assert(STR % G == 0, "input stride must be a multiple of the skew granule");
  static_assert(NS == 1 ? (R <= G) : (NS % G == 0), "output stride must keep the skew affine");
  constexpr int IN_OFF = STR + (STR >> LOGSK);        // sk(jj + r*STR) = sk(jj) + r*IN_OFF
  constexpr int OUT_OFF = NS == 1 ? 1 : NS + (NS >> LOGSK);  // sk(j0 + r*NS) = sk(j0) + r*OUT_OFF
  constexpr int TS = N / (NS * R);                    // twiddle table stride
  Cpx<T> v[BPT][R];
#pragma unroll
  for (int bf = 0; bf < BPT; ++bf) {
    const Cpx<T>* src = S + sk<LOGSK>(j + bf * TN);
#pragma unroll
    for (int r = 0; r < R; ++r) v[bf][r] = src[r * IN_OFF];
  }
  line_sync<TN>();
#pragma unroll
  for (int bf = 0; bf < BPT; ++bf) {
    const int jj = j + bf * TN;
    const int k = jj % NS;
    if (NS > 1) {
      const Cpx<T>* w = tw + k * TS;
#pragma unroll
      for (int r = 1; r < R; ++r) v[bf][r] = cmul(v[bf][r], w[(r - 1) * k * TS]);   // tw[r*k*TS]
    }
    Dft<T, R>::run(v[bf]);
    Cpx<T>* dst = S + sk<LOGSK>((jj / NS) * NS * R + k);
#pragma unroll
    for (int r = 0; r < R; ++r) dst[r * OUT_OFF] = v[bf][r];
  }
  line_sync<TN>();
}

template <int TN> __device__ __forceinline__ void line_sync() {
  // a line's TN threads live in one warp when TN <= 32: warp-level sync is enough
  if (TN <= 32) __syncwarp(); else __syncthreads();
}


// Per-thread view of the tile: which (local line, axis index) a thread touches in stage iteration
// `it`, and where that line starts in the input / output arrays.  Everything that does not depend
// on `it` is computed once.
template <typename T, int N, int LAY> struct TileMap {
  static constexpr int TN = Geo<N>::TN;
  static constexpr bool WARP_LINES = (TN <= 32);
  static constexpr int LW = WARP_LINES ? 32 / TN : 1;
  int tid, lane, warp, nthr, lpb;
  int ll_fixed, i_first;      // strided
  long long lines, line0, real_lines;
  int n_in, n_out;
  long long inner;
  size_t sb_in, sb_out;       // strided: element offset of (line, i = 0)

  __device__ __forceinline__ TileMap(const FftArgs& a)
      : tid(threadIdx.x), lane(threadIdx.x & 31), warp(threadIdx.x >> 5), nthr(blockDim.x), lpb(a.lpb),
        lines(a.lines), line0((long long)blockIdx.x * a.lpb), real_lines(a.real_lines), n_in(a.n_in),
        n_out(a.n_out), inner(a.inner) {
    ll_fixed = 0; i_first = 0; sb_in = sb_out = 0;
    if (LAY == LAY_STRIDED) {
      ll_fixed = tid & (lpb - 1);
      i_first = tid >> a.log_lpb;
      long long l = line0 + ll_fixed;
      if (l >= lines) l = lines - 1;   // clamp: masked later
      const long long o = l / inner, b = l - o * inner;
      sb_in = (size_t)(o * n_in * inner + b);
      sb_out = (size_t)(o * n_out * inner + b);
    }
  }
  // (local line, axis index) of stage iteration `it` (compile-time after unrolling)
  __device__ __forceinline__ void elem(int it, int& ll, int& i) const {
    if (LAY == LAY_STRIDED) { ll = ll_fixed; i = i_first + it * TN; }
    else if (WARP_LINES) {
      if (N >= 32) { ll = warp * LW + (it * 32) / N; i = ((it * 32) % N) + lane; }
      else { const int e = it * 32 + lane; ll = warp * LW + e / N; i = e % N; }
    } else { const int e = tid + it * nthr; ll = e / N; i = e % N; }
  }
  __device__ __forceinline__ bool line_ok(int ll) const { return line0 + ll < lines; }
  __device__ __forceinline__ Cpx<T> load(const void* in, int ll, int i) const {
    if (LAY == LAY_REALPAIR) {
      const long long l = line0 + ll;
      const T* __restrict__ p = reinterpret_cast<const T*>(in) + (size_t)(2 * l) * n_in + i;
      Cpx<T> v;
      v.x = p[0];
      v.y = (2 * l + 1 < real_lines) ? p[n_in] : T(0);
      return v;
    }
    const Cpx<T>* __restrict__ p = reinterpret_cast<const Cpx<T>*>(in);
    if (LAY == LAY_STRIDED) return p[sb_in + (size_t)i * inner];
    return p[(size_t)(line0 + ll) * n_in + i];
  }
  __device__ __forceinline__ void store(void* out, int ll, int i, Cpx<T> v) const {
    if (LAY == LAY_REALPAIR) {
      const long long l = line0 + ll;
      T* p = reinterpret_cast<T*>(out) + (size_t)(2 * l) * n_out + i;
      p[0] = v.x;
      if (2 * l + 1 < real_lines) p[n_out] = v.y;
      return;
    }
    Cpx<T>* p = reinterpret_cast<Cpx<T>*>(out);
    if (LAY == LAY_STRIDED) p[sb_out + (size_t)i * inner] = v;
    else p[(size_t)(line0 + ll) * n_out + i] = v;
  }
};

template <typename T, int N, int KIND, int LAY>
__device__ __forceinline__ void stage_in(const FftArgs& a, const TileMap<T, N, LAY>& tm, Cpx<T>* __restrict__ S) {
  constexpr int PITCH = Geo<N>::PITCH, LOGSK = Geo<N>::LOGSK, RMAX = Geo<N>::RMAX;
  constexpr int CH = (RMAX < 4) ? RMAX : 4;   // independent global requests in flight per thread
  const Cpx<T>* __restrict__ half = reinterpret_cast<const Cpx<T>*>(a.half);
  const Cpx<T>* __restrict__ pre = reinterpret_cast<const Cpx<T>*>(a.pre);
  const int n_in = a.n_in;
#pragma unroll
  for (int it0 = 0; it0 < RMAX; it0 += CH) {
    Cpx<T> z0[CH], z1[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      int ll, i;
      tm.elem(it0 + c, ll, i);
      z0[c] = Cpx<T>{T(0), T(0)};
      z1[c] = Cpx<T>{T(0), T(0)};
      if (!tm.line_ok(ll)) continue;
      if (KIND == K_FOUR_BWD) {
        const int hlf = n_in >> 1;                 // padded spectrum position i <- coefficient p
        int p = i;
        if (n_in != N) p = (i < hlf) ? i : ((i >= N - (n_in - hlf)) ? i - (N - n_in) : -1);
        if (p >= 0) { z0[c] = tm.load(a.in, ll, p); if (pre) z1[c] = pre[p]; }
      } else if (KIND == K_CHEB_BWD) {
        if (i < n_in) z0[c] = tm.load(a.in, ll, i);
        const int m = N - i;
        if (i > 0 && m < n_in) z1[c] = tm.load(a.in, ll, m);
      } else {
        z0[c] = tm.load(a.in, ll, i);               // forward kinds: n_in == N
      }
    }
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      int ll, i;
      tm.elem(it0 + c, ll, i);
      Cpx<T> z = z0[c];
      int dst = i;
      if (KIND == K_FOUR_BWD) {
        if (pre) z = cmul(z, z1[c]);
        z = conj_(z);                                // inverse DFT = conj(FFT(conj(.)))
      } else if (KIND == K_CHEB_BWD) {
        // z_0 = A_0 ; z_k = e^{+i pi k/(2n)}/2 (A_k - i A_{n-k}),  A_k = c_k (-1)^k (0 beyond n_in)
        Cpx<T> ak = z, am = z1[c];
        if (i & 1) { ak.x = -ak.x; ak.y = -ak.y; }   // N even: (N - i) has the parity of i
        if (i & 1) { am.x = -am.x; am.y = -am.y; }
        if (i != 0) {
          const Cpx<T> w{ak.x + am.y, ak.y - am.x};   // A_k - i A_{n-k}
          const Cpx<T> t = half[i];                   // e^{-i pi k/(2n)}; we need its conjugate
          z = Cpx<T>{T(0.5) * (t.x * w.x + t.y * w.y), T(0.5) * (t.x * w.y - t.y * w.x)};
        } else z = ak;
        z = conj_(z);
      } else if (KIND == K_CHEB_FWD) {
        dst = (i & 1) ? (N - 1 - (i >> 1)) : (i >> 1);  // DCT-II: v[m] = x[2m], v[n-1-m] = x[2m+1]
      }
      S[ll * PITCH + sk<LOGSK>(dst)] = z;
    }
  }
}

template <typename T, int N, int KIND, int LAY>
__device__ __forceinline__ void stage_out(const FftArgs& a, const TileMap<T, N, LAY>& tm, const Cpx<T>* __restrict__ S) {
  constexpr int PITCH = Geo<N>::PITCH, LOGSK = Geo<N>::LOGSK, RMAX = Geo<N>::RMAX;
  const Cpx<T>* __restrict__ half = reinterpret_cast<const Cpx<T>*>(a.half);
  const T scale = (T)a.scale;
  const int nout = a.n_out;
#pragma unroll
  for (int it = 0; it < RMAX; ++it) {
    int ll, q;
    tm.elem(it, ll, q);
    if (!tm.line_ok(ll) || q >= nout) continue;
    const Cpx<T>* Sl = S + ll * PITCH;
    Cpx<T> v;
    if (KIND == K_FOUR_BWD) {
      v = conj_(Sl[sk<LOGSK>(q)]);
    } else if (KIND == K_CHEB_BWD) {
      const int m = (q & 1) ? (N - 1 - (q >> 1)) : (q >> 1);
      v = conj_(Sl[sk<LOGSK>(m)]);
    } else if (KIND == K_CHEB_FWD) {
      // C_k = t_k W[k] + conj(t_k) W[n-k],  t_k = e^{-i pi k/(2n)}
      const Cpx<T> t = half[q];
      const Cpx<T> wk = Sl[sk<LOGSK>(q)], wm = Sl[sk<LOGSK>((N - q) & (N - 1))];
      v = cmul(t, wk) + cmul(conj_(t), wm);
      T s = (q & 1) ? -scale : scale;
      if (q == 0 && a.kind == FAST_CHEB_FORWARD) s *= T(0.5);
      v.x *= s; v.y *= s;
    } else {
      const int nm = a.n_modes;                   // Fourier forward: wavenumber gather when truncating
      int i = q;
      if (N > nm && q >= ((nm + 1) >> 1)) i = N + q - nm;
      v = Sl[sk<LOGSK>(i)];
      v.x *= scale; v.y *= scale;
    }
    tm.store(a.out, ll, q, v);
  }
}

template <typename T, int N, int KIND, int LAY>
__device__ __forceinline__ void fft_tile(const FftArgs& a, Cpx<T>* S) {
  constexpr int TN = Geo<N>::TN, PITCH = Geo<N>::PITCH;
  constexpr bool WARP_LOCAL = (TN <= 32) && (LAY != LAY_STRIDED);  // stage loops touch only the warp's own lines
  const TileMap<T, N, LAY> tm(a);
  stage_in<T, N, KIND, LAY>(a, tm, S);
  if (WARP_LOCAL) __syncwarp(); else __syncthreads();
  {
    const int ll = threadIdx.x / TN, j = threadIdx.x % TN;
    Cpx<T>* Sl = S + ll * PITCH;
    const Cpx<T>* __restrict__ tw = reinterpret_cast<const Cpx<T>*>(a.tw);
    using P = Plan<N>;
    fft_pass<T, N, P::R0, 1>(Sl, j, tw);
    fft_pass<T, N, P::R1, P::R0>(Sl, j, tw);
    if constexpr (P::R2 > 1) fft_pass<T, N, P::R2, P::R0 * P::R1>(Sl, j, tw);
  }
  if (!WARP_LOCAL) __syncthreads();
  stage_out<T, N, KIND, LAY>(a, tm, S);
}

template <typename T, int N>
__global__ void __launch_bounds__(Geo<N>::TN > 128 ? 512 : 256)
fft_axis_kernel(const FftArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Cpx<T>* S = reinterpret_cast<Cpx<T>*>(smem_raw);
  // uniform dispatch: one specialised (kind, layout) path per launch, no per-element branching
  const int lay = a.real_pair ? LAY_REALPAIR : (a.inner > 1 ? LAY_STRIDED : LAY_CONTIG);
  int k4;
  switch (a.kind) {
    case FAST_CHEB_BACKWARD: k4 = K_CHEB_BWD; break;
    case FAST_CHEB_FORWARD: case FAST_CHEB_SCALAR: k4 = K_CHEB_FWD; break;
    case FAST_FOURIER_BACKWARD: k4 = K_FOUR_BWD; break;
    default: k4 = K_FOUR_FWD;
  }
#define JFX_CASE(K, L) if (k4 == K && lay == L) { fft_tile<T, N, K, L>(a, S); return; }
  JFX_CASE(K_CHEB_BWD, LAY_CONTIG) JFX_CASE(K_CHEB_BWD, LAY_STRIDED) JFX_CASE(K_CHEB_BWD, LAY_REALPAIR)
  JFX_CASE(K_CHEB_FWD, LAY_CONTIG) JFX_CASE(K_CHEB_FWD, LAY_STRIDED) JFX_CASE(K_CHEB_FWD, LAY_REALPAIR)
  JFX_CASE(K_FOUR_BWD, LAY_CONTIG) JFX_CASE(K_FOUR_BWD, LAY_STRIDED)
  JFX_CASE(K_FOUR_FWD, LAY_CONTIG) JFX_CASE(K_FOUR_FWD, LAY_STRIDED)
#undef JFX_CASE
}

// ---------------------------------------------------------------------------------------------------
static bool supported_n(int n) {
  switch (n) {
    case 16: case 32: case 64: case 128: case 256: case 512: case 1024: case 2048: case 4096: return true;
    case 48: case 96: case 192: case 384: return true;   // 3 * 2^m: second-generation kernel only
    case 80: case 160: case 320: return true;  // 5 * 2^m: second-generation kernel only
  }
  return false;
}

bool fast_geometry_ok(const AxisGeom& g, int dtype) {
  // real data: last axis (pairs of lines) or an even inner extent (neighbouring lines = re/im)
  if (dtype_is_complex(dtype)) return true;
  return g.inner == 1 || g.inner % 2 == 0;
}

bool fast_available(int basis, int n, int dtype) {
  if (basis != JFX_BASIS_CHEBYSHEV && basis != JFX_BASIS_FOURIER) return false;
  if (basis == JFX_BASIS_FOURIER && !dtype_is_complex(dtype)) return false;
  return supported_n(n);
}

template <typename T>
static int upload_complex(const std::vector<long double>& re, const std::vector<long double>& im, void** dptr) {
  std::vector<T> h(2 * re.size());
  for (size_t i = 0; i < re.size(); ++i) { h[2 * i] = (T)re[i]; h[2 * i + 1] = (T)im[i]; }
  JFX_CUDA_OK(cudaMalloc(dptr, h.size() * sizeof(T)));
  JFX_CUDA_OK(cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return JFX_OK;
}

// exp(-i pi num/den) with exact octant reduction, long double
static void expmipi(long long num, long long den, long double* c, long double* s) {
  num %= 2 * den;
  if (num < 0) num += 2 * den;
  const long double PI = 3.141592653589793238462643383279502884L;
  // reduce to [0, den/2] using symmetries about pi/2 and pi
  long long r = num;
  int sgn_c = 1, sgn_s = 1;
  if (r >= den) { r -= den; sgn_c = -sgn_c; sgn_s = -sgn_s; }   // angle - pi
  if (2 * r > den) { r = den - r; sgn_c = -sgn_c; }               // pi - angle
  long double ang = PI * (long double)r / (long double)den;
  long double cc, ss;
  if (4 * r <= den) { cc = cosl(ang); ss = sinl(ang); }
  else { long double a2 = PI * (long double)(den - 2 * r) / (long double)(2 * den); cc = sinl(a2); ss = cosl(a2); }
  *c = sgn_c * cc;
  *s = -sgn_s * ss;  // exp(-i theta) = cos - i sin
}

template <int N> static void radices_of(int* r0, int* r1, int* r2) {
  *r0 = Plan<N>::R0; *r1 = Plan<N>::R1; *r2 = Plan<N>::R2;
}
static bool plan_radices(int n, int* r0, int* r1, int* r2) {
  switch (n) {
#define JFX_R(N) case N: radices_of<N>(r0, r1, r2); return true;
    JFX_R(16) JFX_R(32) JFX_R(64) JFX_R(128) JFX_R(256) JFX_R(512) JFX_R(1024) JFX_R(2048) JFX_R(4096)
    JFX_R(48) JFX_R(96) JFX_R(192) JFX_R(384) JFX_R(80) JFX_R(160) JFX_R(320)
#undef JFX_R
  }
  return false;
}

int fast_tables_create(const FastParams& p, int dtype, FastTables** out) {
  *out = nullptr;
  const int n = p.n_quad;
  JFX_REQUIRE(supported_n(n), JFX_ERR_UNSUPPORTED, "no fast transform kernel for n=%d", n);
  const bool cheb = p.kind <= FAST_CHEB_SCALAR;
  JFX_REQUIRE(!(cheb && p.deriv != 0), JFX_ERR_UNSUPPORTED,
              "Chebyshev derivative orders are folded into dense tables by the host");
  FastTables* t = new FastTables;
  t->n = n;
  t->dbl = dtype_is_double(dtype);
  std::vector<long double> re(n), im(n);
  for (int m = 0; m < n; ++m) expmipi(2LL * m, n, &re[m], &im[m]);
  {
    // per-pass tables behind the base entries (TwLayout<N>, fft_common.cuh): the same values, stored by butterfly index
    int R0 = 0, R1 = 0, R2 = 0;
    JFX_REQUIRE(plan_radices(n, &R0, &R1, &R2), JFX_ERR_UNSUPPORTED, "no radix plan for n=%d", n);
    const bool three = R2 > 1;
    const int RL = three ? R2 : R1;
    std::vector<long double> wr(re), wi(im);
    if (three) {
      const int TS = n / (R0 * R1);
      for (int r = 1; r < R1; ++r)
        for (int k = 0; k < R0; ++k) { wr.push_back(re[(size_t)r * k * TS]); wi.push_back(im[(size_t)r * k * TS]); }
    }
    for (int r = 1; r < RL; ++r)
      for (int k = 0; k < n / RL; ++k) { wr.push_back(re[(size_t)r * k]); wi.push_back(im[(size_t)r * k]); }
    const int rc0 = t->dbl ? upload_complex<double>(wr, wi, &t->d_tw) : upload_complex<float>(wr, wi, &t->d_tw);
    if (rc0 != JFX_OK) { delete t; return rc0; }
  }
  int rc = JFX_OK;
  if (rc == JFX_OK && cheb) {
    for (int k = 0; k < n; ++k) expmipi(k, 2LL * n, &re[k], &im[k]);
    rc = t->dbl ? upload_complex<double>(re, im, &t->d_half) : upload_complex<float>(re, im, &t->d_half);
  }
  if (rc == JFX_OK && p.kind == FAST_FOURIER_BACKWARD && p.deriv > 0) {
    // (i m)^k * df^k, Nyquist mode removed for odd k   (Fourier.py:206-219, orthogonal.py:245)
    const int nc = p.n_modes;
    std::vector<long double> pr(nc), pi_(nc);
    const long double dfk = powl((long double)p.domain_factor, p.deriv);
    for (int q = 0; q < nc; ++q) {
      long long m = (q < (nc + 1) / 2) ? q : q - nc;
      if ((p.deriv & 1) && (nc % 2 == 0) && q == nc / 2) m = 0;
      long double mag = powl((long double)m, p.deriv) * dfk;   // m^k (sign kept for odd k)
      switch (p.deriv & 3) {  // i^k
        case 0: pr[q] = mag; pi_[q] = 0; break;
        case 1: pr[q] = 0; pi_[q] = mag; break;
        case 2: pr[q] = -mag; pi_[q] = 0; break;
        default: pr[q] = 0; pi_[q] = -mag; break;
      }
    }
    rc = t->dbl ? upload_complex<double>(pr, pi_, &t->d_pre) : upload_complex<float>(pr, pi_, &t->d_pre);
  }
  if (rc != JFX_OK) { delete t; return rc; }
  *out = t;
  return JFX_OK;
}

void fast_tables_destroy(FastTables* t) { delete t; }

template <typename T, int N>
static int launch_n(cudaStream_t s, const FftArgs& a_in, bool strided) {
  FftArgs a = a_in;
  constexpr int TN = Geo<N>::TN;
  const int max_threads = TN > 128 ? 512 : 256;
  const size_t line_bytes = (size_t)Geo<N>::PITCH * sizeof(Cpx<T>);
  int lpb = max_threads / TN;
  if (lpb < 1) lpb = 1;
  const size_t smem_cap = 200 * 1024;
  while (lpb > 1 && (size_t)lpb * line_bytes > smem_cap) lpb /= 2;
  // strided axes want >= 8 neighbouring lines (128 B rows); small problems shrink the tile
  while (lpb > 1 && (long long)(lpb / 2) >= a.lines) lpb /= 2;
  if (TN <= 32 && lpb < 32 / TN) lpb = 32 / TN;   // whole warps: a warp owns 32/TN lines
  a.lpb = lpb;
  a.log_lpb = 0;
  while ((1 << a.log_lpb) < lpb) ++a.log_lpb;
  const size_t smem = (size_t)lpb * line_bytes;
  static bool attr_done = false;
  if (!attr_done) {
    JFX_CUDA_OK(cudaFuncSetAttribute(fft_axis_kernel<T, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cap + 8192));
    attr_done = true;
  }
  const long long blocks = (a.lines + lpb - 1) / lpb;
  JFX_REQUIRE(blocks < (1ll << 31), JFX_ERR_UNSUPPORTED, "too many lines");
  (void)strided;
  fft_axis_kernel<T, N><<<(unsigned)blocks, lpb * TN, smem, s>>>(a);
  JFX_CUDA_OK(cudaGetLastError());
  return JFX_OK;
}

template <typename T>
static int launch_t(cudaStream_t s, const FftArgs& a, int n, bool strided) {
  switch (n) {
    case 16: return launch_n<T, 16>(s, a, strided);
    case 32: return launch_n<T, 32>(s, a, strided);
    case 64: return launch_n<T, 64>(s, a, strided);
    case 128: return launch_n<T, 128>(s, a, strided);
    case 256: return launch_n<T, 256>(s, a, strided);
    case 512: return launch_n<T, 512>(s, a, strided);
    case 1024: return launch_n<T, 1024>(s, a, strided);
    case 2048: return launch_n<T, 2048>(s, a, strided);
    case 4096: return launch_n<T, 4096>(s, a, strided);
  }
  set_error("no fast transform kernel for n=%d", n);
  return JFX_ERR_UNSUPPORTED;
}

// Arguments of one axis pass (shared by the single-pass kernels and the plane-fused pair kernel).
// *empty is set when the pass has no lines.
int make_fft_args(const AxisGeom& g, int dtype, const FastParams& p, const FastTables* t, const void* in, void* out,
                  FftArgs* out_args, bool* empty) {
  JFX_REQUIRE(t != nullptr, JFX_ERR_INVALID, "missing fast tables");
  const bool cplx = dtype_is_complex(dtype);
  const bool cheb = p.kind <= FAST_CHEB_SCALAR;
  FftArgs& a = *out_args;
  a = FftArgs{};
  *empty = false;
  a.in = in; a.out = out;
  a.tw = t->d_tw; a.half = t->d_half; a.pre = t->d_pre;
  a.n_in = g.n_in; a.n_out = g.n_out; a.n_modes = p.n_modes; a.kind = p.kind; a.reverse = p.reverse;
  const double n = (double)p.n_quad;
  const double PI = 3.14159265358979323846;
  switch (p.kind) {
    case FAST_CHEB_FORWARD: a.scale = 1.0 / n; break;
    case FAST_CHEB_SCALAR: a.scale = PI / n / 2.0 / p.domain_factor; break;
    case FAST_FOURIER_FORWARD: a.scale = 1.0 / n; break;
    case FAST_FOURIER_SCALAR: a.scale = (1.0 / n) * 2.0 * PI / p.domain_factor; break;
    default: a.scale = 1.0;
  }
  if (g.outer * g.inner == 0) { *empty = true; return JFX_OK; }
  if (cplx) {
    a.inner = g.inner; a.lines = g.outer * g.inner; a.real_pair = 0;
  } else {
    JFX_REQUIRE(cheb, JFX_ERR_INVALID, "Fourier axes need complex data");
    if (g.inner == 1) {
      a.real_pair = 1; a.real_lines = g.outer; a.lines = (g.outer + 1) / 2; a.inner = 1;
    } else if (g.inner % 2 == 0) {
      a.real_pair = 0; a.inner = g.inner / 2; a.lines = g.outer * a.inner;  // neighbouring real lines = (re, im)
    } else {
      set_error("real Chebyshev axis with odd inner extent %lld has no fast path", (long long)g.inner);
      return JFX_ERR_UNSUPPORTED;
    }
  }
  return JFX_OK;
}

int launch_fast_axis(cudaStream_t s, const AxisGeom& g, int dtype, const FastParams& p, const FastTables* t,
                     const void* in, void* out) {
  FftArgs a;
  bool empty;
  { const int rc = make_fft_args(g, dtype, p, t, in, out, &a, &empty); if (rc != JFX_OK) return rc; }
  if (empty) return JFX_OK;
  // second-generation kernel first (direct global<->register passes); the staged kernel below is the
  // fallback for configurations outside its envelope and, with JFX_FFT_V1=1, an A/B switch for tests
  static const bool force_v1 = [] { const char* e = getenv("JFX_FFT_V1"); return e && e[0] == '1'; }();
  // streaming variant (persistent CTAs, bulk-async prefetch of the next tile) for full, unpadded tiles.
  // Opt-in (JFX_FFT_STREAM=1, read per call): measured 2x slower on strided axes (n bulk copies of 128 B
  // per tile) and within +-7 % of the plain kernel on contiguous axes — see DESIGN.md.
  const char* stream_env = getenv("JFX_FFT_STREAM");
  if (!force_v1 && stream_env && (stream_env[0] == '1' || stream_env[0] == '2')) {
    const int rc = launch_fast_axis_stream(s, a, p.n_quad, dtype_is_double(dtype), stream_env[0] - '0');
    if (rc != 0) return rc < 0 ? rc : JFX_OK;
  }
  if (!force_v1) {
    const int rc = launch_fast_axis_v2(s, a, p.n_quad, dtype_is_double(dtype));
    if (rc != 0) return rc < 0 ? rc : JFX_OK;
  }
  return dtype_is_double(dtype) ? launch_t<double>(s, a, p.n_quad, a.inner > 1)
                                : launch_t<float>(s, a, p.n_quad, a.inner > 1);
}

}  // namespace jfx
