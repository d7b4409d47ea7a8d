// Row-fused nonlinear-term kernel: for every line of the LAST (Fourier) axis, in ONE CTA and without
// touching HBM in between,
//
//     leaf_l = ifft( pad( mult_l(k) * A_{g(l)}[row, k] ) )          l = 0 .. n_leaves-1
//     E      = program(leaf_0, ..., statics)                         pointwise (integrators/nonlinear.py:135-217)
//     out    = truncate( scale * fft(E) )
//
// i.e. the last-axis part of  testspace.forward(evaluator(uh))  (integrators/base.py:230-248): every
// leaf's Fourier.backward / backward_primitive (galerkin/Fourier.py:126-148, 206-219, orthogonal.py:229-246)
// followed by the pointwise tree and Fourier.forward / scalar_product (Fourier.py:150-180).  A_g are the
// coefficient rows after the other axes (if any) were taken to physical space; leaves that share the
// derivative orders on those axes share one A_g.  For a 1-D (batched) field the whole nonlinear term is
// this single kernel: one read of uh and one write of the result.
//
// The inverse / forward transforms reuse the register + shared-memory Stockham core of kernels_fft2.cu;
// leaf lines are parked in a small CTA-private scratch that stays resident in L2 (a persistent grid owns
// gridDim * LPB * n_leaves lines in all: ~20 MB) between the inverse transforms, the pointwise
// evaluation and the forward transform; shared memory only holds the FFT exchange buffer, so the kernel
// keeps the occupancy of the plain FFT kernel.
#include <cuda_runtime.h>

#include <cstdlib>

#include "fft_common.cuh"
#include "fft_core.cuh"
#include "pointwise.cuh"

namespace jfx {

#ifndef JFX_FUSED_TPS
#define JFX_FUSED_TPS 512   /* resident threads per SM the register budget is sized for */
#endif
#ifndef JFX_FUSED_THREADS
#define JFX_FUSED_THREADS 64   /* one 1024-point line per CTA: barriers span two warps (round 2: KdV 1.38 -> 1.33 ms) */
#endif
constexpr int FUSED_THREADS = JFX_FUSED_THREADS;

// MODE: where the leaf lines live between their inverse transform and the pointwise evaluation
//   PARK_ALL_L2  every leaf and the pointwise result go through the CTA's L2-resident scratch (general programs: the stack
//                machine indexes its operands at run time)
//   PARK_L2 / PARK_SMEM  polynomial programs: the LAST leaf never leaves the registers, the pointwise result is formed in
//                registers and handed to the forward transform by a compile-time register permutation (the bins a thread
//                holds after an inverse transform are exactly the pass-0 inputs it needs next); only leaves 0 .. L-2 are
//                parked — in shared memory when they fit beside the exchange buffer, else in the L2 scratch.  KdV -u u_x:
//                one parked line per row instead of six scratch transits.
enum { PARK_ALL_L2 = 0, PARK_L2 = 1, PARK_SMEM = 2 };

// PARK_SMEM variants are limited to three CTAs per SM by shared memory: their register budget is sized for 384 threads
template <typename T, int N, bool PAD, int DEPTH, int MODE>
__global__ void __launch_bounds__((N / Geo<N>::RMAX) > FUSED_THREADS ? (N / Geo<N>::RMAX) : FUSED_THREADS,
                                  (MODE == PARK_SMEM && (N / Geo<N>::RMAX) <= FUSED_THREADS ? 384 : JFX_FUSED_TPS) /
                                      ((N / Geo<N>::RMAX) > FUSED_THREADS ? (N / Geo<N>::RMAX) : FUSED_THREADS))
fused_rows_kernel(const __grid_constant__ FusedRowArgs a) {
  using P = Plan<N>;
  constexpr int R0 = P::R0;
  constexpr int E = Geo<N>::RMAX, TN = N / E, PITCH = Geo<N>::PITCH;
  constexpr int RL = (P::R2 > 1) ? P::R2 : P::R1, NSL = N / RL;
  constexpr int STR0 = N / R0, BPT0 = E / R0, BPTL = E / RL;
  constexpr int THREADS = TN > FUSED_THREADS ? TN : FUSED_THREADS;
  constexpr int LPB = THREADS / TN;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  Cpx<T>* __restrict__ S = reinterpret_cast<Cpx<T>*>(smem_raw);          // [LPB][PITCH] exchange buffer

  const int tid = threadIdx.x;
  const int ll = tid / TN, j = tid % TN;
  Cpx<T>* __restrict__ Sl = S + ll * PITCH;
  const Cpx<T>* __restrict__ tw = reinterpret_cast<const Cpx<T>*>(a.tw);
  const int nc = a.n_coeff, hlf = nc >> 1;
  // Leaf lines: a CTA-private slice of the scratch (L2-resident: gridDim * LPB * n_leaves lines in all).
  // A thread only ever reads back the points it wrote itself (the bins a thread holds after the last FFT
  // pass, j + k*TN, are exactly the points whose pass-0 inputs it needs next), so no barrier guards them.
  Cpx<T>* __restrict__ Lb = (MODE == PARK_SMEM)
      ? S + (size_t)LPB * PITCH + (size_t)ll * (size_t)(a.n_leaves - 1) * N          // parked leaves behind the exchange buffer
      : reinterpret_cast<Cpx<T>*>(a.scratch) + ((size_t)blockIdx.x * LPB + ll) * (size_t)a.n_leaves * N;
  const long long ntiles = (a.rows + LPB - 1) / LPB;
  const int last_leaf = a.n_leaves - 1;

  Cpx<T> v[E];
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    long long row = tile * LPB + ll;
    const bool valid = row < a.rows;
    if (!valid) row = a.rows - 1;
    // ---- leaves: inverse transforms ----------------------------------------------------------------
    for (int l = 0; l < a.n_leaves; ++l) {
      const Cpx<T>* __restrict__ in = reinterpret_cast<const Cpx<T>*>(a.src[a.leaf_group[l]]) + (size_t)row * nc;
      const Cpx<T>* __restrict__ mult = reinterpret_cast<const Cpx<T>*>(a.mult[l]);
#pragma unroll
      for (int bf = 0; bf < BPT0; ++bf) {
#pragma unroll
        for (int r = 0; r < R0; ++r) {
          const int m = j + bf * TN + r * STR0;
          // mid-spectrum zero padding (Fourier.py:139-147)
          int p = m;
          bool ok = true;
          if (PAD) {
            if (m < hlf) p = m;
            else if (m >= N - (nc - hlf)) p = m - (N - nc);
            else { p = 0; ok = false; }
          }
          Cpx<T> z = in[p];
          if (mult) z = cmul(z, mult[p]);
          if (PAD && !ok) z = Cpx<T>{T(0), T(0)};
          z.y = -z.y;
          v[bf * R0 + r] = z;
        }
      }
      fft_core<T, N, false>(v, Sl, j, tw);
      if (MODE == PARK_ALL_L2 || l < last_leaf) {
        Cpx<T>* __restrict__ dst = Lb + (size_t)l * N;
#pragma unroll
        for (int bf = 0; bf < BPTL; ++bf)
#pragma unroll
          for (int r = 0; r < RL; ++r) {
            Cpx<T> z = v[bf * RL + r];
            z.y = -z.y;
            dst[j + bf * TN + r * NSL] = z;
          }
      }
      __syncthreads();                                   // exchange buffer is reused by the next transform
    }
    if (MODE != PARK_ALL_L2) {
      // ---- polynomial program on registers: last leaf = v, the others from the parked lines (own points only) -------
      static_assert(E % 4 == 0, "points per thread");
#pragma unroll
      for (int e = 0; e < E; ++e) v[e].y = -v[e].y;      // conjugate: inverse DFT = conj(FFT(conj(.)))
#pragma unroll
      for (int e0 = 0; e0 < E; e0 += 4) {
        auto leaf4 = [&](int l, int c) -> C2<T> {
          const int e = e0 + c, m = j + (e / RL) * TN + (e % RL) * NSL;   // the bin register e holds
          if (l == last_leaf) return C2<T>{v[e].x, v[e].y};
          const Cpx<T> z = Lb[(size_t)l * N + m];
          return C2<T>{z.x, z.y};
        };
        C2<T> r4[4];
        poly_eval_vec<T, 4>(a.poly, leaf4, r4);
#pragma unroll
        for (int c = 0; c < 4; ++c) v[e0 + c] = Cpx<T>{r4[c].re, r4[c].im};
      }
      // ---- register permutation: bin order of the last pass -> input order of pass 0 ----------------------------------
      {
        Cpx<T> w[E];
#pragma unroll
        for (int bf = 0; bf < BPT0; ++bf)
#pragma unroll
          for (int r = 0; r < R0; ++r) {
            const int k = bf + r * BPT0;                 // point j + k * TN
            w[bf * R0 + r] = v[(k % BPTL) * RL + k / BPTL];
          }
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = w[e];
      }
    } else
    // ---- pointwise program over the thread's own points, in place over the line of leaf 0 ------------
    if (a.poly.n_terms > 0) {
      // polynomial normal form, four points at a time (their scratch loads overlap)
      static_assert(E % 4 == 0, "points per thread");
#pragma unroll 1
      for (int e0 = 0; e0 < E; e0 += 4) {
        auto leaf4 = [&](int l, int c) -> C2<T> {
          const Cpx<T> z = Lb[(size_t)l * N + j + (e0 + c) * TN];
          return C2<T>{z.x, z.y};
        };
        C2<T> r4[4];
        poly_eval_vec<T, 4>(a.poly, leaf4, r4);
#pragma unroll
        for (int c = 0; c < 4; ++c) Lb[j + (e0 + c) * TN] = Cpx<T>{r4[c].re, r4[c].im};
      }
    } else {
#pragma unroll 1
      for (int e = 0; e < E; ++e) {
        const int m = j + e * TN;
        auto leaf = [&](int l) -> C2<T> {
          const Cpx<T> z = Lb[(size_t)l * N + m];
          return C2<T>{z.x, z.y};
        };
        auto stat = [&](int s) -> C2<T> {
          const Cpx<T> z = reinterpret_cast<const Cpx<T>*>(a.statics[s])[(size_t)row * N + m];
          return C2<T>{z.x, z.y};
        };
        const C2<T> r = pw_eval<T, true, DEPTH>(a.instr, a.n_instr, a.consts, leaf, stat);
        Lb[m] = Cpx<T>{r.re, r.im};
      }
    }
    if (MODE == PARK_ALL_L2) {
#pragma unroll
      for (int bf = 0; bf < BPT0; ++bf)
#pragma unroll
        for (int r = 0; r < R0; ++r) v[bf * R0 + r] = Lb[j + bf * TN + r * STR0];
    }
    // ---- forward transform, scale, wavenumber gather (Fourier.py:150-180) ---------------------------
    fft_core<T, N, false>(v, Sl, j, tw);
    {
      const T scale = (T)a.scale;
      const int nm = a.n_out;
      Cpx<T>* __restrict__ out = reinterpret_cast<Cpx<T>*>(a.out) + (size_t)row * nm;
      if (valid) {
#pragma unroll
        for (int bf = 0; bf < BPTL; ++bf)
#pragma unroll
          for (int r = 0; r < RL; ++r) {
            const int b = j + bf * TN + r * NSL;
            Cpx<T> z = v[bf * RL + r];
            z.x *= scale; z.y *= scale;
            if (PAD) {
              if (b < ((nm + 1) >> 1)) out[b] = z;
              else if (b >= N - (nm >> 1)) out[b - (N - nm)] = z;
            } else {
              out[b] = z;
            }
          }
      }
    }
    __syncthreads();
  }
}

// Geometry of one launch: persistent grid (CTAs resident at once), scratch bytes it needs.
// shared memory of one CTA: the exchange buffer, plus (PARK_SMEM) the parked lines of leaves 0 .. L-2
template <typename T, int N>
static size_t fused_smem(int n_leaves, int mode) {
  constexpr int E = Geo<N>::RMAX, TN = N / E;
  constexpr int THREADS = TN > FUSED_THREADS ? TN : FUSED_THREADS;
  constexpr int LPB = THREADS / TN;
  size_t smem = (size_t)LPB * Geo<N>::PITCH * sizeof(Cpx<T>);
  if (mode == PARK_SMEM) smem += (size_t)LPB * (size_t)(n_leaves - 1) * N * sizeof(Cpx<T>);
  return smem;
}

template <typename T, int N, bool PAD, int DEPTH, int MODE>
static int fused_geometry(const FusedRowArgs& a, int* grid_out, size_t* smem_out, size_t* scratch_out) {
  constexpr int E = Geo<N>::RMAX, TN = N / E;
  constexpr int THREADS = TN > FUSED_THREADS ? TN : FUSED_THREADS;
  constexpr int LPB = THREADS / TN;
  const size_t smem = fused_smem<T, N>(a.n_leaves, MODE);
  // occupancy depends on the leaf count in PARK_SMEM mode: cached per leaf count
  static int per_sm_l[JFX_MAX_LEAVES + 1], sms = 0;
  static bool init = false;
  if (!init) { for (int& v : per_sm_l) v = -1; init = true; }
  int& per_sm = per_sm_l[a.n_leaves];
  if (per_sm < 0) {
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
      JFX_CUDA_OK(cudaFuncSetAttribute(fused_rows_kernel<T, N, PAD, DEPTH, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_smem = smem;
    }
    int dev = 0;
    JFX_CUDA_OK(cudaGetDevice(&dev));
    JFX_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int nb = 0;
    JFX_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fused_rows_kernel<T, N, PAD, DEPTH, MODE>, THREADS, smem));
    per_sm = nb < 1 ? 1 : nb;
  }
  const long long ntiles = (a.rows + LPB - 1) / LPB;
  long long grid = (long long)per_sm * sms;
  if (grid > ntiles) grid = ntiles;
  if (grid < 1) grid = 1;
  *grid_out = (int)grid;
  *smem_out = smem;
  // scratch is sized for the full persistent grid so that it does not depend on the row count
  *scratch_out = (size_t)per_sm * sms * LPB * (size_t)a.n_leaves * N * sizeof(Cpx<T>);
  return JFX_OK;
}

template <typename T, int N, bool PAD, int DEPTH, int MODE>
static int launch_fused_v(cudaStream_t s, const FusedRowArgs& a, size_t* scratch_query) {
  constexpr int E = Geo<N>::RMAX, TN = N / E;
  constexpr int THREADS = TN > FUSED_THREADS ? TN : FUSED_THREADS;
  int grid;
  size_t smem, scratch;
  int rc = fused_geometry<T, N, PAD, DEPTH, MODE>(a, &grid, &smem, &scratch);
  if (rc != JFX_OK) return rc;
  if (scratch_query) { *scratch_query = scratch; return 1; }
  fused_rows_kernel<T, N, PAD, DEPTH, MODE><<<grid, THREADS, smem, s>>>(a);
  JFX_CUDA_OK(cudaGetLastError());
  return 1;
}

template <typename T, int N>
static int launch_fused_n(cudaStream_t s, const FusedRowArgs& a, size_t* q) {
  // PAD: zero padding on the way in or truncation on the way out; DEPTH: operand stack of the program
  const bool pad = (a.n_coeff != N) || (a.n_out != N);
  const bool deep = a.depth > 4;
  // polynomial programs keep the last leaf and the result in registers (JFX_NL_REG=0: the round-1 scratch route, for A/B runs)
  static const bool reg_off = [] { const char* e = getenv("JFX_NL_REG"); return e && e[0] == '0'; }();
  if (a.poly.n_terms > 0 && !reg_off) {
    // parked leaves: the L2-resident scratch by default — measured on the B200 (KdV 65 536 x 1024: 0.92 ms against 1.06 ms with
    // shared-memory parking, Cahn-Hilliard 1024^2 0.103 against 0.125 ms): the parked lines cost a third of the resident
    // warps (12 instead of 16 per SM), which hurts more than the L2 round trip of ONE line per row.  JFX_NL_PARK=smem selects
    // shared-memory parking where it fits.
    static const int park_env = [] { const char* e = getenv("JFX_NL_PARK"); return !e ? 0 : (e[0] == 's' ? 2 : 1); }();
    const bool in_smem = park_env == 2 && a.n_leaves >= 1 && fused_smem<T, N>(a.n_leaves, PARK_SMEM) <= 200 * 1024;
    if (in_smem) return pad ? launch_fused_v<T, N, true, 4, PARK_SMEM>(s, a, q) : launch_fused_v<T, N, false, 4, PARK_SMEM>(s, a, q);
    return pad ? launch_fused_v<T, N, true, 4, PARK_L2>(s, a, q) : launch_fused_v<T, N, false, 4, PARK_L2>(s, a, q);
  }
  if (pad) return deep ? launch_fused_v<T, N, true, 8, PARK_ALL_L2>(s, a, q) : launch_fused_v<T, N, true, 4, PARK_ALL_L2>(s, a, q);
  return deep ? launch_fused_v<T, N, false, 8, PARK_ALL_L2>(s, a, q) : launch_fused_v<T, N, false, 4, PARK_ALL_L2>(s, a, q);
}

template <typename T>
static int launch_fused_t(cudaStream_t s, int n, const FusedRowArgs& a, size_t* q) {
  switch (n) {
    case 48: return launch_fused_n<T, 48>(s, a, q);
    case 64: return launch_fused_n<T, 64>(s, a, q);
    case 96: return launch_fused_n<T, 96>(s, a, q);
    case 192: return launch_fused_n<T, 192>(s, a, q);
    case 128: return launch_fused_n<T, 128>(s, a, q);
    case 256: return launch_fused_n<T, 256>(s, a, q);
    case 512: return launch_fused_n<T, 512>(s, a, q);
    case 1024: return launch_fused_n<T, 1024>(s, a, q);
    case 2048: return launch_fused_n<T, 2048>(s, a, q);
    case 4096: return launch_fused_n<T, 4096>(s, a, q);
  }
  return 0;
}

// 1 = launched (or, with scratch_query != null, launchable: *scratch_query = scratch bytes the launch needs),
// 0 = outside the envelope, < 0 = error
int launch_fused_rows(cudaStream_t s, int dtype, int n, const FusedRowArgs& a, size_t* scratch_query) {
  if (!dtype_is_complex(dtype)) return 0;
  if (!scratch_query && a.rows <= 0) return 1;
  return dtype == JFX_C128 ? launch_fused_t<double>(s, n, a, scratch_query) : launch_fused_t<float>(s, n, a, scratch_query);
}

}  // namespace jfx
