// Plane-fused two-axis pass: the transforms along the LAST TWO axes of an array [P, n1, n2] in ONE
// persistent launch whose intermediate stays in L2.
//
// For a fixed leading index s the plane [n1, n2] is self-contained for both axes, so the work is cut into
//   A-tiles of plane s : axis n1 (strided),    LPB neighbouring lines each
//   B-tiles of plane s : axis n2 (contiguous), LPB lines (or pairs of real lines) each
// and handed out in ORDER through an atomic ticket: block k of the ticket sequence holds the A-tiles of
// plane k followed by the B-tiles of plane k - LAG.  A B-tile spins on a per-plane counter until all
// A-tiles of its plane have been written (they were ticketed ~LAG planes earlier, so the wait is normally
// zero); because tickets are in order and A-tiles never wait for data, the scheme cannot deadlock whatever
// the grid size.  The intermediate goes through a RING of `ring` plane slots (an A-tile that reuses a slot
// waits until the B-tiles of the previous occupant are done), so it is rewritten while still dirty in L2
// and never reaches HBM: the pair costs one read and one write of the field instead of two of each.
#include <cuda_runtime.h>

#include "fft2_tile.cuh"

namespace jfx {
namespace f2 {

struct PairArgs {
  FftArgs a;              // pass A: in = source array, out = ring (plane-relative addressing: see below)
  FftArgs b;              // pass B: in = ring, out = destination array
  unsigned* ticket;       // [1]
  unsigned* produced;     // [planes]  A-tiles finished per plane
  unsigned* consumed;     // [planes]  B-tiles finished per plane
  int planes, ta, tb;     // tiles per plane of each pass
  int lag, ring;
  long long slot_elems_a_out;   // elements (of pass-A output type units, complex) per plane in the ring
};

__device__ __forceinline__ void spin_until(const unsigned* p, unsigned want) {
  while (*reinterpret_cast<const volatile unsigned*>(p) < want) __nanosleep(64);
}

template <typename T, int N, int KIND, int LAYB>
__global__ void __launch_bounds__(Cta<N, LAY_STRIDED>::THREADS, Cta<N, LAY_STRIDED>::MINB)
fft2_pair_kernel(const __grid_constant__ PairArgs pa) {
  static_assert(Cta<N, LAY_STRIDED>::THREADS == Cta<N, LAYB>::THREADS, "both passes use the same CTA size");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Cpx<T>* S = reinterpret_cast<Cpx<T>*>(smem_raw);
  __shared__ unsigned s_item;
  const unsigned blk_a = pa.ta, blk_b = pa.tb;
  const unsigned P = pa.planes, D = pa.lag;
  const unsigned total = P * (blk_a + blk_b);
  // ticket -> (block k, offset): blocks [0, D) hold A only, [D, P) hold A then B, [P, P + D) hold B only
  const unsigned head = D * blk_a, mid = (P - D) * (blk_a + blk_b);
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_item = atomicAdd(pa.ticket, 1u);
    __syncthreads();
    const unsigned item = s_item;
    if (item >= total) break;
    bool is_a;
    unsigned plane, t;
    if (item < head) { is_a = true; plane = item / blk_a; t = item % blk_a; }
    else if (item < head + mid) {
      const unsigned r = item - head, k = r / (blk_a + blk_b), o = r % (blk_a + blk_b);
      if (o < blk_a) { is_a = true; plane = D + k; t = o; }
      else { is_a = false; plane = k; t = o - blk_a; }
    } else { const unsigned r = item - head - mid; is_a = false; plane = (P - D) + r / blk_b; t = r % blk_b; }
    const unsigned slot = plane % pa.ring;
    if (is_a) {
      if (plane >= (unsigned)pa.ring) {   // the slot's previous occupant must have been consumed
        if (threadIdx.x == 0) spin_until(pa.consumed + (plane - pa.ring), blk_b);
        __syncthreads();
      }
      // pass A reads plane `plane` of the source and writes slot `slot` of the ring: both arrays are
      // addressed by a global tile index, so shift the output base by (slot - plane) planes
      FftArgs a = pa.a;
      a.out = reinterpret_cast<Cpx<T>*>(a.out) + ((long long)slot - (long long)plane) * pa.slot_elems_a_out;
      fft2_tile<T, N, KIND, LAY_STRIDED, false>(a, (long long)plane * blk_a + t, S);
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) atomicAdd(pa.produced + plane, 1u);
    } else {
      if (threadIdx.x == 0) { spin_until(pa.produced + plane, blk_a); __threadfence(); }
      __syncthreads();
      FftArgs b = pa.b;
      b.in = reinterpret_cast<const Cpx<T>*>(b.in) + ((long long)slot - (long long)plane) * pa.slot_elems_a_out;
      fft2_tile<T, N, KIND, LAYB, false, SRC_GLOBAL_CG>(b, (long long)plane * blk_b + t, S);
      __syncthreads();
      if (threadIdx.x == 0) atomicAdd(pa.consumed + plane, 1u);
    }
  }
}

template <typename T, int N, int KIND, int LAYB>
static int launch_pair_variant(cudaStream_t s, const PairArgs& pa) {
  constexpr int THREADS = Cta<N, LAY_STRIDED>::THREADS;
  constexpr int LPB_A = THREADS / (N / Geo<N>::RMAX);
  const size_t smem = (size_t)LPB_A * Geo<N>::PITCH * sizeof(Cpx<T>);
  static int grid = 0;
  if (!grid) {
    JFX_CUDA_OK(cudaFuncSetAttribute(fft2_pair_kernel<T, N, KIND, LAYB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 0, nb = 0;
    JFX_CUDA_OK(cudaGetDevice(&dev));
    JFX_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    JFX_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fft2_pair_kernel<T, N, KIND, LAYB>, THREADS, smem));
    grid = sms * (nb < 1 ? 1 : nb);
  }
  const long long total = (long long)pa.planes * (pa.ta + pa.tb);
  const int g = (int)(total < grid ? total : grid);
  fft2_pair_kernel<T, N, KIND, LAYB><<<g, THREADS, smem, s>>>(pa);
  JFX_CUDA_OK(cudaGetLastError());
  return 1;
}

template <typename T, int N>
static int launch_pair_n(cudaStream_t s, const PairArgs& pa, int k4, int layb) {
#define JFX_CASE(K, L) if (k4 == K && layb == L) return launch_pair_variant<T, N, K, L>(s, pa);
  JFX_CASE(K_CHEB_BWD, LAY_REALPAIR) JFX_CASE(K_CHEB_FWD, LAY_REALPAIR)
  JFX_CASE(K_CHEB_BWD, LAY_CONTIG) JFX_CASE(K_CHEB_FWD, LAY_CONTIG)
  JFX_CASE(K_FOUR_BWD, LAY_CONTIG) JFX_CASE(K_FOUR_FWD, LAY_CONTIG)
#undef JFX_CASE
  return 0;
}

}  // namespace f2

size_t fast_pair_counter_bytes(long long planes) { return (size_t)(2 * planes + 64) * sizeof(unsigned); }

// 1 = launched, 0 = outside the envelope, < 0 = error.  `a` / `b` are the two passes' arguments over the
// WHOLE arrays (a.out and b.in are filled in here: the ring), `counters` holds fast_pair_counter_bytes().
int launch_fast_pair(cudaStream_t s, FftArgs a, FftArgs b, int n, bool dbl, long long planes, void* ring_buf,
                     size_t ring_bytes, void* counters, bool query) {
  using namespace f2;
  if (!dbl) return 0;
  if (a.kind != b.kind && !((a.kind == FAST_CHEB_FORWARD || a.kind == FAST_CHEB_SCALAR) &&
                            (b.kind == FAST_CHEB_FORWARD || b.kind == FAST_CHEB_SCALAR))) return 0;
  if (a.real_pair || a.inner <= 1 || b.inner != 1) return 0;   // A strided, B contiguous
  if (a.n_in != n || a.n_out != n || b.n_in != n || b.n_out != n) return 0;
  if (n != 128 && n != 256) return 0;   // sizes whose strided and contiguous tiles use the same CTA size
  if (planes < 8 || planes >= (1ll << 20)) return 0;
  int k4;
  switch (a.kind) {
    case FAST_CHEB_BACKWARD: k4 = K_CHEB_BWD; break;
    case FAST_CHEB_FORWARD: case FAST_CHEB_SCALAR: k4 = K_CHEB_FWD; break;
    case FAST_FOURIER_BACKWARD: k4 = K_FOUR_BWD; break;
    default: k4 = K_FOUR_FWD;
  }
  if (a.pre || b.pre) return 0;
  const int layb = b.real_pair ? LAY_REALPAIR : LAY_CONTIG;
  // tiles per plane
  int lpb;
  if (n == 128) lpb = Cta<128, LAY_STRIDED>::THREADS / (128 / Geo<128>::RMAX);
  else lpb = Cta<256, LAY_STRIDED>::THREADS / (256 / Geo<256>::RMAX);
  if (a.lines % planes || b.lines % planes) return 0;
  const long long la = a.lines / planes, lb = b.lines / planes;
  if (la % lpb || lb % lpb) return 0;
  if (b.real_pair && (b.real_lines & 1)) return 0;
  PairArgs pa{};
  pa.planes = (int)planes; pa.ta = (int)(la / lpb); pa.tb = (int)(lb / lpb);
  // plane size of the intermediate, in complex elements of T
  pa.slot_elems_a_out = la * n;
  const size_t slot_bytes = (size_t)pa.slot_elems_a_out * 16;
  long long ring = (long long)(ring_bytes / slot_bytes);
  if (ring > planes) ring = planes;
  long long lag = 24;
  if (lag > planes / 2) lag = planes / 2;
  if (ring < 2 * lag + 8 && ring < planes) return 0;
  if (query) return 1;
  pa.lag = (int)lag; pa.ring = (int)ring;
  unsigned* c = reinterpret_cast<unsigned*>(counters);
  JFX_CUDA_OK(cudaMemsetAsync(c, 0, fast_pair_counter_bytes(planes), s));
  pa.ticket = c; pa.produced = c + 64; pa.consumed = c + 64 + planes;
  a.out = ring_buf; b.in = ring_buf;
  pa.a = a; pa.b = b;
  if (n == 128) return launch_pair_n<double, 128>(s, pa, k4, layb);
  return launch_pair_n<double, 256>(s, pa, k4, layb);
}

}  // namespace jfx
