// Parity-folded FP64 tensor-core contraction (half the DMMA work of dgemm_dmma_tma for tables with the mirror
// symmetry of polynomial bases on symmetric nodes).  The math, tile roles and every index formula live in
// dmma_fold.cuh, which tests/emu/fold_emu.cpp executes on the host; this file adds the hardware pipeline:
// a producer warp issuing cp.async.bulk.tensor box copies into a 4-stage ring of 48 KB stages (three 16 KB tiles),
// "full" mbarriers carrying the transaction bytes, "empty" mbarriers with one arrival per MMA warp, a persistent
// grid (one CTA per SM), spin waits bounded by __trap.
//
// Reference semantics: the same contraction as kernels_dense.cu (orthogonal.py:214-277 with the Vandermonde of
// orthogonal.py:131-141); results differ from the unfolded kernel only by summation order.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdlib.h>

#include "jfx_common.h"
#include "dmma_fold.cuh"
#include "dmma_fold_api.h"

namespace jfx {
namespace dmma {
namespace fold {

// 8 MMA warps = 256 threads, up to 255 registers each: no separate producer warp and no register hand-over.  Lane 0 of
// warp 0 issues the TMA copies of k-tile it + AHEAD between its own k-tiles (three box copies per stage — the wide X tensor
// map of dmma_fold.cuh — so the detour costs a few dozen instructions).
// History: round 1 had a producer warpgroup + setmaxnreg (232 registers for the MMA warps); a ninth producer warp without
// setmaxnreg caps every thread at 168 registers (the hardware allocates for 12 warps), which the IN variants do not fit.
constexpr int THREADS = MMA_WARPS * 32;
constexpr int AHEAD = STAGES - 2;      // k-tiles in flight ahead of the compute cursor (the stage refilled at iteration it was
                                       // released by every warp at iteration it - 1: one k-tile of slack for the slowest warp)
#define JFX_FOLD_BOUNDS __launch_bounds__(THREADS, 1)
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 2 * STAGES * 8;
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
    if (++spins > SPIN_LIMIT) __trap();   // a lost signal becomes a launch failure, never a hang
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
__device__ __forceinline__ void tma_5d(unsigned dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, int c4, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_4d(unsigned dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
               ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}

struct MmaOp {
  __device__ __forceinline__ void operator()(double& d0, double& d1, double a, double b) const {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
  }
};

struct GlobalStore {
  double* C;
  __device__ __forceinline__ void s2(long long idx, double v0, double v1) const {
    *reinterpret_cast<double2*>(C + idx) = make_double2(v0, v1);
  }
  __device__ __forceinline__ void s1(long long idx, double v) const { C[idx] = v; }
};

// The whole CTA: pipeline set-up and the eight MMA warps (warp 0 also feeds the pipeline).  epi(tile_m, tile_n, z, wm, wn, g,
// t, acc) stores one warp's share of a finished tile (plain epilogue into this GPU's array, or the scatter epilogue into the
// peers' receive buffers).
template <int V, class EPI>
__device__ __forceinline__ void fold_cta(const CUtensorMap& tmA, const CUtensorMap& tmB, const Args& q, EPI&& epi) {
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte aligned tile ring (the swizzles are functions of the shared-memory address bits 4..9)
  const unsigned base = (s32(smem_raw) + 1023u) & ~1023u;
  unsigned char* gbase = smem_raw + (base - s32(smem_raw));
  const unsigned bar_full = base + STAGES * STAGE_BYTES;
  const unsigned bar_empty = bar_full + STAGES * 8;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, MMA_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int kts = ktiles(q);
  const unsigned total_tiles = (unsigned)q.tiles_n * (unsigned)q.tiles_m * (unsigned)q.batch;   // < 2^31: checked by the host
  const unsigned my_tiles = blockIdx.x < total_tiles ? (total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
  const unsigned total_its = my_tiles * (unsigned)kts;

  // ---- producer state (lane 0 of warp 0): cursor over the k-tiles of this CTA's tiles, AHEAD of the compute cursor ----
  unsigned p_it = 0, p_tl = blockIdx.x;
  int p_kt = 0, p_tm = 0, p_tn = 0, p_z = 0;
  auto p_decode = [&]() {
    const unsigned r = p_tl / (unsigned)q.tiles_n;
    p_tn = (int)(p_tl - r * (unsigned)q.tiles_n);
    p_z = (int)(r / (unsigned)q.tiles_m);
    p_tm = (int)(r - (unsigned)p_z * (unsigned)q.tiles_m) + q.tm_rot;
    if (p_tm >= q.tiles_m) p_tm -= q.tiles_m;
  };
  auto produce = [&]() {      // one k-tile: wait for the stage, announce the bytes, issue the box copies
    if (p_it >= total_its) return;
    const int s = (int)(p_it % STAGES);
    const unsigned ph = (p_it / STAGES) & 1u;
    mbar_wait(bar_empty + 8 * s, ph ^ 1u);
    const unsigned full = bar_full + 8 * s;
    mbar_expect_tx(full, STAGE_BYTES);
    const unsigned dst0 = base + s * STAGE_BYTES;
    stage_copies<V>(q, p_kt, p_tm, p_tn, p_z, [&](int map, int dst, int rank, int c0, int c1, int c2, int c3, int c4) {
      const CUtensorMap* tmap = map == 0 ? &tmA : &tmB;
      const unsigned d = dst0 + (unsigned)dst * 8u;
      if (rank == 2) tma_2d(d, tmap, c0, c1, full);
      else if (rank == 3) tma_3d(d, tmap, c0, c1, c2, full);
      else if (rank == 4) tma_4d(d, tmap, c0, c1, c2, c3, full);
      else tma_5d(d, tmap, c0, c1, c2, c3, c4, full);
    });
    ++p_it;
    if (++p_kt == kts) {
      p_kt = 0;
      p_tl += gridDim.x;
      if (p_it < total_its) p_decode();
    }
  };
  const bool producer = tid == 0;
  if (producer && my_tiles > 0) {
    p_decode();
#pragma unroll 1
    for (int i = 0; i < AHEAD; ++i) produce();
  }

  // ================================ MMA warps ================================
  const int wm = warp / WARPS_N, wn = warp % WARPS_N;
  const int g = lane >> 2, t = lane & 3;
  const Frag fr = frag_init<V>(wm, wn, g, t);
  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  unsigned it = 0;
  int prev = -1;
  for (unsigned tl = blockIdx.x; tl < total_tiles; tl += gridDim.x) {
    for (int kt = 0; kt < kts; ++kt, ++it) {
      const int s = (int)(it % STAGES);
      const unsigned ph = (it / STAGES) & 1u;
      mbar_wait(bar_full + 8 * s, ph);
      // Release of the PREVIOUS stage, here and not at the end of its k-tile.  ptxas schedules an arrive that follows the
      // k-tile directly behind the last fragment loads and ahead of the MMAs that consume them; the arrive then takes
      // effect while those LDS are still queued, and when this warp is the last of the eight the refill (the table tile
      // comes from L2 in a few hundred ns) can overwrite the 1 KB atoms they address.  Seen on the B200 as wrong minus-half
      // accumulators in ~1e-5 of the warp k-tiles whenever the producer was not far ahead, and gone with a slowed-down
      // producer (tools/forensic.py).  After the spin loop of the next wait every MMA of the previous k-tile has issued,
      // i.e. every fragment register has been written: the stage is really free.
      if (prev >= 0) {
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8 * prev);
      }
      prev = s;
      // k-tile it + AHEAD goes into the stage that every warp released one iteration ago
      if (producer) produce();
      __syncwarp();
      const double* S = reinterpret_cast<const double*>(gbase + (size_t)s * STAGE_BYTES);
      ktile<V>(S, fr, ktile_par(q, kt), acc, MmaOp{});
    }
    const unsigned r = tl / (unsigned)q.tiles_n;
    const int tn = (int)(tl - r * (unsigned)q.tiles_n);
    const long long z = r / (unsigned)q.tiles_m;
    int tm = (int)(r - (unsigned)z * (unsigned)q.tiles_m) + q.tm_rot;
    if (tm >= q.tiles_m) tm -= q.tiles_m;
    epi(tm, tn, z, wm, wn, g, t, acc);
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  }
}

template <int V>
__global__ void JFX_FOLD_BOUNDS
dgemm_dmma_fold(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Args q) {
  fold_cta<V>(tmA, tmB, q, [&](int tm, int tn, long long z, int wm, int wn, int g, int t, const double (&acc)[8][4][2]) {
    epilogue<V>(q, tm, tn, z, wm, wn, g, t, acc, GlobalStore{q.C});
  });
}

// Same contraction, result scattered into the slab-exchange receive buffers of all GPUs (NT variants; see Scatter)
template <int V>
__global__ void JFX_FOLD_BOUNDS
dgemm_dmma_fold_scatter(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Args q,
                        const Scatter sc) {
  fold_cta<V>(tmA, tmB, q, [&](int tm, int tn, long long, int wm, int wn, int g, int t, const double (&acc)[8][4][2]) {
    epilogue_scatter<V>(q, sc, tm, tn, wm, wn, g, t, acc, [&](int dest) { return GlobalStore{sc.peer[dest]}; });
  });
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

static bool encode(CUtensorMap* tm, const MapDesc& m) {
  EncodeFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[5], strides[4];
  cuuint32_t box[5];
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  for (int d = 0; d < m.rank; ++d) { dims[d] = m.dims[d]; box[d] = m.box[d]; }
  for (int d = 0; d + 1 < m.rank; ++d) strides[d] = m.strides_bytes[d];
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)m.rank, const_cast<void*>(m.base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, m.swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int V>
static int launch_variant(cudaStream_t s, const CUtensorMap& tmA, const CUtensorMap& tmB, const Args& q, int sms) {
  static bool attr = false;
  if (!attr) {
    JFX_CUDA_OK(cudaFuncSetAttribute(dgemm_dmma_fold<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    attr = true;
  }
  const long long tiles = (long long)q.tiles_n * q.tiles_m * q.batch;
  const unsigned ctas = (unsigned)(tiles < sms ? tiles : sms);
  dgemm_dmma_fold<V><<<ctas, THREADS, SMEM_BYTES, s>>>(tmA, tmB, q);
  JFX_CUDA_OK(cudaGetLastError());
  return 1;
}

}  // namespace fold

// ---- plan-time object ------------------------------------------------------------------------------
struct FoldPlan {
  fold::FoldedTable host;     // geometry (the host copies of the tables are released after the upload)
  double* d_nn = nullptr;     // device copies of the two layouts
  double* d_nt = nullptr;
};

// Read at every plan creation (not cached), so one process can hold folded and plain plans side by side.
bool fold_enabled() {
  const char* e = getenv("JFX_DMMA_FOLD");   // default on; JFX_DMMA_FOLD=0 keeps every table pass on dgemm_dmma_tma
  return !(e && e[0] == '0');
}

int fold_plan_create(const double* table, int rows, int cols, FoldPlan** out) {
  *out = nullptr;
  if (rows < 16 || cols < 16) return JFX_OK;   // tiny tables: the plain kernel's tile is already mostly padding
  const fold::FoldInfo fi = fold::analyze(table, rows, cols);
  if (fi.type == fold::FOLD_NONE) return JFX_OK;
  FoldPlan* fp = new FoldPlan;
  fp->host = fold::build(table, rows, cols, fi);
  const size_t bn = fp->host.nn.size() * sizeof(double), bt = fp->host.nt.size() * sizeof(double);
  if (cudaMalloc(&fp->d_nn, bn) != cudaSuccess || cudaMalloc(&fp->d_nt, bt) != cudaSuccess ||
      cudaMemcpy(fp->d_nn, fp->host.nn.data(), bn, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(fp->d_nt, fp->host.nt.data(), bt, cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaGetLastError();
    fold_plan_destroy(fp);
    set_error("fold_plan_create: device allocation / upload of the folded tables failed");
    return JFX_ERR_CUDA;
  }
  std::vector<double>().swap(fp->host.nn);
  std::vector<double>().swap(fp->host.nt);
  *out = fp;
  return JFX_OK;
}

void fold_plan_destroy(FoldPlan* fp) {
  if (!fp) return;
  if (fp->d_nn) cudaFree(fp->d_nn);
  if (fp->d_nt) cudaFree(fp->d_nt);
  delete fp;
}

int fold_plan_type(const FoldPlan* fp) { return fp ? fp->host.type : 0; }

// multiply-adds the folded pass issues relative to the plain contraction: 1/2, plus the asymmetry-correction k-tiles
double fold_plan_flop_fraction(const FoldPlan* fp) {
  if (!fp || fp->host.kfold <= 0) return 1.0;
  const fold::FoldedTable& f = fp->host;
  return 0.5 * (double)(f.kfold + (f.kts_corr ? f.kfold - f.kcorr0 : 0)) / (double)f.kfold;
}
int fold_plan_corrected_modes(const FoldPlan* fp) { return fp && fp->host.kts_corr ? 2 * (fp->host.kfold - fp->host.kcorr0) : 0; }

// 1 = launched, 0 = outside the envelope (caller uses the plain kernel), < 0 = error
int launch_dmma_fold(cudaStream_t s, const FoldPlan* fp, long long outer, long long inner_real, const double* in,
                     double* out) {
  using namespace fold;
  if (!fp) return 0;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    JFX_CUDA_OK(cudaGetDevice(&dev));
    JFX_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const bool nn = inner_real != 1;
  Args q;
  MapDesc mA, mB;
  if (!make_launch(fp->host, nn, outer, inner_real, nn ? fp->d_nn : fp->d_nt, in, out, &q, &mA, &mB)) return 0;
  CUtensorMap tmA, tmB;
  if (!encode(&tmA, mA) || !encode(&tmB, mB)) return 0;
  switch (q.variant) {
    case OUT_NN: return launch_variant<OUT_NN>(s, tmA, tmB, q, sms);
    case IN_NN: return launch_variant<IN_NN>(s, tmA, tmB, q, sms);
    case OUT_NT: return launch_variant<OUT_NT>(s, tmA, tmB, q, sms);
    default: return launch_variant<IN_NT>(s, tmA, tmB, q, sms);
  }
}

// Last-axis pass with the scatter epilogue.  1 = launched, 0 = not applicable (not an NT fold pass / envelope), < 0 = error.
int launch_dmma_fold_scatter(cudaStream_t s, const FoldPlan* fp, long long outer, const double* in, int mode, int parts,
                             int src, int A, int B, double* const* peers) {
  using namespace fold;
  if (!fp || parts < 1 || parts > MAX_PEERS || (mode != 1 && mode != 2) || (long long)A * B != outer) return 0;
  if ((mode == 1 ? B : A) % parts != 0 || src < 0 || src >= parts) return 0;
  int dev = 0, sms = 0;
  JFX_CUDA_OK(cudaGetDevice(&dev));
  JFX_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  Args q;
  MapDesc mA, mB;
  // the C pointer only decides vector-store alignment here: every receive buffer must be 16-byte aligned
  for (int p = 0; p < parts; ++p)
    if (!peers[p] || (reinterpret_cast<uintptr_t>(peers[p]) & 15)) return 0;
  if (!make_launch(fp->host, /*nn=*/false, outer, 1, fp->d_nt, in, peers[src], &q, &mA, &mB)) return 0;
  // Every rank walks its rows in the same order; with split a (mode 2) the destination rank is a / (A / P), so without a
  // rotation all P ranks would store into the SAME peer at any moment (one NVLink ingress port against P - 1 egress ports:
  // measured at P = 8, 512^3: 0.74 ms for this pass instead of 0.31 ms with local stores).  Rank r starts at the tiles of
  // destination r and proceeds cyclically, like the steps of a ring all-to-all.  (Mode 1 splits b, which varies fastest.)
  q.tm_rot = (mode == 2 && q.tiles_m >= parts) ? (int)(((long long)src * q.tiles_m) / parts) : 0;
  Scatter sc{};
  sc.mode = mode; sc.parts = parts; sc.src = src; sc.A = A; sc.B = B;
  for (int p = 0; p < parts; ++p) sc.peer[p] = peers[p];
  CUtensorMap tmA, tmB;
  if (!encode(&tmA, mA) || !encode(&tmB, mB)) return 0;
  const long long tiles = (long long)q.tiles_n * q.tiles_m * q.batch;
  const unsigned ctas = (unsigned)(tiles < sms ? tiles : sms);
  if (q.variant == OUT_NT) {
    JFX_CUDA_OK(cudaFuncSetAttribute(dgemm_dmma_fold_scatter<OUT_NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    dgemm_dmma_fold_scatter<OUT_NT><<<ctas, THREADS, SMEM_BYTES, s>>>(tmA, tmB, q, sc);
  } else if (q.variant == IN_NT) {
    JFX_CUDA_OK(cudaFuncSetAttribute(dgemm_dmma_fold_scatter<IN_NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    dgemm_dmma_fold_scatter<IN_NT><<<ctas, THREADS, SMEM_BYTES, s>>>(tmA, tmB, q, sc);
  } else {
    return 0;
  }
  JFX_CUDA_OK(cudaGetLastError());
  return 1;
}

// ---- complex data on a last table axis (CPLX_NT) ---------------------------------------------------
struct CplxPlan {
  fold::CplxTable host;   // geometry (host copy of the table released after the upload)
  double* d_nt = nullptr;
};

// default (validated on the B200 against the NN route through the C ABI, tools/fold_check --extra: 25-35x faster);
// JFX_CPLX_NT=0 (read at plan creation) keeps the NN order with two real columns per batch, which wastes 63 / 64 of every tile
bool cplx_nt_enabled() {
  const char* e = getenv("JFX_CPLX_NT");
  return !(e && e[0] == '0');
}

int cplx_plan_create(const double* table, int n_out, int n_in, CplxPlan** out) {
  *out = nullptr;
  if (n_out < 1 || n_in < 1) return JFX_OK;
  CplxPlan* cp = new CplxPlan;
  cp->host = fold::build_cplx(table, n_out, n_in);
  const size_t b = cp->host.nt.size() * sizeof(double);
  if (cudaMalloc(&cp->d_nt, b) != cudaSuccess ||
      cudaMemcpy(cp->d_nt, cp->host.nt.data(), b, cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaGetLastError();
    cplx_plan_destroy(cp);
    set_error("cplx_plan_create: device allocation / upload of the table failed");
    return JFX_ERR_CUDA;
  }
  std::vector<double>().swap(cp->host.nt);
  *out = cp;
  return JFX_OK;
}

void cplx_plan_destroy(CplxPlan* cp) {
  if (!cp) return;
  if (cp->d_nt) cudaFree(cp->d_nt);
  delete cp;
}

// rows = complex lines [outer][n_in]; 1 = launched, 0 = outside the envelope, < 0 = error
int launch_dmma_cplx_nt(cudaStream_t s, const CplxPlan* cp, long long outer, const double* in, double* out) {
  using namespace fold;
  if (!cp) return 0;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    JFX_CUDA_OK(cudaGetDevice(&dev));
    JFX_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  Args q;
  MapDesc mA, mB;
  if (!make_launch_cplx(cp->host, outer, cp->d_nt, in, out, &q, &mA, &mB)) return 0;
  CUtensorMap tmA, tmB;
  if (!encode(&tmA, mA) || !encode(&tmB, mB)) return 0;
  return launch_variant<CPLX_NT>(s, tmA, tmB, q, sms);
}

}  // namespace dmma
}  // namespace jfx
