// C ABI of the jfx engine: plan construction, execution, nonlinear-term composition.
// See include/jfx.h for the contract of every entry point.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <memory>
#include <mutex>
#include <utility>
#include <vector>

#include "jfx_common.h"
#include "fft_common.cuh"
#include "dmma_fold_api.h"

namespace jfx {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

struct Pass {
  int axis = 0;
  AxisGeom geom{};
  bool fast = false;
  void* d_table = nullptr;  // owned
  bool table_complex = false;
  bool dmma = false;
  dmma::CplxPlan* cplx = nullptr;  // owned; complex data on a last table axis as one NT launch (default; JFX_CPLX_NT=0 disables)
  dmma::FoldPlan* fold = nullptr;  // owned; parity-folded tables when the table has the mirror symmetry (default; JFX_DMMA_FOLD=0 disables)
  FastParams fp{};
  FastTables* ft = nullptr;  // owned
};

}  // namespace jfx

struct jfx_plan {
  jfx_plan_desc desc{};
  int ndim = 0;
  int64_t shape_in[JFX_MAX_DIMS]{};
  int64_t shape_out[JFX_MAX_DIMS]{};
  std::vector<jfx::Pass> passes;
  size_t buf_bytes = 0;  // one ping-pong buffer
  size_t ws_bytes = 0;
  int slabs = 1;         // > 1: the last two passes run L2-blocked over slabs of the leading axis
  bool pair = false;     // the last two passes run as ONE plane-fused launch (kernels_fft2_pair.cu)
  size_t pair_counter_off = 0, pair_ring_off = 0, pair_ring_bytes = 0;
  double flops = 0, bytes = 0;
  double flops_executed = 0;   // multiply-adds actually issued: folded passes count half
  // host-pointer path (lazy, guarded)
  std::mutex host_mu;
  void* h_in = nullptr;
  void* h_out = nullptr;
  void* h_ws = nullptr;
  ~jfx_plan() {
    for (auto& p : passes) {
      if (p.d_table) cudaFree(p.d_table);
      if (p.fold) jfx::dmma::fold_plan_destroy(p.fold);
      if (p.cplx) jfx::dmma::cplx_plan_destroy(p.cplx);
      if (p.ft) jfx::fast_tables_destroy(p.ft);
    }
    if (h_in) cudaFree(h_in);
    if (h_out) cudaFree(h_out);
    if (h_ws) cudaFree(h_ws);
  }
};

namespace jfx {

static int64_t prod(const int64_t* s, int a, int b) {
  int64_t p = 1;
  for (int i = a; i < b; ++i) p *= s[i];
  return p;
}
static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static size_t slab_target_bytes();

static int build_plan(const jfx_plan_desc* d, jfx_plan* pl) {
  JFX_REQUIRE(d->abi_version == JFX_ABI_VERSION, JFX_ERR_INVALID, "ABI version %d != %d", d->abi_version,
              JFX_ABI_VERSION);
  JFX_REQUIRE(d->ndim >= 1 && d->ndim <= JFX_MAX_DIMS, JFX_ERR_INVALID, "ndim %d out of range", d->ndim);
  JFX_REQUIRE(d->dtype >= JFX_F32 && d->dtype <= JFX_C128, JFX_ERR_INVALID, "bad dtype %d", d->dtype);
  JFX_REQUIRE(d->op >= JFX_OP_FORWARD && d->op <= JFX_OP_APPLY && d->op != JFX_OP_NONLINEAR, JFX_ERR_INVALID,
              "bad op %d (nonlinear terms use jfx_nonlinear_create)", d->op);
  JFX_REQUIRE(d->slab_size <= 1, JFX_ERR_UNSUPPORTED,
              "a plan is a single-device object: slab transforms are jfx_slab objects (jfx_slab_create)");
  pl->desc = *d;
  pl->ndim = d->ndim;
  const bool to_physical = (d->op == JFX_OP_BACKWARD || d->op == JFX_OP_BACKWARD_PRIMITIVE);
  int64_t cur[JFX_MAX_DIMS];
  for (int i = 0; i < d->ndim; ++i) {
    JFX_REQUIRE(d->shape_in[i] >= 0, JFX_ERR_INVALID, "negative extent");
    cur[i] = pl->shape_in[i] = d->shape_in[i];
  }
  const size_t es = dtype_size(d->dtype);
  size_t max_inter = 0;
  pl->flops = 0;
  const int64_t in_elems = prod(cur, 0, d->ndim);

  for (int ax = 0; ax < d->ndim; ++ax) {
    const jfx_axis_desc& a = d->axis[ax];
    if (a.basis == JFX_BASIS_NONE) continue;
    // the pass lives in the plan from the start: an early error return then releases its device tables with the plan
    pl->passes.emplace_back();
    Pass& p = pl->passes.back();
    p.axis = ax;
    int n_in = (int)cur[ax], n_out;
    if (a.basis == JFX_BASIS_TABLE || a.basis == JFX_BASIS_CTABLE) {
      JFX_REQUIRE(a.table != nullptr, JFX_ERR_INVALID, "axis %d: table basis without a table", ax);
      JFX_REQUIRE(a.table_cols == n_in, JFX_ERR_INVALID, "axis %d: table has %d columns, array extent is %d", ax,
                  a.table_cols, n_in);
      JFX_REQUIRE(a.table_rows >= 0, JFX_ERR_INVALID, "axis %d: negative table rows", ax);
      n_out = a.table_rows;
      p.table_complex = (a.basis == JFX_BASIS_CTABLE);
      JFX_REQUIRE(!(p.table_complex && !dtype_is_complex(d->dtype)), JFX_ERR_INVALID,
                  "axis %d: complex table needs a complex array dtype", ax);
      const size_t tes = (dtype_is_double(d->dtype) ? 8 : 4) * (p.table_complex ? 2 : 1);
      const size_t tbytes = (size_t)n_out * n_in * tes;
      if (tbytes) {
        JFX_CUDA_OK(cudaMalloc(&p.d_table, tbytes));
        // the host table is always double precision; narrow for f32 plans
        if (dtype_is_double(d->dtype)) {
          JFX_CUDA_OK(cudaMemcpy(p.d_table, a.table, tbytes, cudaMemcpyHostToDevice));
        } else {
          const size_t cnt = (size_t)n_out * n_in * (p.table_complex ? 2 : 1);
          std::vector<float> tmp(cnt);
          const double* src = (const double*)a.table;
          for (size_t i = 0; i < cnt; ++i) tmp[i] = (float)src[i];
          JFX_CUDA_OK(cudaMemcpy(p.d_table, tmp.data(), tbytes, cudaMemcpyHostToDevice));
        }
      }
    } else if (a.basis == JFX_BASIS_CHEBYSHEV || a.basis == JFX_BASIS_FOURIER) {
      JFX_REQUIRE(d->op != JFX_OP_APPLY, JFX_ERR_INVALID, "axis %d: APPLY plans take table bases only", ax);
      JFX_REQUIRE(a.n_quad >= a.n_modes && a.n_modes >= 1, JFX_ERR_INVALID, "axis %d: need n_quad >= n_modes >= 1", ax);
      JFX_REQUIRE(!(a.basis == JFX_BASIS_FOURIER && !dtype_is_complex(d->dtype)), JFX_ERR_INVALID,
                  "axis %d: Fourier axes need a complex array dtype", ax);
      JFX_REQUIRE(fast_available(a.basis, a.n_quad, d->dtype), JFX_ERR_UNSUPPORTED,
                  "axis %d: no fast kernel for basis %d at n=%d; supply a dense table instead", ax, a.basis, a.n_quad);
      p.fast = true;
      const bool cheb = a.basis == JFX_BASIS_CHEBYSHEV;
      if (to_physical) {
        JFX_REQUIRE(n_in <= a.n_modes, JFX_ERR_INVALID, "axis %d: %d coefficients > n_modes %d", ax, n_in, a.n_modes);
        n_out = a.n_quad;
        p.fp.kind = cheb ? FAST_CHEB_BACKWARD : FAST_FOURIER_BACKWARD;
        p.fp.n_modes = n_in;
      } else {
        JFX_REQUIRE(n_in == a.n_quad, JFX_ERR_INVALID, "axis %d: physical extent %d != n_quad %d", ax, n_in, a.n_quad);
        n_out = a.n_modes;
        const bool sp = d->op == JFX_OP_SCALAR_PRODUCT;
        p.fp.kind = cheb ? (sp ? FAST_CHEB_SCALAR : FAST_CHEB_FORWARD) : (sp ? FAST_FOURIER_SCALAR : FAST_FOURIER_FORWARD);
        p.fp.n_modes = a.n_modes;
      }
      p.fp.n_quad = a.n_quad;
      p.fp.deriv = (d->op == JFX_OP_BACKWARD_PRIMITIVE) ? a.deriv : 0;
      p.fp.domain_factor = a.domain_factor;
      int rc = fast_tables_create(p.fp, d->dtype, &p.ft);
      if (rc != JFX_OK) return rc;
    } else {
      set_error("axis %d: unknown basis %d", ax, a.basis);
      return JFX_ERR_INVALID;
    }
    p.geom.outer = prod(cur, 0, ax);
    p.geom.inner = prod(cur, ax + 1, d->ndim);
    p.geom.n_in = n_in;
    p.geom.n_out = n_out;
    if (p.fast)
      JFX_REQUIRE(fast_geometry_ok(p.geom, d->dtype), JFX_ERR_UNSUPPORTED,
                  "axis %d: real data with odd inner extent %lld has no fast kernel; supply a dense table", ax,
                  (long long)p.geom.inner);
    if (!p.fast) {
      p.dmma = table_apply_uses_dmma(p.geom, d->dtype, p.table_complex);
      if (p.dmma && d->dtype == JFX_C128 && p.geom.inner == 1 && a.table != nullptr && dmma::cplx_nt_enabled()) {
        int rc = dmma::cplx_plan_create((const double*)a.table, n_out, n_in, &p.cplx);
        if (rc != JFX_OK) return rc;
      }
      if (p.dmma && !p.cplx && dmma::fold_enabled() && a.table != nullptr) {
        int rc = dmma::fold_plan_create((const double*)a.table, n_out, n_in, &p.fold);
        if (rc != JFX_OK) return rc;
      }
      const double cm = dtype_is_complex(d->dtype) ? (p.table_complex ? 4.0 : 2.0) : 1.0;
      const double fl = 2.0 * cm * (double)p.geom.outer * p.geom.inner * (double)n_in * n_out;
      pl->flops += fl;
      pl->flops_executed += p.fold ? dmma::fold_plan_flop_fraction(p.fold) * fl : fl;
    }
    cur[ax] = n_out;
    max_inter = std::max(max_inter, (size_t)prod(cur, 0, d->ndim) * es);
  }
  for (int i = 0; i < d->ndim; ++i) pl->shape_out[i] = cur[i];
  const int64_t out_elems = prod(cur, 0, d->ndim);
  pl->bytes = (double)es * ((double)in_elems + (double)out_elems);
  pl->buf_bytes = pl->passes.size() > 1 ? align_up(max_inter, 256) : 0;
  pl->ws_bytes = pl->passes.size() > 2 ? 2 * pl->buf_bytes : pl->buf_bytes;
  // L2 blocking of the last two passes (see execute_plan): both on axes >= 1, both bandwidth-bound
  // (fast transforms), intermediate larger than what the L2 keeps anyway
  pl->slabs = 1;
  const size_t npass = pl->passes.size();
  if (npass >= 2 && slab_target_bytes() > 0) {
    const Pass& A = pl->passes[npass - 2];
    const Pass& B = pl->passes[npass - 1];
    const int64_t L = pl->shape_out[0];
    const size_t inter = (size_t)A.geom.outer * A.geom.n_out * A.geom.inner * es;
    if (A.axis >= 1 && B.axis >= 1 && A.fast && B.fast && L >= 2 && inter > 2 * slab_target_bytes()) {
      int64_t want = (int64_t)((inter + slab_target_bytes() - 1) / slab_target_bytes());
      pl->slabs = (int)std::min<int64_t>(want, L);
    }
  }
  // L2 reuse between consecutive fast passes whose tiles are both ordered by the leading array axis (axes >= 1 of a 3-D
  // array): the second one walks its tiles last-to-first, so it starts on the planes the first one wrote last — about a third
  // of a 134 MB field is still in the 126 MB L2 (dirty) when a pass ends.  JFX_FFT_REVERSE=0 (read at plan creation) disables it.
  {
    const char* e = getenv("JFX_FFT_REVERSE");
    const bool on = !(e && e[0] == '0');
    bool prev_rev = false;
    for (size_t i = 1; i < npass && on; ++i) {
      Pass& A = pl->passes[i - 1];
      Pass& B = pl->passes[i];
      const bool both = A.fast && B.fast && A.axis >= 1 && B.axis >= 1 && d->ndim >= 3;
      B.fp.reverse = (both && !prev_rev) ? 1 : 0;     // alternate: a reversed pass ends on the FIRST planes
      prev_rev = B.fp.reverse != 0;
    }
  }
  // plane-fused pair (preferred over slabs): both passes fast, same length, no padding
  // opt-in (JFX_PAIR=1, read at plan creation): measured on par with / slightly slower than two plain passes
  // at 256^3 — the single passes are latency-bound, not DRAM-bound, so saving the HBM round trip of the
  // intermediate does not pay yet
  const char* pair_env = getenv("JFX_PAIR");
  const bool pair_on = pair_env && pair_env[0] == '1';
  if (pair_on && npass >= 2 && pl->slabs == 1) {
    const Pass& A = pl->passes[npass - 2];
    const Pass& B = pl->passes[npass - 1];
    if (A.fast && B.fast && A.axis >= 1 && B.axis == d->ndim - 1 && A.axis == d->ndim - 2 && pl->shape_out[0] >= 8) {
      FftArgs fa, fb;
      bool ea, eb;
      // dummy non-null pointers: only the geometry matters for the query
      if (make_fft_args(A.geom, d->dtype, A.fp, A.ft, (void*)16, (void*)16, &fa, &ea) == JFX_OK &&
          make_fft_args(B.geom, d->dtype, B.fp, B.ft, (void*)16, (void*)16, &fb, &eb) == JFX_OK && !ea && !eb) {
        // ring: up to 64 MB of the intermediate (L2-resident), counters after it
        const size_t inter = (size_t)A.geom.outer * A.geom.n_out * A.geom.inner * es;
        const size_t plane_bytes = inter / (size_t)pl->shape_out[0];
        size_t ring_bytes = std::min<size_t>(inter, std::max<size_t>(32u << 20, 64 * plane_bytes));
        ring_bytes = ring_bytes / plane_bytes * plane_bytes;
        if (launch_fast_pair(nullptr, fa, fb, A.fp.n_quad, dtype_is_double(d->dtype), pl->shape_out[0], (void*)16,
                             ring_bytes, (void*)16, /*query=*/true) == 1 && A.fp.n_quad == B.fp.n_quad) {
          pl->pair = true;
          pl->pair_ring_bytes = ring_bytes;
          // workspace layout: [ping-pong buffers as before][ring][counters]
          pl->pair_ring_off = align_up(pl->ws_bytes, 256);
          pl->pair_counter_off = pl->pair_ring_off + align_up(ring_bytes, 256);
          pl->ws_bytes = pl->pair_counter_off + align_up(fast_pair_counter_bytes(pl->shape_out[0]), 256);
        }
      }
    }
  }
  return JFX_OK;
}

static int run_pass_geom(cudaStream_t s, const Pass& p, const AxisGeom& g, int dtype, const void* src, void* dst) {
  if (p.fast) return launch_fast_axis(s, g, dtype, p.fp, p.ft, src, dst);
  if (p.cplx && g.outer * g.n_out != 0) {
    const int rc = dmma::launch_dmma_cplx_nt(s, p.cplx, g.outer, (const double*)src, (double*)dst);
    if (rc < 0) return rc;
    if (rc == 1) return JFX_OK;
  }
  if (p.fold && g.outer * g.inner * g.n_out != 0) {
    const int rc = dmma::launch_dmma_fold(s, p.fold, g.outer, g.inner * (dtype == JFX_C128 ? 2 : 1), (const double*)src,
                                          (double*)dst);
    if (rc < 0) return rc;
    if (rc == 1) return JFX_OK;
  }
  return launch_table_apply(s, g, dtype, p.d_table, p.table_complex, src, dst, nullptr);
}
static int run_pass(cudaStream_t s, const Pass& p, int dtype, const void* src, void* dst) {
  return run_pass_geom(s, p, p.geom, dtype, src, dst);
}

// L2-blocked execution of the last two passes.  When both act on axes >= 1 the leading array axis is a
// pure batch dimension for them, so the array can be walked in slabs of leading-axis planes: pass A
// writes a slab-sized scratch that pass B consumes while it is still resident in the 126 MB L2.  The
// intermediate array of the pair then never travels to HBM (one read + one write of the field saved).
static size_t slab_target_bytes() {
  static const size_t v = [] {
    const char* e = getenv("JFX_SLAB_MB");
    const long mb = e ? atol(e) : 0;   // opt-in: measured slower than plain passes at 256^3 (launch tails), see DESIGN.md
    return (size_t)(mb < 0 ? 0 : mb) << 20;
  }();
  return v;
}

static int execute_plan(const jfx_plan* pl, cudaStream_t s, const void* in, void* out, void* ws) {
  const size_t np = pl->passes.size();
  const size_t es = dtype_size(pl->desc.dtype);
  if (np == 0) {
    const size_t bytes = (size_t)prod(pl->shape_in, 0, pl->ndim) * es;
    if (bytes) JFX_CUDA_OK(cudaMemcpyAsync(out, in, bytes, cudaMemcpyDeviceToDevice, s));
    return JFX_OK;
  }
  JFX_REQUIRE(pl->ws_bytes == 0 || ws != nullptr, JFX_ERR_INVALID, "plan needs a %zu byte workspace", pl->ws_bytes);
  char* w0 = (char*)ws;
  char* w1 = w0 + pl->buf_bytes;
  const void* src = in;
  size_t first_pair = np;   // index of pass A when the last two passes run slab-blocked
  if (pl->slabs > 1 || pl->pair) first_pair = np - 2;
  for (size_t i = 0; i < first_pair; ++i) {
    void* dst = (i + 1 == np) ? out : (void*)((i & 1) ? w1 : w0);
    int rc = run_pass(s, pl->passes[i], pl->desc.dtype, src, dst);
    if (rc != JFX_OK) return rc;
    src = dst;
  }
  if (pl->pair) {
    const Pass& A = pl->passes[np - 2];
    const Pass& B = pl->passes[np - 1];
    FftArgs fa, fb;
    bool ea, eb;
    char* ring = w0 + pl->pair_ring_off;
    int rc = make_fft_args(A.geom, pl->desc.dtype, A.fp, A.ft, src, ring, &fa, &ea);
    if (rc != JFX_OK) return rc;
    rc = make_fft_args(B.geom, pl->desc.dtype, B.fp, B.ft, ring, out, &fb, &eb);
    if (rc != JFX_OK) return rc;
    rc = launch_fast_pair(s, fa, fb, A.fp.n_quad, dtype_is_double(pl->desc.dtype), pl->shape_out[0], ring,
                          pl->pair_ring_bytes, w0 + pl->pair_counter_off, false);
    if (rc < 0) return rc;
    JFX_REQUIRE(rc == 1, JFX_ERR_UNSUPPORTED, "plane-fused pass refused a configuration it accepted at plan time");
    return JFX_OK;
  }
  if (first_pair < np) {
    const Pass& A = pl->passes[np - 2];
    const Pass& B = pl->passes[np - 1];
    const int64_t L = pl->shape_out[0];
    // scratch = the ping-pong buffer `src` does not live in (src is `in` when the pair is the whole plan)
    char* scratch = (src == (const void*)w0) ? w1 : w0;
    const int64_t oA = A.geom.outer / L, oB = B.geom.outer / L;
    const size_t inA = (size_t)oA * A.geom.n_in * A.geom.inner * es;     // bytes per leading-axis plane
    const size_t outB = (size_t)oB * B.geom.n_out * B.geom.inner * es;
    const int64_t per = (L + pl->slabs - 1) / pl->slabs;
    for (int64_t a = 0; a < L; a += per) {
      const int64_t b = std::min<int64_t>(L, a + per);
      AxisGeom gA = A.geom, gB = B.geom;
      gA.outer = (b - a) * oA;
      gB.outer = (b - a) * oB;
      int rc = run_pass_geom(s, A, gA, pl->desc.dtype, (const char*)src + (size_t)a * inA, scratch);
      if (rc != JFX_OK) return rc;
      rc = run_pass_geom(s, B, gB, pl->desc.dtype, scratch, (char*)out + (size_t)a * outB);
      if (rc != JFX_OK) return rc;
    }
  }
  return JFX_OK;
}

}  // namespace jfx

// =================================================================================================
struct jfx_nonlinear {
  std::vector<jfx_plan*> leaves;
  jfx_plan* final_plan = nullptr;
  // ---- row-fused path (kernels_fused.cu): last axis Fourier, leaves + pointwise + forward in one kernel
  bool fused = false;
  std::vector<jfx_plan*> pre_plans;        // one per leaf group: the other axes -> physical (may be empty)
  jfx_plan* post_plan = nullptr;           // the other axes of the final transform (may be null)
  std::vector<jfx::FastTables*> tables;    // owned: twiddles + per-leaf derivative multipliers
  jfx::FusedRowArgs fargs{};
  int fused_n = 0;
  size_t pre_bytes = 0, row_out_bytes = 0; // one pre-array / the row kernel's output
  size_t scratch_bytes = 0;                // L2-resident leaf lines of the persistent grid
  int n_groups = 0;
  jfx::PointwiseProgram prog{};
  const void* statics[JFX_MAX_LEAVES]{};
  int dtype = JFX_F64;
  int64_t phys_elems = 0;
  size_t field_bytes = 0;  // one physical field, aligned
  size_t sub_ws = 0;       // max workspace of the sub-plans
  size_t ws_bytes = 0;
  ~jfx_nonlinear() {
    for (auto* p : leaves) delete p;
    delete final_plan;
    for (auto* p : pre_plans) delete p;
    delete post_plan;
    for (auto* t : tables) jfx::fast_tables_destroy(t);
  }
};

namespace jfx {

// Decide whether the nonlinear term can run row-fused (kernels_fused.cu) and build what that path needs.
// Conditions: complex data, the LAST array axis is a fast Fourier axis in the final transform and in every
// leaf (same n_quad, same coefficient count), and the leaf lines of one row fit in shared memory.
// Leaves are grouped by what they need from the OTHER axes (their derivative orders there): one
// "pre-plan" per group takes those axes to physical space and leaves the last axis in coefficient space.
static int try_fuse_rows(const jfx_nonlinear_desc* d, jfx_nonlinear* nl) {
  static const bool disabled = [] { const char* e = getenv("JFX_NL_FUSE"); return e && e[0] == '0'; }();
  if (disabled) return JFX_OK;
  const jfx_plan_desc* fd = d->final_transform;
  const int nd = fd->ndim, last = nd - 1;
  if (!dtype_is_complex(fd->dtype)) return JFX_OK;
  const jfx_axis_desc& fa = fd->axis[last];
  if (fa.basis != JFX_BASIS_FOURIER) return JFX_OK;
  const int n = fa.n_quad;
  if (n < 48 || !fast_available(JFX_BASIS_FOURIER, n, fd->dtype)) return JFX_OK;
  if (fd->shape_in[last] != n) return JFX_OK;
  const int64_t n_coeff = d->leaves[0]->shape_in[last];
  for (int l = 0; l < d->n_leaves; ++l) {
    const jfx_axis_desc& la = d->leaves[l]->axis[last];
    if (la.basis != JFX_BASIS_FOURIER || la.n_quad != n || d->leaves[l]->shape_in[last] != n_coeff) return JFX_OK;
    if (la.domain_factor != d->leaves[0]->axis[last].domain_factor) return JFX_OK;
  }
  FusedRowArgs& fa_ = nl->fargs;
  fa_ = FusedRowArgs{};
  fa_.n_leaves = d->n_leaves;
  fa_.rows = 1;   // placeholder for the query
  fa_.n_coeff = (int)n_coeff;
  fa_.n_out = fa.n_modes;
  {
    int rc = validate_program(nl->prog, nl->statics);
    if (rc != JFX_OK) return rc;
  }
  fa_.depth = program_depth(nl->prog);
  if (!program_to_poly(nl->prog, &fa_.poly)) fa_.poly.n_terms = 0;
  size_t scratch_bytes = 0;
  if (launch_fused_rows(nullptr, fd->dtype, n, fa_, &scratch_bytes) != 1) return JFX_OK;
  nl->scratch_bytes = align_up(scratch_bytes, 256);

  // ---- groups: leaves with identical specs on the other axes ----------------------------------------
  bool other_axes = false;
  for (int ax = 0; ax < last; ++ax) other_axes |= (fd->axis[ax].basis != JFX_BASIS_NONE);
  std::vector<int> group_rep;   // representative leaf of each group
  for (int l = 0; l < d->n_leaves; ++l) {
    int g = -1;
    for (size_t k = 0; k < group_rep.size() && g < 0; ++k) {
      bool same = true;
      for (int ax = 0; ax < last; ++ax) same &= (d->leaves[l]->axis[ax].deriv == d->leaves[group_rep[k]]->axis[ax].deriv);
      if (same) g = (int)k;
    }
    if (g < 0) { g = (int)group_rep.size(); group_rep.push_back(l); }
    fa_.leaf_group[l] = g;
  }
  nl->n_groups = (int)group_rep.size();
  const size_t es = dtype_size(fd->dtype);
  int64_t rows = 1;
  for (int ax = 0; ax < last; ++ax) rows *= fd->shape_in[ax];
  if (other_axes) {
    for (int rep : group_rep) {
      jfx_plan_desc pd = *d->leaves[rep];
      pd.axis[last] = jfx_axis_desc{};
      pd.axis[last].basis = JFX_BASIS_NONE;
      bool any = false;
      for (int ax = 0; ax < last; ++ax) any |= (pd.axis[ax].deriv != 0);
      pd.op = any ? JFX_OP_BACKWARD_PRIMITIVE : JFX_OP_BACKWARD;
      jfx_plan* p = nullptr;
      int rc = jfx_plan_create(&pd, &p);
      if (rc != JFX_OK) return rc;
      nl->pre_plans.push_back(p);
      nl->sub_ws = std::max(nl->sub_ws, p->ws_bytes);
      for (int ax = 0; ax < last; ++ax)
        JFX_REQUIRE(p->shape_out[ax] == fd->shape_in[ax], JFX_ERR_INVALID, "leaf physical shape differs on axis %d", ax);
    }
    jfx_plan_desc qd = *fd;
    qd.axis[last] = jfx_axis_desc{};
    qd.axis[last].basis = JFX_BASIS_NONE;
    qd.shape_in[last] = fa.n_modes;
    int rc = jfx_plan_create(&qd, &nl->post_plan);
    if (rc != JFX_OK) return rc;
    nl->sub_ws = std::max(nl->sub_ws, nl->post_plan->ws_bytes);
    nl->pre_bytes = align_up((size_t)rows * n_coeff * es, 256);
    nl->row_out_bytes = align_up((size_t)rows * fa.n_modes * es, 256);
  }
  // ---- tables: twiddles and per-leaf derivative multipliers --------------------------------------------
  {
    FastParams fp{};
    fp.kind = FAST_FOURIER_FORWARD; fp.n_modes = fa.n_modes; fp.n_quad = n; fp.deriv = 0; fp.domain_factor = fa.domain_factor;
    FastTables* t = nullptr;
    int rc = fast_tables_create(fp, fd->dtype, &t);
    if (rc != JFX_OK) return rc;
    nl->tables.push_back(t);
    fa_.tw = t->d_tw;
  }
  for (int l = 0; l < d->n_leaves; ++l) {
    const jfx_axis_desc& la = d->leaves[l]->axis[last];
    const int k = d->leaves[l]->op == JFX_OP_BACKWARD_PRIMITIVE ? la.deriv : 0;
    fa_.mult[l] = nullptr;
    if (k > 0) {
      FastParams fp{};
      fp.kind = FAST_FOURIER_BACKWARD; fp.n_modes = (int)n_coeff; fp.n_quad = n; fp.deriv = k; fp.domain_factor = la.domain_factor;
      FastTables* t = nullptr;
      int rc = fast_tables_create(fp, fd->dtype, &t);
      if (rc != JFX_OK) return rc;
      nl->tables.push_back(t);
      fa_.mult[l] = t->d_pre;
    }
  }
  const double PI = 3.14159265358979323846;
  fa_.scale = 1.0 / n;
  if (fd->op == JFX_OP_SCALAR_PRODUCT) fa_.scale *= 2.0 * PI / fa.domain_factor;
  fa_.rows = rows;
  fa_.n_coeff = (int)n_coeff;
  fa_.n_out = fa.n_modes;
  fa_.n_instr = nl->prog.n_instr;
  memcpy(fa_.instr, nl->prog.instr, sizeof(jfx_pw_instr) * nl->prog.n_instr);
  memcpy(fa_.consts, nl->prog.consts, sizeof(fa_.consts));
  for (int i = 0; i < JFX_MAX_LEAVES; ++i) fa_.statics[i] = nl->statics[i];
  {
    int rc = validate_program(nl->prog, nl->statics);
    if (rc != JFX_OK) return rc;
  }
  nl->fused = true;
  nl->fused_n = n;
  nl->ws_bytes = (size_t)nl->n_groups * nl->pre_bytes + nl->row_out_bytes + nl->scratch_bytes + align_up(nl->sub_ws, 256);
  return JFX_OK;
}

}  // namespace jfx

extern "C" {

int jfx_abi_version(void) { return JFX_ABI_VERSION; }
const char* jfx_last_error(void) { return jfx::get_error(); }

int jfx_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int jfx_fast_path_available(int basis, int n, int dtype) { return jfx::fast_available(basis, n, dtype) ? 1 : 0; }

int jfx_plan_create(const jfx_plan_desc* desc, jfx_plan** out) {
  using namespace jfx;
  JFX_REQUIRE(desc && out, JFX_ERR_INVALID, "null argument");
  *out = nullptr;
  JFX_REQUIRE(jfx_device_count() > 0, JFX_ERR_CUDA, "no CUDA device: the jfx engine has no CPU fallback");
  std::unique_ptr<jfx_plan> pl(new (std::nothrow) jfx_plan);
  JFX_REQUIRE(pl, JFX_ERR_NOMEM, "out of host memory");
  int rc = build_plan(desc, pl.get());
  if (rc != JFX_OK) return rc;
  // The constant tables were uploaded with cudaMemcpy from pageable host memory, which returns once the data are STAGED: the
  // DMA into device memory may still be in flight, and it is ordered with the legacy default stream only.  A first execution
  // on the caller's stream could read a partially written table (seen on the B200: 32 wrong values in the first 512^3 pass
  // after plan creation, none afterwards).  Plan creation may synchronise; jfx_execute never does.
  {
    const char* e = getenv("JFX_NO_CREATE_SYNC");   // debugging switch for exactly that race (tools/stress_first_call.py)
    if (!(e && e[0] == '1')) JFX_CUDA_OK(cudaDeviceSynchronize());
  }
  // tables were copied: do not keep dangling host pointers
  for (int i = 0; i < JFX_MAX_DIMS; ++i) pl->desc.axis[i].table = nullptr;
  *out = pl.release();
  return JFX_OK;
}

void jfx_plan_destroy(jfx_plan* plan) { delete plan; }

int jfx_plan_ndim(const jfx_plan* plan) { return plan ? plan->ndim : JFX_ERR_INVALID; }

int jfx_plan_shape_out(const jfx_plan* plan, int64_t* shape_out) {
  using namespace jfx;
  JFX_REQUIRE(plan && shape_out, JFX_ERR_INVALID, "null argument");
  for (int i = 0; i < JFX_MAX_DIMS; ++i) shape_out[i] = i < plan->ndim ? plan->shape_out[i] : 1;
  return JFX_OK;
}

int jfx_plan_workspace_bytes(const jfx_plan* plan, size_t* bytes) {
  using namespace jfx;
  JFX_REQUIRE(plan && bytes, JFX_ERR_INVALID, "null argument");
  *bytes = plan->ws_bytes;
  return JFX_OK;
}

int jfx_plan_work(const jfx_plan* plan, double* flops, double* bytes) {
  using namespace jfx;
  JFX_REQUIRE(plan, JFX_ERR_INVALID, "null argument");
  if (flops) *flops = plan->flops;
  if (bytes) *bytes = plan->bytes;
  return JFX_OK;
}

int jfx_plan_executed_flops(const jfx_plan* plan, double* flops) {
  using namespace jfx;
  JFX_REQUIRE(plan && flops, JFX_ERR_INVALID, "null argument");
  *flops = plan->flops_executed;
  return JFX_OK;
}

int jfx_plan_launches(const jfx_plan* plan) {
  if (!plan) return JFX_ERR_INVALID;
  if (plan->passes.empty()) return 0;
  if (plan->pair) return (int)plan->passes.size() - 1;
  if (plan->slabs > 1) {
    const int64_t L = plan->shape_out[0], per = (L + plan->slabs - 1) / plan->slabs;
    return (int)plan->passes.size() - 2 + 2 * (int)((L + per - 1) / per);
  }
  return (int)plan->passes.size();
}

// ---- slab exchange fused into the last pass ----------------------------------------------------------
static int scatter_geometry(const jfx_plan* pl, int parts, int split_axis, int* mode, int* A, int* B) {
  using namespace jfx;
  JFX_REQUIRE(pl, JFX_ERR_INVALID, "null plan");
  JFX_REQUIRE(pl->ndim == 3 && !pl->passes.empty(), JFX_ERR_UNSUPPORTED, "scatter execution: 3-D plans only");
  JFX_REQUIRE(parts >= 1 && parts <= 8, JFX_ERR_UNSUPPORTED, "scatter execution: 1..8 peers");
  JFX_REQUIRE(split_axis == 0 || split_axis == 1, JFX_ERR_INVALID, "split_axis must be 0 or 1");
  JFX_REQUIRE(pl->slabs == 1 && !pl->pair, JFX_ERR_UNSUPPORTED, "scatter execution: plain pass sequence only");
  const Pass& last = pl->passes.back();
  JFX_REQUIRE(last.axis == 2 && last.fold != nullptr && pl->desc.dtype == JFX_F64, JFX_ERR_UNSUPPORTED,
              "scatter execution needs a parity-folded fp64 table pass along the last axis as the final pass");
  *A = (int)pl->shape_out[0];
  *B = (int)pl->shape_out[1];
  *mode = split_axis == 1 ? 1 : 2;
  JFX_REQUIRE((split_axis == 1 ? *B : *A) % parts == 0, JFX_ERR_INVALID, "split axis %d of extent %d is not divisible by %d",
              split_axis, split_axis == 1 ? *B : *A, parts);
  return JFX_OK;
}

int jfx_plan_scatter_supported(const jfx_plan* plan, int parts, int split_axis) {
  int mode, A, B;
  return scatter_geometry(plan, parts, split_axis, &mode, &A, &B) == JFX_OK ? 1 : 0;
}

int jfx_execute_scatter(const jfx_plan* plan, void* stream, const void* in, void* const* peer_out, int parts, int rank,
                        int split_axis, void* workspace) {
  using namespace jfx;
  JFX_REQUIRE(plan && in && peer_out, JFX_ERR_INVALID, "null argument");
  int mode, A, B;
  int rc = scatter_geometry(plan, parts, split_axis, &mode, &A, &B);
  if (rc != JFX_OK) return rc;
  JFX_REQUIRE(rank >= 0 && rank < parts, JFX_ERR_INVALID, "rank %d outside 0..%d", rank, parts - 1);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t np = plan->passes.size();
  JFX_REQUIRE(np == 1 || workspace != nullptr, JFX_ERR_INVALID, "plan needs a %zu byte workspace", plan->ws_bytes);
  char* w0 = (char*)workspace;
  char* w1 = w0 + plan->buf_bytes;
  const void* src = in;
  for (size_t i = 0; i + 1 < np; ++i) {
    void* dst = (void*)((i & 1) ? w1 : w0);
    rc = run_pass(s, plan->passes[i], plan->desc.dtype, src, dst);
    if (rc != JFX_OK) return rc;
    src = dst;
  }
  const Pass& last = plan->passes.back();
  rc = dmma::launch_dmma_fold_scatter(s, last.fold, last.geom.outer, (const double*)src, mode, parts, rank, A, B,
                                      (double* const*)peer_out);
  if (rc < 0) return rc;
  JFX_REQUIRE(rc == 1, JFX_ERR_UNSUPPORTED, "scatter pass outside the folded kernel's envelope (alignment)");
  return JFX_OK;
}

int jfx_execute(const jfx_plan* plan, void* stream, const void* in, void* out, void* workspace) {
  using namespace jfx;
  JFX_REQUIRE(plan, JFX_ERR_INVALID, "null argument");
  // empty arrays are valid requests (and have null data pointers in most array libraries)
  if (prod(plan->shape_in, 0, plan->ndim) == 0 || prod(plan->shape_out, 0, plan->ndim) == 0) return JFX_OK;
  JFX_REQUIRE(in && out, JFX_ERR_INVALID, "null argument");
  return execute_plan(plan, (cudaStream_t)stream, in, out, workspace);
}

int jfx_execute_host(jfx_plan* plan, void* stream, const void* in_host, void* out_host) {
  using namespace jfx;
  JFX_REQUIRE(plan && in_host && out_host, JFX_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(plan->host_mu);
  const size_t es = dtype_size(plan->desc.dtype);
  const size_t in_b = (size_t)prod(plan->shape_in, 0, plan->ndim) * es;
  const size_t out_b = (size_t)prod(plan->shape_out, 0, plan->ndim) * es;
  if (!plan->h_in && in_b) JFX_CUDA_OK(cudaMalloc(&plan->h_in, in_b));
  if (!plan->h_out && out_b) JFX_CUDA_OK(cudaMalloc(&plan->h_out, out_b));
  if (!plan->h_ws && plan->ws_bytes) JFX_CUDA_OK(cudaMalloc(&plan->h_ws, plan->ws_bytes));
  cudaStream_t s = (cudaStream_t)stream;
  if (in_b) JFX_CUDA_OK(cudaMemcpyAsync(plan->h_in, in_host, in_b, cudaMemcpyHostToDevice, s));
  if (in_b && out_b) {
    int rc = execute_plan(plan, s, plan->h_in, plan->h_out, plan->h_ws);
    if (rc != JFX_OK) return rc;
  }
  if (out_b) JFX_CUDA_OK(cudaMemcpyAsync(out_host, plan->h_out, out_b, cudaMemcpyDeviceToHost, s));
  JFX_CUDA_OK(cudaStreamSynchronize(s));
  return JFX_OK;
}

int jfx_host_alloc(void** ptr, size_t bytes) {
  using namespace jfx;
  JFX_REQUIRE(ptr, JFX_ERR_INVALID, "null argument");
  JFX_CUDA_OK(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
  return JFX_OK;
}
int jfx_host_free(void* ptr) {
  using namespace jfx;
  if (ptr) JFX_CUDA_OK(cudaFreeHost(ptr));
  return JFX_OK;
}

// ---- nonlinear terms -----------------------------------------------------------------------
int jfx_nonlinear_create(const jfx_nonlinear_desc* d, jfx_nonlinear** out) {
  using namespace jfx;
  JFX_REQUIRE(d && out, JFX_ERR_INVALID, "null argument");
  *out = nullptr;
  JFX_REQUIRE(d->abi_version == JFX_ABI_VERSION, JFX_ERR_INVALID, "ABI version mismatch");
  JFX_REQUIRE(d->n_leaves >= 1 && d->n_leaves <= JFX_MAX_LEAVES, JFX_ERR_INVALID, "n_leaves must be 1..%d", JFX_MAX_LEAVES);
  JFX_REQUIRE(d->final_transform, JFX_ERR_INVALID, "missing final transform");
  JFX_REQUIRE(d->n_program >= 1 && d->n_program <= JFX_MAX_PROGRAM, JFX_ERR_INVALID, "bad program length");
  JFX_REQUIRE(d->n_consts >= 0 && d->n_consts <= 32, JFX_ERR_INVALID, "bad constant count");
  std::unique_ptr<jfx_nonlinear> nl(new (std::nothrow) jfx_nonlinear);
  JFX_REQUIRE(nl, JFX_ERR_NOMEM, "out of host memory");
  for (int l = 0; l < d->n_leaves; ++l) {
    JFX_REQUIRE(d->leaves[l], JFX_ERR_INVALID, "leaf %d missing", l);
    JFX_REQUIRE(d->leaves[l]->op == JFX_OP_BACKWARD || d->leaves[l]->op == JFX_OP_BACKWARD_PRIMITIVE,
                JFX_ERR_INVALID, "leaf %d must be a backward transform", l);
    jfx_plan* p = nullptr;
    int rc = jfx_plan_create(d->leaves[l], &p);
    if (rc != JFX_OK) return rc;
    nl->leaves.push_back(p);
  }
  JFX_REQUIRE(d->final_transform->op == JFX_OP_FORWARD || d->final_transform->op == JFX_OP_SCALAR_PRODUCT,
              JFX_ERR_INVALID, "final transform must be forward or scalar_product");
  int rc = jfx_plan_create(d->final_transform, &nl->final_plan);
  if (rc != JFX_OK) return rc;
  nl->dtype = d->final_transform->dtype;
  const jfx_plan* f = nl->final_plan;
  nl->phys_elems = prod(f->shape_in, 0, f->ndim);
  for (auto* p : nl->leaves) {
    JFX_REQUIRE(p->desc.dtype == nl->dtype, JFX_ERR_INVALID, "leaf dtype differs from final transform dtype");
    JFX_REQUIRE(p->ndim == f->ndim, JFX_ERR_INVALID, "leaf rank differs");
    for (int i = 0; i < f->ndim; ++i)
      JFX_REQUIRE(p->shape_out[i] == f->shape_in[i], JFX_ERR_INVALID, "leaf physical shape differs on axis %d", i);
    for (int i = 0; i < f->ndim; ++i)
      JFX_REQUIRE(p->shape_in[i] == nl->leaves[0]->shape_in[i], JFX_ERR_INVALID, "leaf coefficient shapes differ");
    nl->sub_ws = std::max(nl->sub_ws, p->ws_bytes);
  }
  nl->sub_ws = std::max(nl->sub_ws, f->ws_bytes);
  nl->prog.n_instr = d->n_program;
  memcpy(nl->prog.instr, d->program, sizeof(jfx_pw_instr) * d->n_program);
  nl->prog.n_consts = d->n_consts;
  memcpy(nl->prog.consts, d->consts, sizeof(d->consts));
  nl->prog.n_leaves = d->n_leaves;
  for (int i = 0; i < JFX_MAX_LEAVES; ++i) nl->statics[i] = i < d->n_statics ? d->statics[i] : nullptr;
  nl->field_bytes = align_up((size_t)nl->phys_elems * dtype_size(nl->dtype), 256);
  // workspace = leaf fields + pointwise result + sub-plan workspace
  nl->ws_bytes = (size_t)(d->n_leaves + 1) * nl->field_bytes + align_up(nl->sub_ws, 256);
  {
    const int rc = try_fuse_rows(d, nl.get());
    if (rc != JFX_OK) return rc;
  }
  JFX_CUDA_OK(cudaDeviceSynchronize());   // table uploads complete before the first execution (see jfx_plan_create)
  *out = nl.release();
  return JFX_OK;
}

void jfx_nonlinear_destroy(jfx_nonlinear* nl) { delete nl; }

int jfx_nonlinear_workspace_bytes(const jfx_nonlinear* nl, size_t* bytes) {
  using namespace jfx;
  JFX_REQUIRE(nl && bytes, JFX_ERR_INVALID, "null argument");
  *bytes = nl->ws_bytes;
  return JFX_OK;
}

int jfx_nonlinear_shape_out(const jfx_nonlinear* nl, int64_t* shape_out, int* ndim) {
  using namespace jfx;
  JFX_REQUIRE(nl && shape_out, JFX_ERR_INVALID, "null argument");
  if (ndim) *ndim = nl->final_plan->ndim;
  return jfx_plan_shape_out(nl->final_plan, shape_out);
}

int jfx_nonlinear_launches(const jfx_nonlinear* nl) {
  if (!nl) return JFX_ERR_INVALID;
  if (nl->fused) {
    int n = 1 + (nl->post_plan ? jfx_plan_launches(nl->post_plan) : 0);
    for (auto* p : nl->pre_plans) n += jfx_plan_launches(p);
    return n;
  }
  int n = 1 + jfx_plan_launches(nl->final_plan);
  for (auto* p : nl->leaves) n += std::max(1, jfx_plan_launches(p));
  return n;
}

int jfx_nonlinear_execute(const jfx_nonlinear* nl, void* stream, const void* uh, void* out, void* workspace) {
  using namespace jfx;
  JFX_REQUIRE(nl && uh && out && workspace, JFX_ERR_INVALID, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  char* base = (char*)workspace;
  if (nl->fused) {
    FusedRowArgs fa = nl->fargs;
    fa.scratch = base + (size_t)nl->n_groups * nl->pre_bytes + nl->row_out_bytes;
    char* sub = (char*)fa.scratch + nl->scratch_bytes;
    if (nl->pre_plans.empty()) {
      for (int g = 0; g < nl->n_groups; ++g) fa.src[g] = uh;
      fa.out = out;
    } else {
      for (int g = 0; g < nl->n_groups; ++g) {
        void* dst = base + (size_t)g * nl->pre_bytes;
        int rc = execute_plan(nl->pre_plans[g], s, uh, dst, sub);
        if (rc != JFX_OK) return rc;
        fa.src[g] = dst;
      }
      fa.out = base + (size_t)nl->n_groups * nl->pre_bytes;
    }
    int rc = launch_fused_rows(s, nl->dtype, nl->fused_n, fa, nullptr);
    if (rc < 0) return rc;
    JFX_REQUIRE(rc == 1, JFX_ERR_UNSUPPORTED, "row-fused nonlinear kernel refused a configuration it accepted at plan time");
    if (nl->post_plan) return execute_plan(nl->post_plan, s, fa.out, out, sub);
    return JFX_OK;
  }
  const size_t nleaf = nl->leaves.size();
  char* sub = base + (nleaf + 1) * nl->field_bytes;
  const void* fields[JFX_MAX_LEAVES];
  for (size_t l = 0; l < nleaf; ++l) {
    void* f = base + l * nl->field_bytes;
    int rc = execute_plan(nl->leaves[l], s, uh, f, sub);
    if (rc != JFX_OK) return rc;
    fields[l] = f;
  }
  void* e = base + nleaf * nl->field_bytes;
  int rc = launch_pointwise(s, nl->prog, fields, nl->statics, e, nl->phys_elems, nl->dtype);
  if (rc != JFX_OK) return rc;
  return execute_plan(nl->final_plan, s, e, out, sub);
}

int jfx_pointwise(void* stream, const jfx_pw_instr* program, int n_program, const double (*consts)[2],
                  int n_consts, const void* const* leaves, int n_leaves, const void* const* statics,
                  void* out, int64_t n, int dtype) {
  using namespace jfx;
  JFX_REQUIRE(program && leaves && out, JFX_ERR_INVALID, "null argument");
  JFX_REQUIRE(n_program >= 1 && n_program <= JFX_MAX_PROGRAM, JFX_ERR_INVALID, "bad program length");
  JFX_REQUIRE(n_consts >= 0 && n_consts <= 32 && n_leaves >= 0 && n_leaves <= JFX_MAX_LEAVES, JFX_ERR_INVALID, "bad counts");
  PointwiseProgram p{};
  p.n_instr = n_program;
  memcpy(p.instr, program, sizeof(jfx_pw_instr) * n_program);
  p.n_consts = n_consts;
  if (n_consts) memcpy(p.consts, consts, sizeof(double) * 2 * n_consts);
  p.n_leaves = n_leaves;
  return launch_pointwise((cudaStream_t)stream, p, leaves, statics, out, n, dtype);
}

// ---- slab / stage arithmetic / calibration ---------------------------------------------------
int jfx_slab_pack(void* stream, const void* in, void* out, const int64_t* shape, int ndim, int split_axis,
                  int parts, int dtype) {
  using namespace jfx;
  JFX_REQUIRE(in && out && shape, JFX_ERR_INVALID, "null argument");
  return launch_slab_pack((cudaStream_t)stream, in, out, shape, ndim, split_axis, parts, dtype);
}
int jfx_slab_unpack(void* stream, const void* in, void* out, const int64_t* shape_out, int ndim, int concat_axis,
                    int parts, int dtype) {
  using namespace jfx;
  JFX_REQUIRE(in && out && shape_out, JFX_ERR_INVALID, "null argument");
  return launch_slab_unpack((cudaStream_t)stream, in, out, shape_out, ndim, concat_axis, parts, dtype);
}
int jfx_axpby_diag(void* stream, int n_terms, const void* const* coeff, const double* alpha, const void* const* x,
                   void* out, int64_t n, int dtype, int coeff_is_complex) {
  using namespace jfx;
  JFX_REQUIRE(x && out, JFX_ERR_INVALID, "null argument");
  return launch_axpby_diag((cudaStream_t)stream, n_terms, coeff, alpha, x, out, n, dtype, coeff_is_complex);
}
// ---- plan registry: serialisable keys instead of raw plan pointers (XLA FFI attributes) ------------------------------
namespace {
struct RegistryEntry {
  jfx_plan_desc desc{};
  std::vector<std::vector<unsigned char>> tables;     // deep copies of the host tables, one per axis (may be empty)
  std::vector<std::pair<int, jfx_plan*>> plans;       // (device ordinal, plan)
};
std::mutex g_reg_mu;
std::vector<std::pair<uint64_t, std::unique_ptr<RegistryEntry>>> g_registry;

uint64_t fnv1a(uint64_t h, const void* p, size_t n) {
  const unsigned char* b = (const unsigned char*)p;
  for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
  return h;
}
size_t table_bytes(const jfx_plan_desc& d, int ax) {
  const jfx_axis_desc& a = d.axis[ax];
  if (!a.table || (a.basis != JFX_BASIS_TABLE && a.basis != JFX_BASIS_CTABLE)) return 0;
  return (size_t)a.table_rows * (size_t)a.table_cols * (a.basis == JFX_BASIS_CTABLE ? 16 : 8);
}
}  // namespace

int jfx_registry_register(const jfx_plan_desc* desc, uint64_t* key_out) {
  using namespace jfx;
  JFX_REQUIRE(desc && key_out, JFX_ERR_INVALID, "null argument");
  JFX_REQUIRE(desc->ndim >= 1 && desc->ndim <= JFX_MAX_DIMS, JFX_ERR_INVALID, "ndim %d out of range", desc->ndim);
  // key = hash of every field that defines the plan, incl. the table CONTENTS (never of a pointer value)
  uint64_t h = 1469598103934665603ull;
  const int32_t head[4] = {desc->abi_version, desc->op, desc->dtype, desc->ndim};
  h = fnv1a(h, head, sizeof(head));
  h = fnv1a(h, desc->shape_in, sizeof(int64_t) * desc->ndim);
  for (int ax = 0; ax < desc->ndim; ++ax) {
    const jfx_axis_desc& a = desc->axis[ax];
    const int32_t f[6] = {a.basis, a.n_modes, a.n_quad, a.deriv, a.table_rows, a.table_cols};
    h = fnv1a(h, f, sizeof(f));
    h = fnv1a(h, &a.domain_factor, sizeof(double));
    const size_t tb = table_bytes(*desc, ax);
    if (tb) h = fnv1a(h, a.table, tb);
  }
  std::lock_guard<std::mutex> lk(g_reg_mu);
  for (auto& e : g_registry)
    if (e.first == h) { *key_out = h; return JFX_OK; }
  std::unique_ptr<RegistryEntry> e(new (std::nothrow) RegistryEntry);
  JFX_REQUIRE(e, JFX_ERR_NOMEM, "out of host memory");
  e->desc = *desc;
  e->tables.resize(JFX_MAX_DIMS);
  for (int ax = 0; ax < desc->ndim; ++ax) {
    const size_t tb = table_bytes(*desc, ax);
    if (tb) {
      e->tables[ax].assign((const unsigned char*)desc->axis[ax].table, (const unsigned char*)desc->axis[ax].table + tb);
      e->desc.axis[ax].table = e->tables[ax].data();
    } else {
      e->desc.axis[ax].table = nullptr;
    }
  }
  g_registry.emplace_back(h, std::move(e));
  *key_out = h;
  return JFX_OK;
}

int jfx_registry_acquire(uint64_t key, const jfx_plan** out) {
  using namespace jfx;
  JFX_REQUIRE(out, JFX_ERR_INVALID, "null argument");
  *out = nullptr;
  int dev = 0;
  JFX_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_reg_mu);
  for (auto& e : g_registry) {
    if (e.first != key) continue;
    for (auto& p : e.second->plans)
      if (p.first == dev) { *out = p.second; return JFX_OK; }
    jfx_plan* pl = nullptr;                      // first use on this device: build its tables here (may synchronise)
    const int rc = jfx_plan_create(&e.second->desc, &pl);
    if (rc != JFX_OK) return rc;
    e.second->plans.emplace_back(dev, pl);
    *out = pl;
    return JFX_OK;
  }
  set_error("plan key %llu is not registered in this process (jfx_registry_register)", (unsigned long long)key);
  return JFX_ERR_INVALID;
}

void jfx_registry_clear(void) {
  std::lock_guard<std::mutex> lk(g_reg_mu);
  for (auto& e : g_registry)
    for (auto& p : e.second->plans) delete p.second;
  g_registry.clear();
}

int jfx_point_contract(void* stream, const void* y, const void* w, void* out, int64_t outer, int32_t n, int64_t points,
                       int dtype, int w_is_complex) {
  using namespace jfx;
  JFX_REQUIRE(y && w && out, JFX_ERR_INVALID, "null argument");
  JFX_REQUIRE(dtype >= JFX_F32 && dtype <= JFX_C128, JFX_ERR_INVALID, "bad dtype %d", dtype);
  return launch_point_contract((cudaStream_t)stream, y, w, out, outer, n, points, dtype, w_is_complex);
}
// ---- slab-decomposed transform behind one call (sharding.py:43-105 of the reference) ------------------------------
}  // extern "C"

struct jfx_slab {
  int rank = 0, size = 1, sharding = 0, ndim = 0, dtype = JFX_F64;
  jfx_plan* phase1 = nullptr;   // the unsharded axes on the local input block
  jfx_plan* phase2 = nullptr;   // the originally sharded axis on the exchanged block
  bool fused = false;           // phase 1 ends in the scatter epilogue (peer stores); else 2-D peer copies
  int64_t mid[JFX_MAX_DIMS]{};  // phase-1 output shape
  int64_t exch[JFX_MAX_DIMS]{}; // exchanged block = phase-2 input shape
  size_t mid_bytes = 0, recv_bytes = 0, ws1_off = 0, ws2_off = 0, ws_bytes = 0;
  std::vector<void*> recv[2];   // [turn][peer]: receive buffers, peer-mapped on this device
  std::vector<void*> signal;    // [peer]: flag words (size * 4 bytes each, zero-initialised), peer-mapped
  void** d_signal = nullptr;    // device copy of the pad table (read by the barrier kernel)
  int turn = 0;
  std::mutex mu;
  ~jfx_slab() {
    delete phase1;
    delete phase2;
    if (d_signal) cudaFree(d_signal);
  }
};

namespace jfx {

// Device-side barrier over the ranks of the box (one thread per peer): raise my flag in the peer's pad, then wait for the
// peer's flag in mine and lower it.  A flag can only be raised when it is down, so consecutive barriers cannot overtake each
// other.  System-scope fences order the peer stores of the kernels enqueued before this one (complete, by stream order)
// ahead of the flag, and the flag ahead of the reads of the kernels enqueued after it.  Spins are bounded (30 s, then trap).
__device__ __forceinline__ unsigned long long slab_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__global__ void slab_barrier_kernel(unsigned* const* __restrict__ pads, int rank, int size) {
  const int p = threadIdx.x;
  if (p >= size) return;
  __threadfence_system();
  unsigned* theirs = pads[p] + rank;   // flag "rank has arrived" in p's pad
  unsigned* mine = pads[rank] + p;     // flag "p has arrived" in my pad
  const unsigned long long t0 = slab_now_ns(), limit = 30ull * 1000000000ull;   // a rank that never arrives: trap, no hang
  while (atomicCAS_system(theirs, 0u, 1u) != 0u) {
    __nanosleep(200);
    if (slab_now_ns() - t0 > limit) __trap();
  }
  while (atomicCAS_system(mine, 1u, 0u) != 1u) {
    __nanosleep(200);
    if (slab_now_ns() - t0 > limit) __trap();
  }
  __threadfence_system();
}

}  // namespace jfx

extern "C" {

int jfx_slab_create(const jfx_plan_desc* desc, int sharding, jfx_slab** out) {
  using namespace jfx;
  JFX_REQUIRE(desc && out, JFX_ERR_INVALID, "null argument");
  *out = nullptr;
  JFX_REQUIRE(desc->abi_version == JFX_ABI_VERSION, JFX_ERR_INVALID, "ABI version %d != %d", desc->abi_version, JFX_ABI_VERSION);
  JFX_REQUIRE(sharding == JFX_SLAB_SPECTRAL || sharding == JFX_SLAB_PHYSICAL, JFX_ERR_INVALID, "bad sharding %d", sharding);
  JFX_REQUIRE(desc->ndim >= 2 && desc->ndim <= JFX_MAX_DIMS, JFX_ERR_INVALID, "slab transforms need 2..%d axes", JFX_MAX_DIMS);
  const int P = desc->slab_size, rank = desc->slab_rank;
  JFX_REQUIRE(P >= 1 && rank >= 0 && rank < P, JFX_ERR_INVALID, "slab rank %d / size %d", rank, P);
  std::unique_ptr<jfx_slab> sl(new (std::nothrow) jfx_slab);
  JFX_REQUIRE(sl, JFX_ERR_NOMEM, "out of host memory");
  sl->rank = rank; sl->size = P; sl->sharding = sharding; sl->ndim = desc->ndim; sl->dtype = desc->dtype;
  const int sh = sharding == JFX_SLAB_SPECTRAL ? 0 : 1;   // sharded axis of the input; the other of {0, 1} is split
  const int split = 1 - sh;
  // phase 1: every axis but the sharded one
  jfx_plan_desc d1 = *desc;
  d1.slab_rank = 0; d1.slab_size = 1;
  d1.axis[sh] = jfx_axis_desc{};
  d1.axis[sh].basis = JFX_BASIS_NONE;
  int rc = jfx_plan_create(&d1, &sl->phase1);
  if (rc != JFX_OK) return rc;
  for (int i = 0; i < desc->ndim; ++i) sl->mid[i] = sl->exch[i] = sl->phase1->shape_out[i];
  JFX_REQUIRE(sl->mid[split] % P == 0, JFX_ERR_INVALID, "split axis %d has extent %lld, not divisible by %d devices", split,
              (long long)sl->mid[split], P);   // sharding.py:59-63
  sl->exch[split] = sl->mid[split] / P;
  sl->exch[sh] = sl->mid[sh] * P;
  // phase 2: the sharded axis alone, on the exchanged block
  jfx_plan_desc d2 = *desc;
  d2.slab_rank = 0; d2.slab_size = 1;
  for (int i = 0; i < desc->ndim; ++i) {
    d2.shape_in[i] = sl->exch[i];
    if (i != sh) { d2.axis[i] = jfx_axis_desc{}; d2.axis[i].basis = JFX_BASIS_NONE; }
  }
  rc = jfx_plan_create(&d2, &sl->phase2);
  if (rc != JFX_OK) return rc;
  const size_t es = dtype_size(desc->dtype);
  sl->mid_bytes = align_up((size_t)prod(sl->mid, 0, sl->ndim) * es, 256);
  sl->recv_bytes = (size_t)prod(sl->exch, 0, sl->ndim) * es;
  sl->fused = P > 1 && desc->ndim == 3 && jfx_plan_scatter_supported(sl->phase1, P, split) == 1;
  {
    static const bool no_fuse = [] { const char* e = getenv("JFX_SLAB_P2P"); return e && e[0] == '0'; }();
    if (no_fuse) sl->fused = false;
  }
  // workspace: [phase-1 result (copy route only)] [phase-1 workspace] [phase-2 workspace]
  sl->ws1_off = sl->fused ? 0 : sl->mid_bytes;
  sl->ws2_off = sl->ws1_off + align_up(sl->phase1->ws_bytes, 256);
  sl->ws_bytes = sl->ws2_off + align_up(sl->phase2->ws_bytes, 256);
  *out = sl.release();
  return JFX_OK;
}

void jfx_slab_destroy(jfx_slab* s) { delete s; }

int jfx_slab_sizes(const jfx_slab* s, size_t* recv_bytes, size_t* signal_bytes, size_t* workspace_bytes,
                   int64_t* shape_out /* [JFX_MAX_DIMS] */) {
  using namespace jfx;
  JFX_REQUIRE(s, JFX_ERR_INVALID, "null argument");
  if (recv_bytes) *recv_bytes = s->recv_bytes;
  if (signal_bytes) *signal_bytes = (size_t)s->size * sizeof(unsigned);
  if (workspace_bytes) *workspace_bytes = s->ws_bytes;
  if (shape_out) for (int i = 0; i < JFX_MAX_DIMS; ++i) shape_out[i] = i < s->ndim ? s->phase2->shape_out[i] : 1;
  return JFX_OK;
}

int jfx_slab_fused(const jfx_slab* s) { return s && s->fused ? 1 : 0; }

int jfx_slab_bind(jfx_slab* s, void* const* recv0, void* const* recv1, void* const* signal) {
  using namespace jfx;
  JFX_REQUIRE(s && recv0 && recv1 && signal, JFX_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(s->mu);
  const int P = s->size;
  for (int p = 0; p < P; ++p)
    JFX_REQUIRE(recv0[p] && recv1[p] && signal[p], JFX_ERR_INVALID, "peer %d: null buffer", p);
  s->recv[0].assign(recv0, recv0 + P);
  s->recv[1].assign(recv1, recv1 + P);
  s->signal.assign(signal, signal + P);
  if (!s->d_signal) JFX_CUDA_OK(cudaMalloc((void**)&s->d_signal, sizeof(void*) * P));
  JFX_CUDA_OK(cudaMemcpy(s->d_signal, s->signal.data(), sizeof(void*) * P, cudaMemcpyHostToDevice));
  JFX_CUDA_OK(cudaDeviceSynchronize());   // binding may synchronise (like plan creation); execution never does
  s->turn = 0;
  return JFX_OK;
}

int jfx_slab_execute(jfx_slab* s, void* stream, const void* in, void* out, void* workspace) {
  using namespace jfx;
  JFX_REQUIRE(s && in && out, JFX_ERR_INVALID, "null argument");
  JFX_REQUIRE(s->ws_bytes == 0 || workspace, JFX_ERR_INVALID, "slab transform needs a %zu byte workspace", s->ws_bytes);
  std::lock_guard<std::mutex> lk(s->mu);
  const int P = s->size, rank = s->rank;
  JFX_REQUIRE((int)s->signal.size() == P, JFX_ERR_INVALID, "jfx_slab_bind has not been called");
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
  const int t = s->turn;
  s->turn ^= 1;
  const int sh = s->sharding == JFX_SLAB_SPECTRAL ? 0 : 1, split = 1 - sh;
  int rc;
  if (s->fused) {
    rc = jfx_execute_scatter(s->phase1, stream, in, s->recv[t].data(), P, rank, split, ws + s->ws1_off);
    if (rc != JFX_OK) return rc;
  } else {
    // phase 1 into the workspace (or straight through when it has no pass), then one strided 2-D copy per peer:
    // the exchange of lax.all_to_all(split_axis, concat_axis, tiled=True) (sharding.py:83-89) without pack / unpack kernels
    const void* mid = in;
    if (!s->phase1->passes.empty()) {
      rc = jfx_execute(s->phase1, stream, in, ws, ws + s->ws1_off);
      if (rc != JFX_OK) return rc;
      mid = ws;
    }
    const size_t es = dtype_size(s->dtype);
    const size_t rest = (size_t)prod(s->mid, 2, s->ndim) * es;      // bytes of one (axis 0, axis 1) element
    const int64_t s0 = s->mid[0], s1 = s->mid[1];
    for (int q = 0; q < P; ++q) {
      const int p = (rank + q) % P;                                   // start with myself, then round the ring
      if (sh == 0) {
        // spectral -> physical: peer p receives mid[:, p*s1/P:(p+1)*s1/P] as block `rank` of its [P, s0, s1/P, ...] buffer
        const size_t row = (size_t)(s1 / P) * rest;
        JFX_CUDA_OK(cudaMemcpy2DAsync((char*)s->recv[t][p] + (size_t)rank * s0 * row, row,
                                      (const char*)mid + (size_t)p * row, (size_t)s1 * rest, row, (size_t)s0,
                                      cudaMemcpyDeviceToDevice, st));
      } else {
        // physical -> spectral: peer p receives mid[p*s0/P:(p+1)*s0/P] as columns rank*s1.. of its [s0/P, P*s1, ...] buffer
        const size_t row = (size_t)s1 * rest;
        JFX_CUDA_OK(cudaMemcpy2DAsync((char*)s->recv[t][p] + (size_t)rank * row, (size_t)P * row,
                                      (const char*)mid + (size_t)p * (s0 / P) * row, row, row, (size_t)(s0 / P),
                                      cudaMemcpyDeviceToDevice, st));
      }
    }
  }
  if (P > 1) {
    slab_barrier_kernel<<<1, 32 * ((P + 31) / 32), 0, st>>>((unsigned* const*)s->d_signal, rank, P);
    JFX_CUDA_OK(cudaGetLastError());
  }
  return jfx_execute(s->phase2, stream, s->recv[t][rank], out, ws + s->ws2_off);
}

int jfx_calibrate_dmma(void* stream, int iters, double* tflops) {
  using namespace jfx;
  JFX_REQUIRE(tflops && iters > 0, JFX_ERR_INVALID, "bad argument");
  return calibrate_dmma((cudaStream_t)stream, iters, tflops);
}
int jfx_calibrate_dfma(void* stream, int iters, double* tflops) {
  using namespace jfx;
  JFX_REQUIRE(tflops && iters > 0, JFX_ERR_INVALID, "bad argument");
  return calibrate_dfma((cudaStream_t)stream, iters, tflops);
}

}  // extern "C"
