// Stand-alone GPU check of the parity-folded contraction (no Python, starts in milliseconds):
//   nvcc -O2 -std=c++17 -o tools/fold_check tools/fold_check.cu -Ljaxfun_b200 -ljfx -Xlinker -rpath -Xlinker '$ORIGIN/../jaxfun_b200'
// For a matrix of shapes it creates the same JFX_OP_APPLY plan twice through the C ABI — JFX_DMMA_FOLD=0 (plain
// dgemm_dmma_tma) and JFX_DMMA_FOLD=1 (folded) — with Legendre Vandermonde tables built here, compares the results
// and times both at the bench size.  Exit code 0 = every case agrees to 1e-12.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../include/jfx.h"

static void gauss_legendre(int n, std::vector<double>& x, std::vector<double>& w) {
  x.resize(n); w.resize(n);
  for (int i = 0; i < (n + 1) / 2; ++i) {
    double z = std::cos(M_PI * (i + 0.75) / (n + 0.5)), pp = 0;
    for (int it = 0; it < 100; ++it) {
      double p1 = 1, p2 = 0;
      for (int j = 0; j < n; ++j) { const double p3 = p2; p2 = p1; p1 = ((2.0 * j + 1) * z * p2 - j * p3) / (j + 1); }
      pp = n * (z * p1 - p2) / (z * z - 1);
      const double dz = p1 / pp;
      z -= dz;
      if (std::fabs(dz) < 1e-16) break;
    }
    x[i] = -z; x[n - 1 - i] = z;
    w[i] = w[n - 1 - i] = 2 / ((1 - z * z) * pp * pp);
  }
}
// backward table [nq][N]: P_k(x_j); forward table [N][nq]: w_j P_k(x_j) (2k+1)/2
static void tables(int N, int nq, std::vector<double>& B, std::vector<double>& F) {
  std::vector<double> x, w;
  gauss_legendre(nq, x, w);
  B.assign((size_t)nq * N, 0); F.assign((size_t)N * nq, 0);
  for (int j = 0; j < nq; ++j) {
    double p0 = 1, p1 = x[j];
    for (int k = 0; k < N; ++k) {
      const double pk = k == 0 ? p0 : (k == 1 ? p1 : 0);
      double v = pk;
      if (k >= 2) { v = ((2.0 * k - 1) * x[j] * p1 - (k - 1) * p0) / k; p0 = p1; p1 = v; }
      B[(size_t)j * N + k] = v;
      F[(size_t)k * nq + j] = w[j] * v * (2 * k + 1) / 2;
    }
  }
}

struct Case { int ndim; long long shape[4]; int tax[4]; int N, nq; bool fwd; int dtype; };

static int g_fail = 0;
static FILE* g_log = nullptr;
#define LOG(...) do { printf(__VA_ARGS__); if (g_log) { fprintf(g_log, __VA_ARGS__); fflush(g_log); } } while (0)

static jfx_plan* make_plan(const Case& c, const std::vector<double>& T, bool fold) {
  setenv("JFX_DMMA_FOLD", fold ? "1" : "0", 1);
  jfx_plan_desc d;
  memset(&d, 0, sizeof(d));
  d.abi_version = JFX_ABI_VERSION; d.op = JFX_OP_APPLY; d.dtype = c.dtype; d.ndim = c.ndim; d.slab_size = 1;
  for (int i = 0; i < c.ndim; ++i) {
    d.shape_in[i] = c.shape[i];
    if (c.tax[i]) {
      d.axis[i].basis = JFX_BASIS_TABLE; d.axis[i].table = T.data();
      d.axis[i].table_rows = c.fwd ? c.N : c.nq; d.axis[i].table_cols = c.fwd ? c.nq : c.N;
      d.axis[i].n_modes = c.N; d.axis[i].n_quad = c.nq; d.axis[i].domain_factor = 1.0;
    }
  }
  jfx_plan* p = nullptr;
  if (jfx_plan_create(&d, &p) != 0) { LOG("plan_create failed: %s\n", jfx_last_error()); return nullptr; }
  return p;
}

static double run_case(const Case& c, int time_iters) {
  std::vector<double> B, F;
  tables(c.N, c.nq, B, F);
  const std::vector<double>& T = c.fwd ? F : B;
  jfx_plan* p0 = make_plan(c, T, false);
  jfx_plan* p1 = make_plan(c, T, true);
  if (!p0 || !p1) { ++g_fail; return -1; }
  long long so[JFX_MAX_DIMS];
  jfx_plan_shape_out(p0, (int64_t*)so);
  const int comp = c.dtype == JFX_C128 ? 2 : 1;
  size_t nin = comp, nout = comp;
  for (int i = 0; i < c.ndim; ++i) { nin *= c.shape[i]; nout *= so[i]; }
  size_t ws0 = 0, ws1 = 0;
  jfx_plan_workspace_bytes(p0, &ws0); jfx_plan_workspace_bytes(p1, &ws1);
  std::vector<double> h(nin);
  unsigned long long sd = 88172645463325252ull;
  for (auto& v : h) { sd ^= sd << 13; sd ^= sd >> 7; sd ^= sd << 17; v = (double)(sd >> 11) / 9007199254740992.0 - 0.5; }
  double *din, *o0, *o1; void* w = nullptr;
  cudaMalloc(&din, nin * 8); cudaMalloc(&o0, nout * 8); cudaMalloc(&o1, nout * 8);
  if (ws0 > ws1) ws1 = ws0;
  if (ws1) cudaMalloc(&w, ws1);
  cudaMemcpy(din, h.data(), nin * 8, cudaMemcpyHostToDevice);
  cudaMemset(o0, 0xff, nout * 8); cudaMemset(o1, 0xff, nout * 8);
  int rc0 = jfx_execute(p0, nullptr, din, o0, w);
  cudaError_t e0 = cudaDeviceSynchronize();
  int rc1 = jfx_execute(p1, nullptr, din, o1, w);
  cudaError_t e1 = cudaDeviceSynchronize();
  double err = -1, t0 = 0, t1 = 0;
  if (rc0 || rc1 || e0 != cudaSuccess || e1 != cudaSuccess) {
    LOG("FAIL execute rc %d %d cuda %s / %s : %s\n", rc0, rc1, cudaGetErrorString(e0), cudaGetErrorString(e1), jfx_last_error());
    ++g_fail;
    if (e0 != cudaSuccess || e1 != cudaSuccess) { LOG("device error: stopping\n"); if (g_log) fclose(g_log); exit(2); }
  } else {
    std::vector<double> a(nout), b(nout);
    cudaMemcpy(a.data(), o0, nout * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(b.data(), o1, nout * 8, cudaMemcpyDeviceToHost);
    double mx = 0, df = 0;
    bool nan = false;
    for (size_t i = 0; i < nout; ++i) {
      if (!(std::fabs(b[i]) < 1e300)) nan = true;
      mx = std::fmax(mx, std::fabs(a[i])); df = std::fmax(df, std::fabs(a[i] - b[i]));
    }
    err = nan ? 1e9 : df / (mx > 0 ? mx : 1);
    if (time_iters > 0) {
      cudaEvent_t ev0, ev1; cudaEventCreate(&ev0); cudaEventCreate(&ev1);
      float ms;
      for (int which = 0; which < 2; ++which) {
        jfx_plan* p = which ? p1 : p0; double* o = which ? o1 : o0;
        for (int i = 0; i < 3; ++i) jfx_execute(p, nullptr, din, o, w);
        cudaEventRecord(ev0);
        for (int i = 0; i < time_iters; ++i) jfx_execute(p, nullptr, din, o, w);
        cudaEventRecord(ev1); cudaEventSynchronize(ev1);
        cudaEventElapsedTime(&ms, ev0, ev1);
        (which ? t1 : t0) = ms / time_iters;
      }
    }
    const bool ok = err < 1e-12;
    if (!ok) ++g_fail;
    LOG("%s %s %s ndim %d shape", ok ? "ok  " : "FAIL", c.fwd ? "fwd" : "bwd", c.dtype == JFX_C128 ? "c128" : "f64 ", c.ndim);
    for (int i = 0; i < c.ndim; ++i) LOG(" %lld%s", c.shape[i], c.tax[i] ? "*" : "");
    LOG("  N %d nq %d  rel diff %.2e", c.N, c.nq, err);
    if (time_iters > 0) LOG("  plain %.4f ms  folded %.4f ms  (x%.2f)", t0, t1, t0 / t1);
    LOG("\n");
  }
  cudaFree(din); cudaFree(o0); cudaFree(o1); if (w) cudaFree(w);
  jfx_plan_destroy(p0); jfx_plan_destroy(p1);
  return err;
}

int main(int argc, char** argv) {
  const bool quick = argc > 1 && !strcmp(argv[1], "--quick");
  g_log = fopen(quick ? "gpurun_out/fold_check_quick.txt" : "gpurun_out/fold_check.txt", "w");
  if (jfx_device_count() < 1) { LOG("no CUDA device\n"); return 3; }
  // timing first (the numbers matter most if the box time runs out): 256^3 backward / forward, f64
  for (int fwd = 0; fwd < 2; ++fwd) {
    Case c{3, {256, 256, 256, 0}, {1, 1, 1, 0}, 256, 256, fwd != 0, JFX_F64};
    run_case(c, 10);
  }
  const int sizes[] = {16, 32, 64, 96, 128, 192};
  for (int n : sizes)
    for (int fwd = 0; fwd < 2; ++fwd) {
      run_case(Case{1, {n, 0, 0, 0}, {1, 0, 0, 0}, n, n, fwd != 0, JFX_F64}, 0);                  // single line
      run_case(Case{2, {300, n, 0, 0}, {0, 1, 0, 0}, n, n, fwd != 0, JFX_F64}, 0);                // batched lines (NT)
      run_case(Case{2, {n, n, 0, 0}, {1, 1, 0, 0}, n, n, fwd != 0, JFX_F64}, 0);                  // NN + NT
      run_case(Case{3, {n, 6, n, 0}, {1, 0, 1, 0}, n, n, fwd != 0, JFX_F64}, 0);                  // middle batch axis
      run_case(Case{2, {n, 34, 0, 0}, {1, 0, 0, 0}, n, n, fwd != 0, JFX_C128}, 0);                // complex columns
      if (!quick) run_case(Case{3, {n, n, n, 0}, {1, 1, 1, 0}, n, n, fwd != 0, JFX_F64}, 0);
    }
  // padded quadrature (3/2 rule) and an odd number of modes
  for (int fwd = 0; fwd < 2; ++fwd) {
    run_case(Case{2, {fwd ? 96 : 64, fwd ? 96 : 64, 0, 0}, {1, 1, 0, 0}, 64, 96, fwd != 0, JFX_F64}, 0);
    run_case(Case{2, {fwd ? 64 : 62, 130, 0, 0}, {1, 0, 0, 0}, 62, 64, fwd != 0, JFX_F64}, 0);
  }
  run_case(Case{2, {64, 63, 0, 0}, {0, 1, 0, 0}, 63, 64, false, JFX_F64}, 0);   // odd mode count: not foldable (OUT), plain path
  run_case(Case{2, {64, 64, 0, 0}, {0, 1, 0, 0}, 63, 64, true, JFX_F64}, 0);    // IN fold with an odd mode count
  if (!quick) {
    run_case(Case{3, {512, 512, 512, 0}, {1, 1, 1, 0}, 512, 512, false, JFX_F64}, 5);
    run_case(Case{2, {65536, 1024, 0, 0}, {0, 1, 0, 0}, 1024, 1024, false, JFX_F64}, 5);
    run_case(Case{2, {65536, 1024, 0, 0}, {0, 1, 0, 0}, 1024, 1024, true, JFX_F64}, 5);
  }
  LOG(g_fail ? "FOLD CHECK: %d FAILURES\n" : "FOLD CHECK: ALL OK\n", g_fail);
  if (g_log) fclose(g_log);
  return g_fail ? 1 : 0;
}
