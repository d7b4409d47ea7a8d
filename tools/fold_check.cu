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

// ---- extras (--extra): the two opt-in paths that have not run on a GPU yet --------------------------------------
// (a) CPLX_NT: complex rows on a last table axis, JFX_CPLX_NT=1 vs the default NN route
static void run_cplx_case(int rows, int N, int nq, bool fwd, int time_iters) {
  std::vector<double> B, F;
  tables(N, nq, B, F);
  const std::vector<double>& T = fwd ? F : B;
  Case c{2, {rows, fwd ? nq : N, 0, 0}, {0, 1, 0, 0}, N, nq, fwd, JFX_C128};
  setenv("JFX_CPLX_NT", "0", 1);
  jfx_plan* p0 = make_plan(c, T, true);
  setenv("JFX_CPLX_NT", "1", 1);
  jfx_plan* p1 = make_plan(c, T, true);
  setenv("JFX_CPLX_NT", "0", 1);
  if (!p0 || !p1) { ++g_fail; return; }
  const size_t nin = 2ull * rows * (fwd ? nq : N), nout = 2ull * rows * (fwd ? N : nq);
  std::vector<double> h(nin);
  unsigned long long sd = 0x2545F4914F6CDD1Dull;
  for (auto& v : h) { sd ^= sd << 13; sd ^= sd >> 7; sd ^= sd << 17; v = (double)(sd >> 11) / 9007199254740992.0 - 0.5; }
  double *din, *o0, *o1;
  cudaMalloc(&din, nin * 8); cudaMalloc(&o0, nout * 8); cudaMalloc(&o1, nout * 8);
  cudaMemcpy(din, h.data(), nin * 8, cudaMemcpyHostToDevice);
  cudaMemset(o1, 0xff, nout * 8);
  int rc0 = jfx_execute(p0, nullptr, din, o0, nullptr), rc1 = jfx_execute(p1, nullptr, din, o1, nullptr);
  cudaError_t e = cudaDeviceSynchronize();
  if (rc0 || rc1 || e != cudaSuccess) {
    LOG("FAIL cplx execute rc %d %d cuda %s : %s\n", rc0, rc1, cudaGetErrorString(e), jfx_last_error());
    ++g_fail;
    if (e != cudaSuccess) { if (g_log) fclose(g_log); exit(2); }
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
    const double err = nan ? 1e9 : df / (mx > 0 ? mx : 1);
    float t0 = 0, t1 = 0;
    if (time_iters > 0) {
      cudaEvent_t ev0, ev1; cudaEventCreate(&ev0); cudaEventCreate(&ev1);
      for (int which = 0; which < 2; ++which) {
        jfx_plan* p = which ? p1 : p0; double* o = which ? o1 : o0;
        jfx_execute(p, nullptr, din, o, nullptr);
        cudaEventRecord(ev0);
        for (int i = 0; i < time_iters; ++i) jfx_execute(p, nullptr, din, o, nullptr);
        cudaEventRecord(ev1); cudaEventSynchronize(ev1);
        cudaEventElapsedTime(which ? &t1 : &t0, ev0, ev1);
      }
    }
    const bool ok = err < 1e-12;
    if (!ok) ++g_fail;
    LOG("%s cplx-last-axis %s rows %d N %d nq %d  rel diff %.2e", ok ? "ok  " : "FAIL", fwd ? "fwd" : "bwd", rows, N, nq, err);
    if (time_iters > 0) LOG("  NN route %.4f ms  CPLX_NT %.4f ms  (x%.1f)", t0 / time_iters, t1 / time_iters, t0 / t1);
    LOG("\n");
  }
  cudaFree(din); cudaFree(o0); cudaFree(o1);
  jfx_plan_destroy(p0); jfx_plan_destroy(p1);
}

// (b) jfx_execute_scatter with P emulated ranks on this one GPU against jfx_execute + the exchange done on the host
static void run_scatter_case(int P, int n0, int n1, int n2, bool fwd) {
  // backward: local spectral block [n0/P][n1][n2], phase-1 axes (1, 2), split axis 1
  // forward : local physical block [n0][n1/P][n2], phase-1 axes (0, 2), split axis 0
  const int split = fwd ? 0 : 1;
  std::vector<double> B, F;
  // one table size per case keeps the tool short: cubic extents
  tables(n2, n2, B, F);
  const std::vector<double>& T = fwd ? F : B;
  Case c{3, {fwd ? n0 : n0 / P, fwd ? n1 / P : n1, n2, 0}, {fwd ? 1 : 0, fwd ? 0 : 1, 1, 0}, n2, n2, fwd, JFX_F64};
  jfx_plan* pl = make_plan(c, T, true);
  if (!pl) { ++g_fail; return; }
  if (!jfx_plan_scatter_supported(pl, P, split)) { LOG("FAIL scatter not supported P %d\n", P); ++g_fail; jfx_plan_destroy(pl); return; }
  const long long s0 = c.shape[0], s1 = c.shape[1], s2 = c.shape[2];
  const size_t nloc = (size_t)s0 * s1 * s2;
  size_t ws = 0;
  jfx_plan_workspace_bytes(pl, &ws);
  void* w = nullptr;
  if (ws) cudaMalloc(&w, ws);
  std::vector<std::vector<double>> hin(P, std::vector<double>(nloc)), hy(P, std::vector<double>(nloc));
  std::vector<double*> recv(P);
  for (int p = 0; p < P; ++p) { cudaMalloc(&recv[p], nloc * 8); cudaMemset(recv[p], 0xff, nloc * 8); }
  double *din, *dy;
  cudaMalloc(&din, nloc * 8); cudaMalloc(&dy, nloc * 8);
  unsigned long long sd = 0x9E3779B97F4A7C15ull;
  int rc = 0;
  for (int r = 0; r < P; ++r) {
    for (auto& v : hin[r]) { sd ^= sd << 13; sd ^= sd >> 7; sd ^= sd << 17; v = (double)(sd >> 11) / 9007199254740992.0 - 0.5; }
    cudaMemcpy(din, hin[r].data(), nloc * 8, cudaMemcpyHostToDevice);
    rc |= jfx_execute(pl, nullptr, din, dy, w);
    cudaMemcpy(hy[r].data(), dy, nloc * 8, cudaMemcpyDeviceToHost);
    rc |= jfx_execute_scatter(pl, nullptr, din, (void* const*)recv.data(), P, r, split, w);
    if (cudaDeviceSynchronize() != cudaSuccess || rc) {
      LOG("FAIL scatter execute rc %d : %s / %s\n", rc, jfx_last_error(), cudaGetErrorString(cudaGetLastError()));
      ++g_fail;
      if (g_log) fclose(g_log);
      exit(2);
    }
  }
  double err = 0, mx = 0;
  bool nan = false;
  std::vector<double> got(nloc);
  for (int p = 0; p < P; ++p) {
    cudaMemcpy(got.data(), recv[p], nloc * 8, cudaMemcpyDeviceToHost);
    for (int r = 0; r < P; ++r)
      for (long long a = 0; a < s0; ++a)
        for (long long b = 0; b < s1; ++b) {
          long long drow;
          if (!fwd) { const long long bp = s1 / P; if (b / bp != p) continue; drow = ((long long)r * s0 + a) * bp + b % bp; }
          else { const long long ap = s0 / P; if (a / ap != p) continue; drow = (a % ap) * (s1 * P) + (long long)r * s1 + b; }
          for (long long k = 0; k < s2; ++k) {
            const double want = hy[r][(a * s1 + b) * s2 + k], g = got[drow * s2 + k];
            if (!(std::fabs(g) < 1e300)) nan = true;
            err = std::fmax(err, std::fabs(want - g)); mx = std::fmax(mx, std::fabs(want));
          }
        }
  }
  const bool ok = !nan && err < 1e-13 * mx;
  if (!ok) ++g_fail;
  LOG("%s scatter %s P %d local %lld x %lld x %lld  max diff %.2e%s\n", ok ? "ok  " : "FAIL", fwd ? "fwd (split 0)" : "bwd (split 1)", P, s0, s1,
      s2, err, nan ? " (unwritten elements)" : "");
  for (int p = 0; p < P; ++p) cudaFree(recv[p]);
  cudaFree(din); cudaFree(dy); if (w) cudaFree(w);
  jfx_plan_destroy(pl);
}

int main(int argc, char** argv) {
  const bool quick = argc > 1 && !strcmp(argv[1], "--quick");
  if (argc > 1 && !strcmp(argv[1], "--extra")) {
    g_log = fopen("gpurun_out/fold_check_extra.txt", "w");
    if (jfx_device_count() < 1) { LOG("no CUDA device\n"); return 3; }
    for (int fwd = 0; fwd < 2; ++fwd) {
      run_cplx_case(300, 16, 16, fwd, 0);
      run_cplx_case(1, 64, 64, fwd, 0);
      run_cplx_case(777, 62, 64, fwd, 0);
      run_cplx_case(1000, 64, 96, fwd, 0);
      run_cplx_case(16384, 128, 128, fwd, 5);
      run_cplx_case(65536, 256, 256, fwd, 5);
    }
    for (int P : {2, 4, 8})
      for (int fwd = 0; fwd < 2; ++fwd) {
        // the two transformed axes share one table here, so their extents are equal: (1, 2) backward, (0, 2) forward
        if (!fwd) { run_scatter_case(P, 32, 64, 64, false); run_scatter_case(P, 16, 96, 96, false); }
        else { run_scatter_case(P, 64, 48, 64, true); run_scatter_case(P, 96, 16, 96, true); }
      }
    LOG(g_fail ? "FOLD CHECK EXTRA: %d FAILURES\n" : "FOLD CHECK EXTRA: ALL OK\n", g_fail);
    if (g_log) fclose(g_log);
    return g_fail ? 1 : 0;
  }
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
