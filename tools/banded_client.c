/* A plain C client of the banded-solver entry points of include/jfx.h (no Python, no torch): what a host written in C would
 * do to solve the per-wavenumber systems of a Fourier x polynomial Helmholtz problem on the GPU.
 *
 *   gcc -O2 -std=c99 -Iinclude -I/usr/local/cuda/include tools/banded_client.c -Ljaxfun_b200 -ljfx \
 *       -L/usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,$PWD/jaxfun_b200 -o /tmp/banded_client
 *
 * Builds n_sys tridiagonal systems B_s = W[0][s] * P_0 + W[1][s] * P_1 of order n from their separable form, lets the
 * library assemble + factor them on the device, solves one complex right-hand-side array laid out [n_sys][n] (polynomial axis
 * last: outer = n_sys, inner = 1) and checks the residual on the host.  Exit codes: 0 ok, 1 wrong result, 3 no CUDA device
 * (the library has no CPU fallback: jfx_banded_create returns JFX_ERR_CUDA), 2 any other failure.
 * Test infrastructure: tests/test_banded_host.py builds it and checks the no-device behaviour on the CPU-only host. */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "jfx.h"

#define N 40
#define NSYS 24

int main(void) {
  static const int32_t offsets[3] = {-1, 0, 1};
  static double W[2][NSYS], P[2][3][N], rhs[NSYS][N][2], x[NSYS][N][2];
  for (int s = 0; s < NSYS; ++s) {
    W[0][s] = 1.0;
    W[1][s] = 0.25 * s * s;                       /* k^2-like weight of the second term */
  }
  for (int j = 0; j < N; ++j) {                   /* column-aligned DIA: diags[t][d][j] = P_t[j - off_d, j] */
    P[0][0][j] = (j < N - 1) ? -1.0 : 0.0;        /* entry (j + 1, j) */
    P[0][1][j] = 2.5 + 0.01 * j;
    P[0][2][j] = (j > 0) ? -1.0 : 0.0;            /* entry (j - 1, j) */
    P[1][0][j] = 0.0;
    P[1][1][j] = 1.0;
    P[1][2][j] = 0.0;
  }
  for (int s = 0; s < NSYS; ++s)
    for (int j = 0; j < N; ++j) {
      rhs[s][j][0] = sin(0.3 * j + s);
      rhs[s][j][1] = cos(0.7 * j - s);
    }

  jfx_banded_desc d = {0};
  d.abi_version = JFX_ABI_VERSION;
  d.dtype = JFX_C128;
  d.band_complex = 0;
  d.n_terms = 2;
  d.n = N;
  d.n_sys = NSYS;
  d.n_diags = 3;
  d.offsets = offsets;
  d.weights = W;
  d.diags = P;
  jfx_banded* b = NULL;
  int rc = jfx_banded_create(&d, &b);
  if (rc == JFX_ERR_CUDA) {
    printf("no CUDA device: %s\n", jfx_last_error());
    return 3;
  }
  if (rc != JFX_OK) {
    printf("jfx_banded_create failed (%d): %s\n", rc, jfx_last_error());
    return 2;
  }
  int32_t p = 0, q = 0;
  size_t bytes = 0;
  jfx_banded_info(b, &p, &q, &bytes);
  void* dev = NULL;
  if (cudaMalloc(&dev, sizeof rhs) != cudaSuccess || cudaMemcpy(dev, rhs, sizeof rhs, cudaMemcpyHostToDevice) != cudaSuccess) return 2;
  rc = jfx_banded_solve(b, NULL, dev, dev, NSYS, 1);          /* in place, legacy default stream */
  if (rc != JFX_OK || cudaMemcpy(x, dev, sizeof x, cudaMemcpyDeviceToHost) != cudaSuccess) {
    printf("solve failed (%d): %s\n", rc, jfx_last_error());
    return 2;
  }
  cudaFree(dev);
  jfx_banded_destroy(b);
  double worst = 0.0;
  for (int s = 0; s < NSYS; ++s)
    for (int i = 0; i < N; ++i)
      for (int c = 0; c < 2; ++c) {
        double r = -rhs[s][i][c];
        for (int dgl = 0; dgl < 3; ++dgl) {
          const int j = i + offsets[dgl];                       /* entry (i, j) lives at column j of diagonal j - i */
          if (j < 0 || j >= N) continue;
          r += (W[0][s] * P[0][dgl][j] + W[1][s] * P[1][dgl][j]) * x[s][j][c];
        }
        if (fabs(r) > worst) worst = fabs(r);
      }
  printf("bandwidths p = %d q = %d, %zu bytes of factors, max residual %.3e\n", p, q, bytes, worst);
  return worst < 1e-12 ? 0 : 1;
}
