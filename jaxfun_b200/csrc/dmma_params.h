// Operand description shared by the FP64 tensor-core contraction kernels (kernels_dense.cu: cp.async
// loaders, kernels_dense_tma.cu: TMA loader).  Internal header.
#pragma once
#include <cstdint>

namespace jfx {
namespace dmma {

struct Params {
  const double* A;  // [M, K] row-major (lda)
  const double* B;  // NT: [N, K] row-major (ldb); NN: [K, N] row-major (ldb)
  double* C;        // [M, N] row-major (ldc)
  int M, N, K;
  int64_t lda, ldb, ldc;
  int64_t strideA, strideB, strideC;  // per batch
};

// TMA-fed persistent kernel.  1 = launched, 0 = outside its envelope (use the cp.async kernel), < 0 = error.
int launch_dmma_tma(cudaStream_t s, const Params& p, bool nn, int batch, int sms);

}  // namespace dmma
}  // namespace jfx
