// Plan-time / launch-time interface of the parity-folded contraction (kernels_dense_fold.cu).  Internal header.
#pragma once
#include <cuda_runtime.h>

namespace jfx {
namespace dmma {

struct FoldPlan;

// On by default for eligible table passes; JFX_DMMA_FOLD=0 (read at every plan creation) turns it off.
bool fold_enabled();
// Analyses the host table [rows][cols]; *out stays null when it has no mirror symmetry (not an error).
int fold_plan_create(const double* table, int rows, int cols, FoldPlan** out);
void fold_plan_destroy(FoldPlan* fp);
int fold_plan_type(const FoldPlan* fp);   // 0 none, 1 OUT (backward-like), 2 IN (forward-like)
// multiply-adds issued relative to the plain contraction: 1/2, plus the asymmetry-correction k-tiles of the highest modes
double fold_plan_flop_fraction(const FoldPlan* fp);
int fold_plan_corrected_modes(const FoldPlan* fp);
// The array is [outer][n_in][inner_real] doubles (complex data: inner_real = 2 * inner).
// 1 = launched, 0 = outside the envelope (use the plain kernel), < 0 = error.
int launch_dmma_fold(cudaStream_t s, const FoldPlan* fp, long long outer, long long inner_real, const double* in,
                     double* out);

// Last-axis (NT) folded pass whose result rows (a, b) = (row / B, row % B) are written straight into the slab-exchange
// receive buffers of `parts` GPUs (peer-mapped pointers): mode 1 splits b (spectral -> physical), mode 2 splits a
// (physical -> spectral, including the unpack).  1 = launched, 0 = not applicable, < 0 = error.
int launch_dmma_fold_scatter(cudaStream_t s, const FoldPlan* fp, long long outer, const double* in, int mode, int parts,
                             int src, int A, int B, double* const* peers);

// Complex interleaved data on a LAST table axis (any real table): one NT launch with (re, im) accumulator groups
// instead of the NN order with two real columns per batch.  Default; JFX_CPLX_NT=0 at plan creation keeps the NN route.
struct CplxPlan;
bool cplx_nt_enabled();
int cplx_plan_create(const double* table, int n_out, int n_in, CplxPlan** out);
void cplx_plan_destroy(CplxPlan* cp);
int launch_dmma_cplx_nt(cudaStream_t s, const CplxPlan* cp, long long outer, const double* in, double* out);

}  // namespace dmma
}  // namespace jfx
