/*
 * jfx.h — C ABI of the B200-native tensor-product spectral transform engine.
 *
 * This is the drop-in boundary for ONE path of spectralDNS/jaxfun: the
 * OrthogonalSpace / TensorProductSpace forward, backward, scalar_product and
 * backward_primitive transforms plus the pseudo-spectral nonlinear-term
 * evaluation the integrators call every stage.  The reference has no FFI of its
 * own (it is 100 % Python on JAX); each entry point below names the reference
 * Python interface it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *  - plain C: pointers, sizes, POD structs; no C++/torch/XLA types.
 *  - every function returns 0 (JFX_OK) or a negative jfx_status; the message of
 *    the last failure on the calling thread is available via jfx_last_error().
 *  - arrays are dense, row-major (C order); complex = interleaved (re, im).
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *  - jfx_execute* never allocates, never synchronises the device and is safe to
 *    capture into a CUDA graph.  Plans are immutable after creation and may be
 *    executed concurrently from several host threads with distinct workspaces.
 *  - there is no CPU fallback: without a CUDA device every compute entry point
 *    fails with JFX_ERR_CUDA.
 */
#ifndef JFX_H_
#define JFX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JFX_ABI_VERSION 1
#define JFX_MAX_DIMS 4      /* up to 3 transformed axes + 1 leading batch axis */
#define JFX_MAX_LEAVES 8    /* backward_primitive leaves of one nonlinear term  */
#define JFX_MAX_PROGRAM 128 /* pointwise bytecode length                        */

typedef enum {
  JFX_OK = 0,
  JFX_ERR_INVALID = -1,     /* bad descriptor / argument                        */
  JFX_ERR_UNSUPPORTED = -2, /* valid request the engine cannot run              */
  JFX_ERR_CUDA = -3,        /* CUDA runtime / driver error (incl. no device)    */
  JFX_ERR_NOMEM = -4,
  JFX_ERR_COMM = -5         /* slab exchange failure                            */
} jfx_status;

typedef enum { JFX_F32 = 0, JFX_F64 = 1, JFX_C64 = 2, JFX_C128 = 3 } jfx_dtype;

/* Which reference method the plan reproduces. */
typedef enum {
  JFX_OP_FORWARD = 0,            /* OrthogonalSpace.forward       galerkin/orthogonal.py:256-262;
                                    TensorProductSpace.forward     galerkin/tensorproductspace.py:395-417 */
  JFX_OP_SCALAR_PRODUCT = 1,     /* OrthogonalSpace.scalar_product orthogonal.py:264-277;
                                    TensorProductSpace.scalar_product tensorproductspace.py:367-393       */
  JFX_OP_BACKWARD = 2,           /* OrthogonalSpace.backward      orthogonal.py:214-227;
                                    TensorProductSpace.backward    tensorproductspace.py:330-365          */
  JFX_OP_BACKWARD_PRIMITIVE = 3, /* backward_primitive            orthogonal.py:229-246;
                                    tensorproductspace.py:419-460                                         */
  JFX_OP_NONLINEAR = 4,          /* BaseIntegrator.nonlinear_rhs / nonlinear_rhs_scalar_product
                                    integrators/base.py:230-248 + integrators/nonlinear.py:89-217         */
  JFX_OP_APPLY = 5               /* generic per-axis table apply (to_orthogonal / evaluate_mesh(uniform) /
                                    eigenvector solves: tensorproductspace.py:462-504, orthogonal.py:180-199) */
} jfx_op;

/* 1-D basis family of one tensor axis. */
typedef enum {
  JFX_BASIS_NONE = 0,       /* axis is not transformed (batch axis)                                */
  JFX_BASIS_TABLE = 1,      /* dense real table supplied by the host: Legendre / Jacobi /
                               Ultraspherical Vandermonde contraction (orthogonal.py:277,
                               Jacobi.py:65-110), or any basis at a size with no fast kernel      */
  JFX_BASIS_CTABLE = 2,     /* dense complex table (Fourier at lengths without an FFT kernel)      */
  JFX_BASIS_CHEBYSHEV = 3,  /* DCT-II / DCT-III fast path     galerkin/Chebyshev.py:225-279        */
  JFX_BASIS_FOURIER = 4     /* c2c FFT fast path              galerkin/Fourier.py:126-180          */
} jfx_basis;

typedef struct {
  int32_t basis;         /* jfx_basis                                                            */
  int32_t n_modes;       /* N: number of spectral coefficients kept along this axis              */
  int32_t n_quad;        /* n: number of quadrature points (>= n_modes; > means padding)         */
  int32_t deriv;         /* derivative order k of backward_primitive along this axis (else 0)    */
  double domain_factor;  /* df = reference length / true length (orthogonal.py:343-354)          */
  /* JFX_BASIS_TABLE / CTABLE: host pointer to the row-major table [n_out, n_in] that maps the
     axis (n_in -> n_out).  All weights / norms / derivative factors are already folded in by the
     host (it is the constant the reference rebuilds inside every jitted call).  The engine copies
     it to the device at plan creation; the caller may free it afterwards.  NULL for fast bases.  */
  const void* table;
  int32_t table_rows;    /* n_out */
  int32_t table_cols;    /* n_in  */
} jfx_axis_desc;

/* Pointwise bytecode of a nonlinear term (integrators/nonlinear.py:135-217).  A tiny stack
   machine evaluated per quadrature point over the leaf values.                                 */
typedef enum {
  JFX_PW_LEAF = 0,   /* push leaf[arg]                       */
  JFX_PW_CONST = 1,  /* push consts[arg] (complex constant)  */
  JFX_PW_ADD = 2,
  JFX_PW_MUL = 3,
  JFX_PW_POWI = 4,   /* top <- top ** arg (integer, may be negative) */
  JFX_PW_ABS = 5,    /* |top| (real result)                  */
  JFX_PW_NEG = 6,
  JFX_PW_FUNC = 7,   /* top <- func[arg](top): see jfx_pw_func */
  JFX_PW_POWR = 8,   /* top <- top ** consts[arg].re         */
  JFX_PW_CONJ = 9,
  JFX_PW_STATIC = 10 /* push statics[arg][point] (mesh-sampled coefficient, nonlinear.py:219-242) */
} jfx_pw_opcode;

typedef enum {
  JFX_FN_EXP = 0, JFX_FN_LOG, JFX_FN_SIN, JFX_FN_COS, JFX_FN_TAN, JFX_FN_SINH, JFX_FN_COSH,
  JFX_FN_TANH, JFX_FN_SQRT, JFX_FN_SIGN, JFX_FN_HEAVISIDE, JFX_FN_ASIN, JFX_FN_ACOS, JFX_FN_ATAN,
  JFX_FN_ASINH, JFX_FN_ACOSH, JFX_FN_ATANH, JFX_FN_RE, JFX_FN_IM
} jfx_pw_func;

typedef struct { int32_t op; int32_t arg; } jfx_pw_instr;

typedef struct {
  int32_t abi_version;            /* JFX_ABI_VERSION                                             */
  int32_t op;                     /* jfx_op                                                      */
  int32_t dtype;                  /* jfx_dtype of input and output arrays                        */
  int32_t ndim;                   /* number of array axes, 1..JFX_MAX_DIMS                       */
  int64_t shape_in[JFX_MAX_DIMS]; /* input array shape                                           */
  jfx_axis_desc axis[JFX_MAX_DIMS];
  /* --- slab decomposition (sharding.py:43-105); size 1 = single device ------------------- */
  int32_t slab_rank;
  int32_t slab_size;
  int32_t reserved[14];
} jfx_plan_desc;

/* Nonlinear term  F(uh) = T( E( leaf_0(uh), leaf_1(uh), ... ) )  (integrators/base.py:230-248):
   every leaf is a BACKWARD / BACKWARD_PRIMITIVE plan on the same coefficient array, E the pointwise
   program over the leaves, T a FORWARD or SCALAR_PRODUCT plan on the physical array. */
typedef struct {
  int32_t abi_version;
  int32_t n_leaves;
  const jfx_plan_desc* leaves[JFX_MAX_LEAVES];
  const jfx_plan_desc* final_transform;
  int32_t n_program;
  jfx_pw_instr program[JFX_MAX_PROGRAM];
  int32_t n_consts;
  double consts[32][2];
  int32_t n_statics;               /* mesh-sampled static coefficients (device pointers, physical shape) */
  const void* statics[JFX_MAX_LEAVES];
  int32_t reserved[8];
} jfx_nonlinear_desc;

typedef struct jfx_plan jfx_plan;
typedef struct jfx_nonlinear jfx_nonlinear;

/* Library / device ------------------------------------------------------------------------ */
int jfx_abi_version(void);
const char* jfx_last_error(void);
/* Number of visible CUDA devices (0 on a CPU-only host; never fails). */
int jfx_device_count(void);
/* 1 when `basis` has a fast (FFT / DCT) kernel at transform length n for `dtype`, else 0. */
int jfx_fast_path_available(int basis, int n, int dtype);

/* Plans ------------------------------------------------------------------------------------ */
int jfx_plan_create(const jfx_plan_desc* desc, jfx_plan** out);
void jfx_plan_destroy(jfx_plan* plan);
int jfx_plan_ndim(const jfx_plan* plan);
int jfx_plan_shape_out(const jfx_plan* plan, int64_t* shape_out /* [JFX_MAX_DIMS] */);
int jfx_plan_workspace_bytes(const jfx_plan* plan, size_t* bytes);
/* Algorithmic work of one execution: flops of dense contractions and the compulsory
   bytes (input read once + output written once). */
int jfx_plan_work(const jfx_plan* plan, double* flops, double* bytes);
/* Flops the plan actually issues: table passes whose table has the mirror symmetry of a polynomial basis on
   symmetric nodes (T[n-1-j, k] = +-(-1)^k T[j, k]) run parity-folded at half the multiply-adds of
   jfx_plan_work's algorithmic count (set JFX_DMMA_FOLD=0 before plan creation to disable the folding). */
int jfx_plan_executed_flops(const jfx_plan* plan, double* flops);
/* Number of kernel launches one jfx_execute enqueues. */
int jfx_plan_launches(const jfx_plan* plan);

/* Plan registry: a SERIALISABLE handle for callers that cannot carry a pointer (XLA FFI attributes, compilation caches,
   one program running on several devices under shard_map).  jfx_registry_register deep-copies the descriptor and its host
   tables and returns a 64-bit key that depends only on the descriptor's contents (incl. the table values); it may be called
   at trace time, any number of times.  jfx_registry_acquire returns the plan of that key for the CURRENT CUDA device,
   creating it (device tables, fold analysis) on first use there — call it from an initialisation stage, not from a stream
   callback that must not synchronise.  Plans obtained this way are owned by the registry (never jfx_plan_destroy them). */
int jfx_registry_register(const jfx_plan_desc* desc, uint64_t* key_out);
int jfx_registry_acquire(uint64_t key, const jfx_plan** out);
void jfx_registry_clear(void);

/* Execution: device pointers, enqueued on `stream`, no host synchronisation.
   `workspace` must hold jfx_plan_workspace_bytes() bytes (may be NULL when that is 0).
   `in` is never written; `out` must not alias `in` or `workspace`. */
int jfx_execute(const jfx_plan* plan, void* stream, const void* in, void* out, void* workspace);

/* Slab exchange fused into the transform (sharding.py:83-103 of the reference: pack + lax.all_to_all(tiled=True) [+ the
   concatenation]): the plan's final pass writes its result rows straight into the receive buffers of all `parts` GPUs
   of the box.  peer_out[p] = device pointer, valid on THIS device (peer-mapped / symmetric memory over NVLink), of rank
   p's receive buffer.  For a 3-D result [s0, s1, s2]:
     split_axis 1 (spectral -> physical): rank p receives out[:, p*s1/P:(p+1)*s1/P, :] as block `rank` of its
                                          [P, s0, s1/P, s2] buffer  == the [P*s0, s1/P, s2] array phase 2 transforms;
     split_axis 0 (physical -> spectral): rank p receives out[p*s0/P:(p+1)*s0/P, :, :] as columns rank*s1 .. of its
                                          [s0/P, P*s1, s2] buffer   (the unpack is included).
   The caller orders the exchange with a barrier across the ranks after this call (and double-buffers the receive side).
   Supported when the final pass is a parity-folded fp64 table pass along the last axis (jfx_plan_scatter_supported);
   otherwise JFX_ERR_UNSUPPORTED and the caller uses jfx_execute + jfx_slab_pack/unpack + its own all-to-all. */
int jfx_plan_scatter_supported(const jfx_plan* plan, int parts, int split_axis);
int jfx_execute_scatter(const jfx_plan* plan, void* stream, const void* in, void* const* peer_out, int parts, int rank,
                        int split_axis, void* workspace);

/* Convenience for host callers (the reference-facing e2e path): copies `in` (host) to the
   device, runs the plan and copies the result back into `out` (host); synchronises `stream`.
   Device buffers are owned and cached by the plan (first call allocates). */
int jfx_execute_host(jfx_plan* plan, void* stream, const void* in_host, void* out_host);

/* Pinned host buffers for the host-pointer path (cudaHostAlloc / cudaFreeHost). */
int jfx_host_alloc(void** ptr, size_t bytes);
int jfx_host_free(void* ptr);

/* Nonlinear-term evaluation: uh (coefficients, device) -> out (coefficients, device). */
int jfx_nonlinear_create(const jfx_nonlinear_desc* desc, jfx_nonlinear** out);
void jfx_nonlinear_destroy(jfx_nonlinear* nl);
int jfx_nonlinear_workspace_bytes(const jfx_nonlinear* nl, size_t* bytes);
int jfx_nonlinear_shape_out(const jfx_nonlinear* nl, int64_t* shape_out, int* ndim);
int jfx_nonlinear_launches(const jfx_nonlinear* nl);
int jfx_nonlinear_execute(const jfx_nonlinear* nl, void* stream, const void* uh, void* out,
                          void* workspace);

/* Standalone pointwise evaluation of a program over n points (leaves/statics: device arrays). */
int jfx_pointwise(void* stream, const jfx_pw_instr* program, int n_program, const double (*consts)[2],
                  int n_consts, const void* const* leaves, int n_leaves, const void* const* statics,
                  void* out, int64_t n, int dtype);

/* Slab exchange (sharding.py:83-89): local pack / unpack kernels around the all-to-all.
   block p of the send buffer = x[.., split slice p, ..] so that an all-to-all of equal blocks
   followed by concatenation along `concat_axis` reproduces lax.all_to_all(tiled=True). */
int jfx_slab_pack(void* stream, const void* in, void* out, const int64_t* shape, int ndim,
                  int split_axis, int parts, int dtype);
int jfx_slab_unpack(void* stream, const void* in, void* out, const int64_t* shape_out, int ndim,
                    int concat_axis, int parts, int dtype);

/* Slab-decomposed transform behind ONE call — `_apply_separable_spmd_shard_map` (sharding.py:43-105 of the reference) for
   this rank's local block:  phase 1 (the unsharded axes, locally) -> exchange (lax.all_to_all(split_axis = unsharded[0],
   concat_axis = sharded[0], tiled = True), sharding.py:83-89) -> phase 2 (the originally sharded axis).
   `desc` describes the whole transform on the LOCAL input block (shape_in = local block, one axis entry per axis, exactly as
   for a single-device plan) with slab_rank / slab_size set; `sharding` is the sharding of the input: JFX_SLAB_SPECTRAL =
   axis 0 sharded (backward / backward_primitive), JFX_SLAB_PHYSICAL = axis 1 sharded (forward / scalar_product); the result
   carries the other one (tests/galerkin/test_forward_backward_spmd.py:71-75).
   No communication library is involved: the exchange is written by the GPUs themselves into receive buffers that every rank
   maps from every peer (CUDA IPC / VMM / symmetric memory over NVLink) — by the contraction epilogue of phase 1 where that
   exists (jfx_execute_scatter), by one strided 2-D peer copy per rank otherwise (any dtype, any basis; no pack / unpack
   kernels) — followed by a device-side barrier on peer-mapped flag words.  The caller provides, per rank, TWO receive
   buffers of `recv_bytes` and one zero-initialised flag pad of `signal_bytes`, and binds the peer-mapped pointers of all
   ranks (index = rank) once.  jfx_slab_execute enqueues everything on `stream` and never synchronises the host; every rank
   of the box must call it the same number of times (it contains a barrier).  `in` is not written; `out` = local block of
   the result (shape from jfx_slab_sizes). */
#define JFX_SLAB_SPECTRAL 0
#define JFX_SLAB_PHYSICAL 1
typedef struct jfx_slab jfx_slab;
int jfx_slab_create(const jfx_plan_desc* desc, int sharding, jfx_slab** out);
void jfx_slab_destroy(jfx_slab* slab);
int jfx_slab_sizes(const jfx_slab* slab, size_t* recv_bytes, size_t* signal_bytes, size_t* workspace_bytes,
                   int64_t* shape_out /* [JFX_MAX_DIMS] */);
/* 1 when phase 1 ends in the peer-store epilogue, 0 when the exchange is done by peer copies. */
int jfx_slab_fused(const jfx_slab* slab);
int jfx_slab_bind(jfx_slab* slab, void* const* recv0 /* [size] */, void* const* recv1 /* [size] */,
                  void* const* signal /* [size] */);
int jfx_slab_execute(jfx_slab* slab, void* stream, const void* in, void* out, void* workspace);

/* Stage arithmetic of the integrators (etdrk4.py:152-166, rk4.py:14-20): out = sum_i c_i * x_i with
   per-element diagonal coefficients c_i (or NULL = 1) scaled by alpha_i.  Complex or real. */
int jfx_axpby_diag(void* stream, int n_terms, const void* const* coeff, const double* alpha,
                   const void* const* x, void* out, int64_t n, int dtype, int coeff_is_complex);

/* Scattered-point evaluation (TensorProductSpace.evaluate, galerkin/tensorproductspace.py:263-321: einsum "i,j,ij" per point).
   The last axis is contracted with the basis values of all points by an ordinary JFX_OP_APPLY plan; every remaining axis is
   then reduced with per-point weights:   out[o, p] = sum_j y[o, j, p] * w[p, j]
   y: [outer, n, points] (dtype), w: [points, n] values of the axis' basis functions at the points — real (float64 for
   f64 / c128, float32 for f32 / c64), or with w_is_complex != 0 the array dtype itself (Fourier axes) —, out: [outer, points].
   Device pointers. */
int jfx_point_contract(void* stream, const void* y, const void* w, void* out, int64_t outer, int32_t n, int64_t points,
                       int dtype, int w_is_complex);

/* Wavenumber-batched banded solves of Fourier x polynomial tensor-product systems — `tpmats_wavenumber_factor` /
   `TPMatricesWavenumberSolver.solve` (la/tpmatrix.py:1236-1354, 686-1014 of the reference) with the no-pivot banded LU of
   la/diamatrix.py:1937-1973.  System s (one per combination of Fourier wavenumbers, C order over the Fourier axes) has the
   banded matrix  B_s = sum_t weights[t][s] * P_t  on the polynomial axis (tpmatrix.py:1306-1347: W[t, s] = scale_t * product of
   the Fourier-axis diagonals at s; P_t = polynomial-axis matrix of term t in DIA form on the union `offsets`).
   jfx_banded_create assembles all B_s on the device, factors them in place (L unit lower, U upper; no pivoting) and
   fails with JFX_ERR_UNSUPPORTED when a pivot is zero or not finite (diamatrix.py:461-471 raises there).  It may synchronise.
   jfx_banded_solve solves every system for one right-hand-side array addressed as [outer, n, inner] (row-major; system
   s = o * inner + i, outer * inner = n_sys), i.e. the polynomial axis may be any axis of the array and nothing is
   transposed; `rhs` and `out` are device pointers of the descriptor's dtype and may be the same buffer.  It enqueues one
   launch on `stream`, allocates nothing and never synchronises. */
#define JFX_BANDED_MAX_TERMS 8
typedef struct {
  int32_t abi_version;   /* JFX_ABI_VERSION                                                                 */
  int32_t dtype;         /* jfx_dtype of the right-hand sides; the factors are kept in its precision        */
  int32_t band_complex;  /* 0: weights / diags are float64, 1: complex128 (needs a complex dtype)            */
  int32_t n_terms;       /* 1 .. JFX_BANDED_MAX_TERMS                                                       */
  int64_t n;             /* length of the polynomial axis = order of every banded system                    */
  int64_t n_sys;         /* number of systems = product of the Fourier extents                              */
  int32_t n_diags;       /* number of diagonals in `offsets`                                                */
  int32_t reserved0;
  const int32_t* offsets;/* host, [n_diags], strictly increasing, must contain 0                            */
  const void* weights;   /* host, [n_terms][n_sys]                                                          */
  const void* diags;     /* host, [n_terms][n_diags][n], column-aligned DIA: diags[t][d][j] = P_t[j - offsets[d], j] */
  int32_t reserved[8];
} jfx_banded_desc;
typedef struct jfx_banded jfx_banded;
int jfx_banded_create(const jfx_banded_desc* desc, jfx_banded** out);
void jfx_banded_destroy(jfx_banded* b);
/* lower / upper bandwidth and the bytes of device memory the factors occupy */
int jfx_banded_info(const jfx_banded* b, int32_t* p, int32_t* q, size_t* factor_bytes);
/* copy the factored band to the host: [p + q + 1][n][n_sys] elements (row p - s = multipliers of sub-diagonal s, rows
   p .. p + q = U; band[p + off][j] = entry (j - off, j)), real or complex of the dtype's precision.  Synchronises. */
int jfx_banded_factors(const jfx_banded* b, void* lu_host);
int jfx_banded_solve(const jfx_banded* b, void* stream, const void* rhs, void* out, int64_t outer, int64_t inner);

/* Calibration helpers used by bench.py (not on the product path). */
int jfx_calibrate_dmma(void* stream, int iters, double* tflops);
int jfx_calibrate_dfma(void* stream, int iters, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* JFX_H_ */
