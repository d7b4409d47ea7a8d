// Elementwise / layout kernels around the transforms:
//   * pointwise nonlinear-term evaluation (integrators/nonlinear.py:135-217) as a small stack
//     machine over the backward_primitive leaves,
//   * diagonal stage arithmetic of ETDRK4 / RK4 / IMEX (integrators/etdrk4.py:152-166),
//   * slab pack / unpack around the all-to-all (sharding.py:83-89).
// All are HBM-bound streaming kernels: 16-byte accesses, grid = multiple of 148 SMs.
#include <cuda_runtime.h>
#include <math.h>

#include "jfx_common.h"

namespace jfx {

// ---------------------------------------------------------------------------------------------
// complex helpers
// ---------------------------------------------------------------------------------------------
template <typename T> struct C2 { T re, im; };

template <typename T> __device__ __forceinline__ C2<T> cmul(C2<T> a, C2<T> b) {
  return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
template <typename T> __device__ __forceinline__ C2<T> cdiv(C2<T> a, C2<T> b) {
  T d = b.re * b.re + b.im * b.im;
  return {(a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d};
}
template <typename T> __device__ __forceinline__ C2<T> cexp(C2<T> a) {
  T e = exp(a.re), s, c;
  sincos(a.im, &s, &c);
  return {e * c, e * s};
}
template <typename T> __device__ __forceinline__ C2<T> clog(C2<T> a) {
  return {log(hypot(a.re, a.im)), atan2(a.im, a.re)};
}
template <typename T> __device__ __forceinline__ C2<T> cpowi(C2<T> a, int n) {
  bool neg = n < 0;
  unsigned m = neg ? (unsigned)(-n) : (unsigned)n;
  C2<T> r{T(1), T(0)};
  while (m) {
    if (m & 1u) r = cmul(r, a);
    a = cmul(a, a);
    m >>= 1;
  }
  if (neg) r = cdiv(C2<T>{T(1), T(0)}, r);
  return r;
}
template <typename T> __device__ __forceinline__ C2<T> csqrt_(C2<T> a) {
  if (a.im == T(0)) {
    if (a.re >= T(0)) return {sqrt(a.re), T(0)};
    return {T(0), sqrt(-a.re)};
  }
  T m = hypot(a.re, a.im);
  T sr = sqrt((m + a.re) * T(0.5));
  T si = sqrt((m - a.re) * T(0.5));
  return {sr, a.im < 0 ? -si : si};
}

template <typename T> __device__ C2<T> apply_func(int fn, C2<T> v) {
  const bool real_arg = (v.im == T(0));
  switch (fn) {
    case JFX_FN_EXP: return cexp(v);
    case JFX_FN_LOG: return clog(v);
    case JFX_FN_SIN: {
      if (real_arg) return {sin(v.re), T(0)};
      return {sin(v.re) * cosh(v.im), cos(v.re) * sinh(v.im)};
    }
    case JFX_FN_COS: {
      if (real_arg) return {cos(v.re), T(0)};
      return {cos(v.re) * cosh(v.im), -sin(v.re) * sinh(v.im)};
    }
    case JFX_FN_TAN: {
      if (real_arg) return {tan(v.re), T(0)};
      C2<T> s{sin(v.re) * cosh(v.im), cos(v.re) * sinh(v.im)};
      C2<T> c{cos(v.re) * cosh(v.im), -sin(v.re) * sinh(v.im)};
      return cdiv(s, c);
    }
    case JFX_FN_SINH: {
      if (real_arg) return {sinh(v.re), T(0)};
      return {sinh(v.re) * cos(v.im), cosh(v.re) * sin(v.im)};
    }
    case JFX_FN_COSH: {
      if (real_arg) return {cosh(v.re), T(0)};
      return {cosh(v.re) * cos(v.im), sinh(v.re) * sin(v.im)};
    }
    case JFX_FN_TANH: {
      if (real_arg) return {tanh(v.re), T(0)};
      C2<T> s{sinh(v.re) * cos(v.im), cosh(v.re) * sin(v.im)};
      C2<T> c{cosh(v.re) * cos(v.im), sinh(v.re) * sin(v.im)};
      return cdiv(s, c);
    }
    case JFX_FN_SQRT: return csqrt_(v);
    case JFX_FN_SIGN: {
      if (real_arg) return {T((v.re > 0) - (v.re < 0)), T(0)};
      T m = hypot(v.re, v.im);
      return m == T(0) ? C2<T>{T(0), T(0)} : C2<T>{v.re / m, v.im / m};
    }
    case JFX_FN_HEAVISIDE: return {v.re > 0 ? T(1) : (v.re < 0 ? T(0) : T(0.5)), T(0)};
    case JFX_FN_ASIN: return {asin(v.re), T(0)};
    case JFX_FN_ACOS: return {acos(v.re), T(0)};
    case JFX_FN_ATAN: return {atan(v.re), T(0)};
    case JFX_FN_ASINH: return {asinh(v.re), T(0)};
    case JFX_FN_ACOSH: return {acosh(v.re), T(0)};
    case JFX_FN_ATANH: return {atanh(v.re), T(0)};
    case JFX_FN_RE: return {v.re, T(0)};
    case JFX_FN_IM: return {v.im, T(0)};
  }
  return v;
}

// ---------------------------------------------------------------------------------------------
// pointwise stack machine
// ---------------------------------------------------------------------------------------------
struct PwArgs {
  const void* leaves[JFX_MAX_LEAVES];
  const void* statics[JFX_MAX_LEAVES];
  int n_instr;
  jfx_pw_instr instr[JFX_MAX_PROGRAM];
  double consts[32][2];
};

template <typename T, bool CPLX>
__global__ void __launch_bounds__(256) pointwise_kernel(const __grid_constant__ PwArgs a, void* out_,
                                                        int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    C2<T> st[8];
    int sp = 0;
    for (int pc = 0; pc < a.n_instr; ++pc) {
      const int op = a.instr[pc].op, arg = a.instr[pc].arg;
      switch (op) {
        case JFX_PW_LEAF:
          if (CPLX) st[sp++] = reinterpret_cast<const C2<T>*>(a.leaves[arg])[i];
          else st[sp++] = {reinterpret_cast<const T*>(a.leaves[arg])[i], T(0)};
          break;
        case JFX_PW_STATIC:
          if (CPLX) st[sp++] = reinterpret_cast<const C2<T>*>(a.statics[arg])[i];
          else st[sp++] = {reinterpret_cast<const T*>(a.statics[arg])[i], T(0)};
          break;
        case JFX_PW_CONST: st[sp++] = {T(a.consts[arg][0]), T(a.consts[arg][1])}; break;
        case JFX_PW_ADD: --sp; st[sp - 1] = {st[sp - 1].re + st[sp].re, st[sp - 1].im + st[sp].im}; break;
        case JFX_PW_MUL: --sp; st[sp - 1] = cmul(st[sp - 1], st[sp]); break;
        case JFX_PW_POWI: st[sp - 1] = cpowi(st[sp - 1], arg); break;
        case JFX_PW_ABS: st[sp - 1] = {CPLX ? hypot(st[sp - 1].re, st[sp - 1].im) : fabs(st[sp - 1].re), T(0)}; break;
        case JFX_PW_NEG: st[sp - 1] = {-st[sp - 1].re, -st[sp - 1].im}; break;
        case JFX_PW_CONJ: st[sp - 1].im = -st[sp - 1].im; break;
        case JFX_PW_FUNC: st[sp - 1] = apply_func<T>(arg, st[sp - 1]); break;
        case JFX_PW_POWR: {
          const T e = T(a.consts[arg][0]);
          C2<T> v = st[sp - 1];
          if (v.im == T(0) && (v.re >= T(0) || !CPLX)) st[sp - 1] = {pow(v.re, e), T(0)};
          else {
            C2<T> l = clog(v);
            st[sp - 1] = cexp(C2<T>{l.re * e, l.im * e});
          }
        } break;
      }
    }
    if (CPLX) reinterpret_cast<C2<T>*>(out_)[i] = st[0];
    else reinterpret_cast<T*>(out_)[i] = st[0].re;
  }
}

int launch_pointwise(cudaStream_t s, const PointwiseProgram& prog, const void* const* leaves,
                     const void* const* statics, void* out, int64_t n, int dtype) {
  JFX_REQUIRE(prog.n_instr > 0 && prog.n_instr <= JFX_MAX_PROGRAM, JFX_ERR_INVALID, "bad program length");
  // validate stack discipline on the host so the kernel cannot run off its 8-entry stack
  int sp = 0;
  for (int i = 0; i < prog.n_instr; ++i) {
    const int op = prog.instr[i].op, arg = prog.instr[i].arg;
    if (op == JFX_PW_LEAF || op == JFX_PW_CONST || op == JFX_PW_STATIC) {
      if (op == JFX_PW_LEAF) JFX_REQUIRE(arg >= 0 && arg < prog.n_leaves, JFX_ERR_INVALID, "leaf index %d", arg);
      if (op == JFX_PW_CONST) JFX_REQUIRE(arg >= 0 && arg < prog.n_consts, JFX_ERR_INVALID, "const index %d", arg);
      if (op == JFX_PW_STATIC) JFX_REQUIRE(statics && arg >= 0 && arg < JFX_MAX_LEAVES && statics[arg], JFX_ERR_INVALID, "static index %d", arg);
      ++sp;
    } else if (op == JFX_PW_ADD || op == JFX_PW_MUL) {
      JFX_REQUIRE(sp >= 2, JFX_ERR_INVALID, "stack underflow at instr %d", i);
      --sp;
    } else {
      JFX_REQUIRE(sp >= 1, JFX_ERR_INVALID, "stack underflow at instr %d", i);
      if (op == JFX_PW_POWR) JFX_REQUIRE(arg >= 0 && arg < prog.n_consts, JFX_ERR_INVALID, "const index %d", arg);
    }
    JFX_REQUIRE(sp <= 8, JFX_ERR_UNSUPPORTED, "pointwise stack deeper than 8");
  }
  JFX_REQUIRE(sp == 1, JFX_ERR_INVALID, "program leaves %d values on the stack", sp);
  if (n == 0) return JFX_OK;
  PwArgs a{};
  for (int i = 0; i < prog.n_leaves; ++i) a.leaves[i] = leaves[i];
  for (int i = 0; i < JFX_MAX_LEAVES; ++i) a.statics[i] = statics ? statics[i] : nullptr;
  a.n_instr = prog.n_instr;
  memcpy(a.instr, prog.instr, sizeof(jfx_pw_instr) * prog.n_instr);
  memcpy(a.consts, prog.consts, sizeof(a.consts));
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  switch (dtype) {
    case JFX_F32: pointwise_kernel<float, false><<<(unsigned)blocks, 256, 0, s>>>(a, out, n); break;
    case JFX_F64: pointwise_kernel<double, false><<<(unsigned)blocks, 256, 0, s>>>(a, out, n); break;
    case JFX_C64: pointwise_kernel<float, true><<<(unsigned)blocks, 256, 0, s>>>(a, out, n); break;
    case JFX_C128: pointwise_kernel<double, true><<<(unsigned)blocks, 256, 0, s>>>(a, out, n); break;
    default: set_error("bad dtype"); return JFX_ERR_INVALID;
  }
  JFX_CUDA_OK(cudaGetLastError());
  return JFX_OK;
}

// ---------------------------------------------------------------------------------------------
// diagonal stage arithmetic: out = sum_t alpha_t * coeff_t (.) x_t
// ---------------------------------------------------------------------------------------------
struct AxArgs {
  const void* coeff[8];
  const void* x[8];
  double alpha[8];
  int n_terms;
};

template <typename T, bool CPLX, bool CCOEF>
__global__ void __launch_bounds__(256) axpby_kernel(const __grid_constant__ AxArgs a, void* out_, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    C2<T> acc{T(0), T(0)};
#pragma unroll 1
    for (int t = 0; t < a.n_terms; ++t) {
      C2<T> xv;
      if (CPLX) xv = reinterpret_cast<const C2<T>*>(a.x[t])[i];
      else xv = {reinterpret_cast<const T*>(a.x[t])[i], T(0)};
      C2<T> cv{T(a.alpha[t]), T(0)};
      if (a.coeff[t]) {
        if (CCOEF) {
          C2<T> c = reinterpret_cast<const C2<T>*>(a.coeff[t])[i];
          cv = {c.re * cv.re, c.im * cv.re};
        } else {
          cv.re *= reinterpret_cast<const T*>(a.coeff[t])[i];
        }
      }
      C2<T> pr = cmul(cv, xv);
      acc.re += pr.re;
      acc.im += pr.im;
    }
    if (CPLX) reinterpret_cast<C2<T>*>(out_)[i] = acc;
    else reinterpret_cast<T*>(out_)[i] = acc.re;
  }
}

int launch_axpby_diag(cudaStream_t s, int n_terms, const void* const* coeff, const double* alpha,
                      const void* const* x, void* out, int64_t n, int dtype, int coeff_is_complex) {
  JFX_REQUIRE(n_terms >= 1 && n_terms <= 8, JFX_ERR_INVALID, "n_terms must be 1..8");
  JFX_REQUIRE(!(coeff_is_complex && !dtype_is_complex(dtype)), JFX_ERR_INVALID, "complex coefficients on real data");
  if (n == 0) return JFX_OK;
  AxArgs a{};
  a.n_terms = n_terms;
  for (int t = 0; t < n_terms; ++t) {
    a.coeff[t] = coeff ? coeff[t] : nullptr;
    a.x[t] = x[t];
    a.alpha[t] = alpha ? alpha[t] : 1.0;
  }
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  const unsigned b = (unsigned)blocks;
  switch (dtype) {
    case JFX_F32: axpby_kernel<float, false, false><<<b, 256, 0, s>>>(a, out, n); break;
    case JFX_F64: axpby_kernel<double, false, false><<<b, 256, 0, s>>>(a, out, n); break;
    case JFX_C64:
      if (coeff_is_complex) axpby_kernel<float, true, true><<<b, 256, 0, s>>>(a, out, n);
      else axpby_kernel<float, true, false><<<b, 256, 0, s>>>(a, out, n);
      break;
    case JFX_C128:
      if (coeff_is_complex) axpby_kernel<double, true, true><<<b, 256, 0, s>>>(a, out, n);
      else axpby_kernel<double, true, false><<<b, 256, 0, s>>>(a, out, n);
      break;
    default: set_error("bad dtype"); return JFX_ERR_INVALID;
  }
  JFX_CUDA_OK(cudaGetLastError());
  return JFX_OK;
}

// ---------------------------------------------------------------------------------------------
// slab pack / unpack.  Arrays are viewed as [pre, len, post] around the split / concat axis.
// pack:   out[p][pre][len/P][post] = in[pre][p*len/P + j][post]
// unpack: out[pre][p*len/P + j][post] = in[p][pre][len/P][post]
// Element = 8 bytes (f64) or 16 bytes (c128) etc.; copied as raw words.
// ---------------------------------------------------------------------------------------------
template <typename W, bool PACK>
__global__ void __launch_bounds__(256) slab_kernel(const W* __restrict__ in, W* __restrict__ out,
                                                   int64_t pre, int64_t len, int64_t post, int parts) {
  const int64_t blk = len / parts;
  const int64_t total = pre * len * post;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    // idx enumerates the blocked layout [p][pre][blk][post]
    const int64_t w = idx % post;
    const int64_t j = (idx / post) % blk;
    const int64_t a = (idx / (post * blk)) % pre;
    const int64_t p = idx / (post * blk * pre);
    const int64_t full = (a * len + p * blk + j) * post + w;
    if (PACK) out[idx] = in[full];
    else out[full] = in[idx];
  }
}

template <bool PACK>
static int run_slab(cudaStream_t s, const void* in, void* out, const int64_t* shape, int ndim, int axis,
                    int parts, int dtype) {
  JFX_REQUIRE(ndim >= 1 && ndim <= JFX_MAX_DIMS && axis >= 0 && axis < ndim, JFX_ERR_INVALID, "bad axis");
  JFX_REQUIRE(parts >= 1 && shape[axis] % parts == 0, JFX_ERR_INVALID,
              "axis length %lld not divisible by %d devices", (long long)shape[axis], parts);
  int64_t pre = 1, post = 1;
  for (int i = 0; i < axis; ++i) pre *= shape[i];
  for (int i = axis + 1; i < ndim; ++i) post *= shape[i];
  const int64_t len = shape[axis];
  const int64_t total = pre * len * post;
  if (total == 0) return JFX_OK;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  const size_t es = dtype_size(dtype);
  if (es == 16) {
    slab_kernel<double2, PACK><<<(unsigned)blocks, 256, 0, s>>>((const double2*)in, (double2*)out, pre, len, post, parts);
  } else if (es == 8) {
    slab_kernel<double, PACK><<<(unsigned)blocks, 256, 0, s>>>((const double*)in, (double*)out, pre, len, post, parts);
  } else {
    slab_kernel<float, PACK><<<(unsigned)blocks, 256, 0, s>>>((const float*)in, (float*)out, pre, len, post, parts);
  }
  JFX_CUDA_OK(cudaGetLastError());
  return JFX_OK;
}

int launch_slab_pack(cudaStream_t s, const void* in, void* out, const int64_t* shape, int ndim,
                     int split_axis, int parts, int dtype) {
  return run_slab<true>(s, in, out, shape, ndim, split_axis, parts, dtype);
}
int launch_slab_unpack(cudaStream_t s, const void* in, void* out, const int64_t* shape_out, int ndim,
                       int concat_axis, int parts, int dtype) {
  return run_slab<false>(s, in, out, shape_out, ndim, concat_axis, parts, dtype);
}

}  // namespace jfx
