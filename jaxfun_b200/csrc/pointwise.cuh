// Pointwise stack machine of the nonlinear terms (integrators/nonlinear.py:135-217), shared by the
// standalone pointwise kernel (kernels_pointwise.cu) and the row-fused nonlinear kernel (kernels_fused.cu).
#pragma once
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
  PolyProgram poly;   // n_terms > 0: evaluate the polynomial normal form instead of the stack program
};


// Evaluate the program at one point.  leaf(i) / stat(i) return the value of leaf / static i there.
// The operand stack is a SHIFT stack of DEPTH registers (top = st[0]): every access has a compile-time
// index, so it lives in registers instead of local memory.  DEPTH = 4 covers every nonlinear term of the
// reference's examples; programs_depth() (host) picks 4 or 8.
template <typename T, bool CPLX, int DEPTH, typename LeafFn, typename StatFn>
__device__ __forceinline__ C2<T> pw_eval(const jfx_pw_instr* __restrict__ instr, int n_instr,
                                         const double (*__restrict__ consts)[2], LeafFn leaf, StatFn stat) {
  C2<T> st[DEPTH];
#pragma unroll
  for (int i = 0; i < DEPTH; ++i) st[i] = C2<T>{T(0), T(0)};
  auto push = [&](C2<T> x) {
#pragma unroll
    for (int i = DEPTH - 1; i > 0; --i) st[i] = st[i - 1];
    st[0] = x;
  };
  auto pop2 = [&](C2<T> r) {   // replace the two top entries by r
#pragma unroll
    for (int i = 1; i < DEPTH - 1; ++i) st[i] = st[i + 1];
    st[0] = r;
  };
  for (int pc = 0; pc < n_instr; ++pc) {
    const int op = instr[pc].op, arg = instr[pc].arg;
    switch (op) {
      case JFX_PW_LEAF: push(leaf(arg)); break;
      case JFX_PW_STATIC: push(stat(arg)); break;
      case JFX_PW_CONST: push(C2<T>{T(consts[arg][0]), T(consts[arg][1])}); break;
      case JFX_PW_ADD: pop2(C2<T>{st[1].re + st[0].re, st[1].im + st[0].im}); break;
      case JFX_PW_MUL: pop2(cmul(st[1], st[0])); break;
      case JFX_PW_POWI:
        if (arg == 2) st[0] = cmul(st[0], st[0]);
        else st[0] = cpowi(st[0], arg);
        break;
      case JFX_PW_ABS: st[0] = {CPLX ? hypot(st[0].re, st[0].im) : fabs(st[0].re), T(0)}; break;
      case JFX_PW_NEG: st[0] = {-st[0].re, -st[0].im}; break;
      case JFX_PW_CONJ: st[0].im = -st[0].im; break;
      case JFX_PW_FUNC: st[0] = apply_func<T>(arg, st[0]); break;
      case JFX_PW_POWR: {
        const T e = T(consts[arg][0]);
        C2<T> v = st[0];
        if (v.im == T(0) && (v.re >= T(0) || !CPLX)) st[0] = {pow(v.re, e), T(0)};
        else {
          C2<T> l = clog(v);
          st[0] = cexp(C2<T>{l.re * e, l.im * e});
        }
      } break;
    }
  }
  return st[0];
}

// Polynomial normal form: acc += coeff * prod(factors), three live values, no operand stack.
template <typename T, typename LeafFn>
__device__ __forceinline__ C2<T> poly_eval(const PolyProgram& P, LeafFn leaf) {
  C2<T> acc{T(0), T(0)};
  for (int t = 0; t < P.n_terms; ++t) {
    const PolyTerm& tm = P.t[t];
    C2<T> p{T(tm.cre), T(tm.cim)};
    int prev = -1;
    C2<T> x{T(0), T(0)};
    for (int f = 0; f < tm.nf; ++f) {
      const int fc = tm.fac[f];
      if (fc != prev) {
        x = leaf(fc & 0x7f);
        if (fc & 0x80) x.im = -x.im;
        prev = fc;
      }
      p = cmul(p, x);
    }
    acc.re += p.re;
    acc.im += p.im;
  }
  return acc;
}

// The same for C points at once: the C leaf loads of a factor are independent, so their latencies overlap
// (the leaves of the row-fused kernel come from an L2-resident scratch).  leaf(l, c) = leaf l at point c.
template <typename T, int C, typename LeafFn>
__device__ __forceinline__ void poly_eval_vec(const PolyProgram& P, LeafFn leaf, C2<T>* out) {
#pragma unroll
  for (int c = 0; c < C; ++c) out[c] = C2<T>{T(0), T(0)};
  for (int t = 0; t < P.n_terms; ++t) {
    const PolyTerm& tm = P.t[t];
    C2<T> p[C], x[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { p[c] = C2<T>{T(tm.cre), T(tm.cim)}; x[c] = C2<T>{T(0), T(0)}; }
    int prev = -1;
    for (int f = 0; f < tm.nf; ++f) {
      const int fc = tm.fac[f];
      if (fc != prev) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
          x[c] = leaf(fc & 0x7f, c);
          if (fc & 0x80) x[c].im = -x[c].im;
        }
        prev = fc;
      }
#pragma unroll
      for (int c = 0; c < C; ++c) p[c] = cmul(p[c], x[c]);
    }
#pragma unroll
    for (int c = 0; c < C; ++c) { out[c].re += p[c].re; out[c].im += p[c].im; }
  }
}

}  // namespace jfx
