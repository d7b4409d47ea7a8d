// Elementwise / layout kernels around the transforms:
//   * pointwise nonlinear-term evaluation (integrators/nonlinear.py:135-217) as a small stack
//     machine over the backward_primitive leaves,
//   * diagonal stage arithmetic of ETDRK4 / RK4 / IMEX (integrators/etdrk4.py:152-166),
//   * slab pack / unpack around the all-to-all (sharding.py:83-89).
// All are HBM-bound streaming kernels: 16-byte accesses, grid = multiple of 148 SMs.
#include <cuda_runtime.h>
#include <math.h>

#include <algorithm>
#include <complex>
#include <map>
#include <vector>

#include "jfx_common.h"
#include "pointwise.cuh"

namespace jfx {

template <typename T, bool CPLX, int DEPTH>
__global__ void __launch_bounds__(256) pointwise_kernel(const __grid_constant__ PwArgs a, void* out_,
                                                        int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    auto leaf = [&](int l) -> C2<T> {
      if (CPLX) return reinterpret_cast<const C2<T>*>(a.leaves[l])[i];
      return C2<T>{reinterpret_cast<const T*>(a.leaves[l])[i], T(0)};
    };
    auto stat = [&](int l) -> C2<T> {
      if (CPLX) return reinterpret_cast<const C2<T>*>(a.statics[l])[i];
      return C2<T>{reinterpret_cast<const T*>(a.statics[l])[i], T(0)};
    };
    const C2<T> r = a.poly.n_terms > 0 ? poly_eval<T>(a.poly, leaf)
                                       : pw_eval<T, CPLX, DEPTH>(a.instr, a.n_instr, a.consts, leaf, stat);
    if (CPLX) reinterpret_cast<C2<T>*>(out_)[i] = r;
    else reinterpret_cast<T*>(out_)[i] = r.re;
  }
}

int program_depth(const PointwiseProgram& prog) {
  int sp = 0, mx = 0;
  for (int i = 0; i < prog.n_instr; ++i) {
    const int op = prog.instr[i].op;
    if (op == JFX_PW_LEAF || op == JFX_PW_CONST || op == JFX_PW_STATIC) ++sp;
    else if (op == JFX_PW_ADD || op == JFX_PW_MUL) --sp;
    mx = sp > mx ? sp : mx;
  }
  return mx;
}

// postfix program -> polynomial normal form by symbolic execution (false: not a polynomial in the leaves)
bool program_to_poly(const PointwiseProgram& prog, PolyProgram* out) {
  using Mono = std::vector<unsigned char>;                 // sorted factor list
  using Poly = std::map<Mono, std::complex<double>>;
  out->n_terms = 0;
  std::vector<Poly> st;
  auto mul = [](const Poly& a, const Poly& b, Poly* r) -> bool {
    r->clear();
    for (const auto& x : a)
      for (const auto& y : b) {
        Mono m = x.first;
        m.insert(m.end(), y.first.begin(), y.first.end());
        if (m.size() > JFX_POLY_MAX_FACTORS) return false;
        std::sort(m.begin(), m.end());
        (*r)[m] += x.second * y.second;
      }
    return r->size() <= 4 * JFX_POLY_MAX_TERMS;
  };
  for (int i = 0; i < prog.n_instr; ++i) {
    const int op = prog.instr[i].op, arg = prog.instr[i].arg;
    switch (op) {
      case JFX_PW_LEAF: {
        if (arg < 0 || arg >= 0x80) return false;
        Poly p; p[Mono{(unsigned char)arg}] = 1.0; st.push_back(p);
      } break;
      case JFX_PW_CONST: {
        Poly p; p[Mono{}] = std::complex<double>(prog.consts[arg][0], prog.consts[arg][1]); st.push_back(p);
      } break;
      case JFX_PW_ADD: {
        if (st.size() < 2) return false;
        Poly b = st.back(); st.pop_back();
        for (const auto& y : b) st.back()[y.first] += y.second;
      } break;
      case JFX_PW_MUL: {
        if (st.size() < 2) return false;
        Poly b = st.back(); st.pop_back();
        Poly r;
        if (!mul(st.back(), b, &r)) return false;
        st.back() = r;
      } break;
      case JFX_PW_NEG:
        if (st.empty()) return false;
        for (auto& y : st.back()) y.second = -y.second;
        break;
      case JFX_PW_CONJ: {
        if (st.empty()) return false;
        Poly r;
        for (const auto& y : st.back()) {
          Mono m = y.first;
          for (auto& f : m) f ^= 0x80;
          std::sort(m.begin(), m.end());
          r[m] += std::conj(y.second);
        }
        st.back() = r;
      } break;
      case JFX_PW_ABS: {
        // |z|^(2k) = (z conj z)^k: only in front of an even integer power
        if (st.empty() || i + 1 >= prog.n_instr || prog.instr[i + 1].op != JFX_PW_POWI) return false;
        const int n = prog.instr[i + 1].arg;
        if (n < 2 || (n & 1) || n > 8) return false;
        Poly c;
        for (const auto& y : st.back()) {
          Mono m = y.first;
          for (auto& f : m) f ^= 0x80;
          std::sort(m.begin(), m.end());
          c[m] += std::conj(y.second);
        }
        Poly zz, r;
        if (!mul(st.back(), c, &zz)) return false;
        r = zz;
        for (int k = 1; k < n / 2; ++k) { Poly tmp; if (!mul(r, zz, &tmp)) return false; r = tmp; }
        st.back() = r;
        ++i;   // the POWI was consumed
      } break;
      case JFX_PW_POWI: {
        if (st.empty() || arg < 0 || arg > 8) return false;
        Poly r; r[Mono{}] = 1.0;
        for (int k = 0; k < arg; ++k) { Poly tmp; if (!mul(r, st.back(), &tmp)) return false; r = tmp; }
        st.back() = r;
      } break;
      default: return false;   // statics, functions, real powers: stack machine
    }
  }
  if (st.size() != 1) return false;
  int n = 0;
  for (const auto& y : st.back()) {
    if (y.second == std::complex<double>(0.0, 0.0)) continue;
    if (n >= JFX_POLY_MAX_TERMS) return false;
    PolyTerm& t = out->t[n++];
    t.cre = y.second.real(); t.cim = y.second.imag();
    t.nf = (int)y.first.size();
    for (int f = 0; f < JFX_POLY_MAX_FACTORS; ++f) t.fac[f] = f < t.nf ? y.first[f] : 0;
  }
  if (n == 0) { out->t[0] = PolyTerm{}; n = 1; }   // identically zero: one zero term
  out->n_terms = n;
  return true;
}

int validate_program(const PointwiseProgram& prog, const void* const* statics) {
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
  return JFX_OK;
}

int launch_pointwise(cudaStream_t s, const PointwiseProgram& prog, const void* const* leaves,
                     const void* const* statics, void* out, int64_t n, int dtype) {
  JFX_REQUIRE(prog.n_instr > 0 && prog.n_instr <= JFX_MAX_PROGRAM, JFX_ERR_INVALID, "bad program length");
  { const int rc = validate_program(prog, statics); if (rc != JFX_OK) return rc; }
  if (n == 0) return JFX_OK;
  PwArgs a{};
  for (int i = 0; i < prog.n_leaves; ++i) a.leaves[i] = leaves[i];
  for (int i = 0; i < JFX_MAX_LEAVES; ++i) a.statics[i] = statics ? statics[i] : nullptr;
  a.n_instr = prog.n_instr;
  memcpy(a.instr, prog.instr, sizeof(jfx_pw_instr) * prog.n_instr);
  memcpy(a.consts, prog.consts, sizeof(a.consts));
  if (!program_to_poly(prog, &a.poly)) a.poly.n_terms = 0;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  const bool deep = program_depth(prog) > 4;
#define JFX_PW_LAUNCH(TT, CP) \
  do { \
    if (deep) pointwise_kernel<TT, CP, 8><<<(unsigned)blocks, 256, 0, s>>>(a, out, n); \
    else pointwise_kernel<TT, CP, 4><<<(unsigned)blocks, 256, 0, s>>>(a, out, n); \
  } while (0)
  switch (dtype) {
    case JFX_F32: JFX_PW_LAUNCH(float, false); break;
    case JFX_F64: JFX_PW_LAUNCH(double, false); break;
    case JFX_C64: JFX_PW_LAUNCH(float, true); break;
    case JFX_C128: JFX_PW_LAUNCH(double, true); break;
    default: set_error("bad dtype"); return JFX_ERR_INVALID;
  }
#undef JFX_PW_LAUNCH
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
// Per-point contraction of scattered evaluation (TensorProductSpace.evaluate, tensorproductspace.py:263-273 of the
// reference: einsum("i,j,ij") per point).  After the last axis has been contracted with the basis values of all points
// (a table pass), every remaining axis is reduced with weights that differ per point:
//     out[o, p] = sum_j y[o, j, p] * w[p, j]          y: [outer, n, P] data dtype,  w: [P, n] real (double / float)
// One thread per (o, p): consecutive threads read consecutive p of y (coalesced); w[p, :] stays in L1 / L2.
// ---------------------------------------------------------------------------------------------
template <typename T, bool CPLX, bool WCPLX>
__global__ void __launch_bounds__(256) point_contract_kernel(const void* __restrict__ y_, const void* __restrict__ w_,
                                                             void* __restrict__ out_, int64_t outer, int n, int64_t P) {
  const int64_t total = outer * P;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t o = idx / P, p = idx - o * P;
    if (CPLX) {
      const C2<T>* y = reinterpret_cast<const C2<T>*>(y_) + o * (int64_t)n * P + p;
      T re = 0, im = 0;
      for (int j = 0; j < n; ++j) {
        const C2<T> v = y[(int64_t)j * P];
        if (WCPLX) {
          const C2<T> ww = reinterpret_cast<const C2<T>*>(w_)[p * n + j];
          re += v.re * ww.re - v.im * ww.im;
          im += v.re * ww.im + v.im * ww.re;
        } else {
          const T ww = reinterpret_cast<const T*>(w_)[p * n + j];
          re += v.re * ww;
          im += v.im * ww;
        }
      }
      reinterpret_cast<C2<T>*>(out_)[idx] = C2<T>{re, im};
    } else {
      const T* y = reinterpret_cast<const T*>(y_) + o * (int64_t)n * P + p;
      const T* w = reinterpret_cast<const T*>(w_);
      T acc = 0;
      for (int j = 0; j < n; ++j) acc += y[(int64_t)j * P] * w[p * n + j];
      reinterpret_cast<T*>(out_)[idx] = acc;
    }
  }
}

int launch_point_contract(cudaStream_t s, const void* y, const void* w, void* out, int64_t outer, int n, int64_t P, int dtype,
                          int w_is_complex) {
  JFX_REQUIRE(outer >= 0 && n >= 0 && P >= 0, JFX_ERR_INVALID, "negative extent");
  JFX_REQUIRE(!(w_is_complex && !dtype_is_complex(dtype)), JFX_ERR_INVALID, "complex weights on real data");
  if (outer * P == 0) return JFX_OK;
  int64_t blocks = (outer * P + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  const unsigned b = (unsigned)blocks;
  switch (dtype) {
    case JFX_F32: point_contract_kernel<float, false, false><<<b, 256, 0, s>>>(y, w, out, outer, n, P); break;
    case JFX_F64: point_contract_kernel<double, false, false><<<b, 256, 0, s>>>(y, w, out, outer, n, P); break;
    case JFX_C64:
      if (w_is_complex) point_contract_kernel<float, true, true><<<b, 256, 0, s>>>(y, w, out, outer, n, P);
      else point_contract_kernel<float, true, false><<<b, 256, 0, s>>>(y, w, out, outer, n, P);
      break;
    case JFX_C128:
      if (w_is_complex) point_contract_kernel<double, true, true><<<b, 256, 0, s>>>(y, w, out, outer, n, P);
      else point_contract_kernel<double, true, false><<<b, 256, 0, s>>>(y, w, out, outer, n, P);
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
