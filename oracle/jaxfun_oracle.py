"""CPU ORACLE — test infrastructure only, never part of the product path.

A NumPy/SciPy/SymPy restatement of the reference algorithms for the transform hot path of
spectralDNS/jaxfun.  Every function names the reference lines it follows
(paths relative to the reference checkout, `src/jaxfun/...`).  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import
this module; nothing under `jaxfun_b200/` does.

Pinning status: PINNED against the reference's own source files.  jax is not installable here,
so the reference cannot run as shipped; instead its unmodified modules (galerkin/*.py,
utils/fastgl.py, galerkin/tensorproductspace.py, integrators/nonlinear.py) are executed on a
numpy stand-in for the jax API (tools/jaxshim) by `tests/golden/make_golden.py`, and the
resulting vectors (`tests/golden/reference_vectors.npz`: 1-D transforms of all six bases, padded /
derivative / complex-line variants, mixed-basis tensor products, nonlinear terms, FastGL node
tables) are replayed against this oracle in `tests/test_golden.py` (nodes / weights / wavenumbers
bit-exact, transforms < 1e-12; observed <= 3e-15).  What that does NOT pin is XLA's last-bit rounding
of cos / FFT / dot versus numpy / scipy ("bit level unpinned" — see DESIGN.md).
The widening rows are pinned less directly: Composite / DirectSum through the reference's own `get_bc_basis`
output (`tests/golden/reference_bc_basis.json`, made by tests/golden/make_golden_bc.py) and its examples' acceptance
criteria; `DirectSumTPS` is PARITY UNPINNED (the reference class needs its flax-based `la` package, which cannot be
mounted here) and is checked through the property that defines it — the prescribed boundary functions are reproduced.

Third-party arithmetic restated here: jax.numpy.fft / jax.scipy.fft.dct (jaxlib 0.11.0) ->
numpy.fft / scipy.fft; scipy.special.roots_jacobi (scipy 1.17.0 pinned, 1.18.1 here).
"""
from __future__ import annotations

import functools
import json
import os

import numpy as np
import scipy.fft
import sympy as sp
from scipy.special import roots_jacobi

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")   # the checker owns its copy of the FastGL constants
WORKERS = os.cpu_count() or 1  # scipy.fft threads (the CPU baseline uses every host core)
n_sym = sp.Symbol("n", integer=True)
alf, bet = sp.symbols("a,b", real=True)
delta = sp.KroneckerDelta


# =================================================================================================
# FastGL  (utils/fastgl.py)
# =================================================================================================
@functools.lru_cache(maxsize=1)
def _gl_data():
    tab = np.load(os.path.join(_DATA, "fastgl_tables.npz"))
    with open(os.path.join(_DATA, "fastgl_coeffs.json")) as f:
        co = {k: [float(s) for s in v] for k, v in json.load(f).items()}
    return tab["theta"], tab["weight"], tab["cl"], co


def besseljzero(k: int) -> float:
    """utils/fastgl.py:232-271."""
    co = _gl_data()[3]
    if k > 20:
        z = np.pi * (k - 0.25)
        r = 1.0 / z
        r2 = r * r
        c = co["kb20"]
        acc = c[7] + c[8] * r2
        for i in range(6, -1, -1):
            acc = c[i] + r2 * acc
        return z + r * acc
    return co["JZ"][k - 1]


def besselj1squared(k: int) -> float:
    """utils/fastgl.py:276-314."""
    co = _gl_data()[3]
    if k > 21:
        x = 1.0 / (k - 0.25)
        x2 = x * x
        c = co["km21"]
        acc = c[7] + c[8] * x2
        for i in range(6, 0, -1):
            acc = c[i] + x2 * acc
        return x * (c[0] + x2 * x2 * acc)
    return co["J1"][k - 1]


def _poly_desc(c, x):
    acc = c[0] * x + c[1]
    for v in c[2:]:
        acc = acc * x + v
    return acc


def GLPairS(n: int, k: int):
    """Asymptotic (theta, weight) pair, utils/fastgl.py:335-508."""
    co = _gl_data()[3]
    w = 1.0 / (n + 0.5)
    nu = besseljzero(k)
    theta = w * nu
    x = theta * theta
    B = besselj1squared(k)
    SF1T, SF2T, SF3T = (_poly_desc(co[s], x) for s in ("SF1T", "SF2T", "SF3T"))
    WSF1T, WSF2T, WSF3T = (_poly_desc(co[s], x) for s in ("WSF1T", "WSF2T", "WSF3T"))
    NuoSin = nu / np.sin(theta)
    BNuoSin = B * NuoSin
    WInvSinc = w * w * NuoSin
    WIS2 = WInvSinc * WInvSinc
    theta = w * (nu + theta * WInvSinc * (SF1T + WIS2 * (SF2T + WIS2 * SF3T)))
    Deno = BNuoSin + BNuoSin * WIS2 * (WSF1T + WIS2 * (WSF2T + WIS2 * WSF3T))
    return theta, (2.0 * w) / Deno


def GLPairTabulated(n: int, k: int):
    """utils/fastgl.py:512-544 (k is zero based here, as in the reference)."""
    TH, W, CL, _ = _gl_data()
    if n % 2 == 1:
        n2 = (n - 1) // 2
        if k == n2:
            return np.pi / 2, 2.0 / (CL[n] * CL[n])
        if k < n2:
            return TH[n, n2 - k - 1], W[n, n2 - k - 1]
        return np.pi - TH[n, k - n2 - 1], W[n, k - n2 - 1]
    n2 = n // 2
    if k < n2:
        return TH[n, n2 - k - 1], W[n, n2 - k - 1]
    return np.pi - TH[n, k - n2], W[n, k - n2]


def GLPair(n: int, k: int):
    """utils/fastgl.py:548-559."""
    if n < 101:
        return GLPairTabulated(n, k - 1)
    if 2 * k - 1 > n:
        t, w = GLPairS(n, n - k + 1)
        return np.pi - t, w
    return GLPairS(n, k)


@functools.lru_cache(maxsize=64)
def leggauss(N: int) -> np.ndarray:
    """utils/fastgl.py:562-567 -> array (2, N): x = cos(theta) ascending, weights."""
    pairs = [GLPair(N, N - i) for i in range(N)]
    return np.array([[np.cos(t) for t, _ in pairs], [w for _, w in pairs]])


# =================================================================================================
# helpers
# =================================================================================================
def _along(fn, x: np.ndarray, axis: int) -> np.ndarray:
    """Apply a function that works along axis 0 of a [n, m] array along `axis` of x."""
    x = np.moveaxis(np.asarray(x), axis, 0)
    shp = x.shape
    y = fn(x.reshape(shp[0], -1))
    return np.moveaxis(y.reshape((y.shape[0],) + shp[1:]), 0, axis)


def _lamb(expr, N):
    """sp.lambdify(n, expr, modules=...)(arange(N)) with scalar results broadcast (Jacobi.py:81-89)."""
    v = sp.lambdify(n_sym, expr, modules=["scipy", "numpy"])(np.arange(N))
    if np.ndim(v) == 0:
        v = np.full(N, float(v))
    return np.asarray(v, dtype=float)


# =================================================================================================
# 1-D spaces
# =================================================================================================
class OrthogonalSpace:
    """galerkin/orthogonal.py:37-444 (transform-relevant subset)."""

    def __init__(self, N, domain=None):
        self.N = N
        self.num_quad_points = N
        self.domain = tuple(self.reference_domain) if domain is None else tuple(domain)

    # orthogonal.py:343-354
    @property
    def domain_factor(self):
        a, b = (float(v) for v in self.domain)
        c, d = (float(v) for v in self.reference_domain)
        L, R = b - a, d - c
        return R / L if abs(L - R) > 1e-12 else 1

    # orthogonal.py:396-426
    def map_reference_domain(self, x):
        if tuple(map(float, self.domain)) != tuple(map(float, self.reference_domain)):
            return float(self.reference_domain[0]) + (x - float(self.domain[0])) * float(self.domain_factor)
        return x

    def map_true_domain(self, X):
        if tuple(map(float, self.domain)) != tuple(map(float, self.reference_domain)):
            return float(self.domain[0]) + (X - float(self.reference_domain[0])) / float(self.domain_factor)
        return X

    # orthogonal.py:428-444
    def mesh(self, kind="quadrature", N=None):
        N = self.num_quad_points if N is None else N
        if kind == "quadrature":
            return self.map_true_domain(self.quad_points_and_weights(N)[0])
        a, b = self.domain
        return np.linspace(float(a), float(b), N)

    # orthogonal.py:131-141, 168-178 (jacn(k=0) == plain vmap of eval_basis_functions)
    def vandermonde(self, X):
        return self.eval_basis_functions(np.asarray(X, dtype=float))

    # orthogonal.py:102-129: evaluate -> _evaluate
    def evaluate(self, x, c, axis=-1):
        X = self.map_reference_domain(np.asarray(x, dtype=float))
        return _along(lambda cm: self._evaluate(X, cm), c, axis)

    def _evaluate(self, X, c):
        # c: [nc, m]; generic: eval_basis_functions(X)[..., :len(c)] @ c
        return self.eval_basis_functions(X)[:, : c.shape[0]] @ c

    # orthogonal.py:214-227
    def backward(self, c, N=None, axis=-1):
        xj = self.mesh("quadrature", N)
        return self.evaluate(xj, c, axis)

    # orthogonal.py:229-246
    def backward_primitive(self, c, k=0, N=None, axis=-1):
        df = float(self.domain_factor ** k)
        return df * self.backward(self.derivative_coeffs(c, k, axis), N=N, axis=axis)

    #: coordinate system: None = Cartesian (sg == 1); else an object with `.sg` (SymPy expression in the true coordinate)
    #: and `.base_scalars()` (jaxfun.coordinates.CoordSys protocol)
    system = None

    # orthogonal.py:264-277
    def scalar_product(self, u, axis=-1):
        def f(um):
            N = um.shape[0]
            xj, wj = self.quad_points_and_weights(N)
            Pi = self.vandermonde(xj)
            sg = (sp.Integer(1) if self.system is None else sp.sympify(self.system.sg)) / self.domain_factor   # :270
            if sp.sympify(sg).is_number:
                wj = wj * float(sg)
            else:                                                                                            # :273-276
                x = self.system.base_scalars()[0]
                a, c, d = float(self.domain[0]), float(self.reference_domain[0]), self.domain_factor
                sgx = sp.lambdify(x, sg.xreplace({x: a + (x - c) / d}), modules="numpy")(np.asarray(xj))        # map_expr_true_domain
                wj = wj * sgx
            return ((um.T * wj) @ np.conj(Pi)).T
        return _along(f, u, axis)

    # orthogonal.py:256-262
    def forward(self, u, axis=-1):
        Lp = self.scalar_product(u, axis)
        A = self.norm_squared() / float(self.domain_factor)
        shp = [1] * Lp.ndim
        shp[axis] = -1
        return Lp / A.reshape(shp)

    def evaluate_mesh(self, c, kind="quadrature", N=None, axis=-1):
        if kind == "quadrature":
            return self.backward(c, N, axis)
        return self.evaluate(self.mesh(kind, N), c, axis)

    def derivative_coeffs(self, c, k=0, axis=-1):
        if k == 0:
            return c
        if k > 1:
            return self.derivative_coeffs(self.derivative_coeffs(c, k - 1, axis), 1, axis)
        return _along(self._derivative1, c, axis)


class Jacobi(OrthogonalSpace):
    """galerkin/Jacobi.py:25-447."""
    reference_domain = (-1, 1)

    def __init__(self, N, domain=None, alpha=0, beta=0):
        self.alpha, self.beta = sp.nsimplify(alpha), sp.nsimplify(beta)
        super().__init__(N, domain)

    # Jacobi.py:344-357
    def gn(self, n):
        return sp.S.One

    # Jacobi.py:359-374
    def _a(self, i, j):
        a, b = self.alpha, self.beta
        return (
            2 * (j + a) * (j + b) / ((2 * j + a + b + 1) * (2 * j + a + b)) * delta(i + 1, j)
            - (a**2 - b**2) / ((2 * j + a + b + 2) * (2 * j + a + b)) * delta(i, j)
            + 2 * (j + 1) * (j + a + b + 1) / ((2 * j + a + b + 2) * (2 * j + a + b + 1)) * delta(i - 1, j)
        )

    # Jacobi.py:421-447
    def a(self, i, j):
        return sp.simplify(self.gn(j) / self.gn(i) * self._a(i, j))

    # Jacobi.py:376-391
    def _b(self, i, j):
        a, b = self.alpha, self.beta
        d = lambda m, n: int(m == n)  # noqa: E731
        f = (2 * (i + a + b) / ((2 * i + a + b) * (2 * i + a + b - 1))) * d(i, j + 1) - (
            (2 * (i + a + 1) * (i + b + 1)) / ((2 * i + a + b + 3) * (2 * i + a + b + 2) * (i + a + b + 1))
        ) * d(i, j - 1)
        if a != b:
            f += ((2 * (a**2 - b**2)) / ((a + b) * (2 * i + a + b + 2) * (2 * i + a + b))) * d(i, j)
        return f

    # Jacobi.py:393-419
    def b(self, i, j):
        return sp.simplify(self.gn(j) / self.gn(i) * self._b(i, j))

    # Jacobi.py:290-342
    def psi(self, n, k):
        return sp.rf(n + self.alpha + self.beta + 1, k) * sp.Rational(1, 2**k)

    # Jacobi.py:257-288: k-th derivative of basis function i at X = -1 / X = +1 (reference coordinate)
    def bnd_values(self, k=0):
        alpha, beta = self.alpha, self.beta

        def gam(i):
            return self.psi(i, k) if k > 0 else 1

        def left_fn(i):
            return self.gn(i) * (-1) ** (k + i) * gam(i) * sp.binomial(i + beta, i - k)

        def right_fn(i):
            return self.gn(i) * gam(i) * sp.binomial(i + alpha, i - k)

        return left_fn, right_fn

    def h(self, n, k=0):
        assert k == 0, "only the k = 0 norm is on the transform path"
        f = sp.rf(n + 1, alf) / sp.rf(n + bet + 1, alf) * 2 ** (alf + bet + 1) / (2 * n + alf + bet + 1)
        f = sp.simplify(f.subs([(alf, self.alpha), (bet, self.beta)]))
        return sp.simplify(self.gn(n) ** 2 * f)

    # Jacobi.py:249-251
    def norm_squared(self):
        return _lamb(self.h(n_sym, 0), self.N)

    @functools.lru_cache(maxsize=None)
    def _rec(self, N):
        am = _lamb(self.a(n_sym + 1, n_sym), N)
        ap = _lamb(self.a(n_sym, n_sym + 1), N)
        aa = np.zeros_like(am)
        if self.alpha != self.beta:
            aa = _lamb(self.a(n_sym, n_sym), N)
        return am, ap, aa

    # Jacobi.py:112-124
    def quad_points_and_weights(self, N=None):
        N = self.num_quad_points if N is None else N
        x, w = roots_jacobi(N, float(self.alpha), float(self.beta))
        return np.array(x), np.array(w)

    #: bench-only switch: evaluate the series as Vandermonde @ c (the generic orthogonal.py:129 form,
    #: BLAS-backed) instead of the reference's N-step recurrence scan; same sums, different rounding
    fast_backward = False

    # Jacobi.py:65-110 (vectorised over points X and columns of c)
    def _evaluate(self, X, c):
        if self.fast_backward:
            return OrthogonalSpace._evaluate(self, X, c)
        N = c.shape[0]
        am, ap, aa = self._rec(N)
        X = np.asarray(X, dtype=float)[:, None]
        x0 = np.ones_like(X)
        if N == 1:
            return c[0] * x0
        x1 = (X - aa[0]) / am[0] * x0
        if N == 2:
            return c[0] * x0 + c[1] * x1
        x1_first = x1
        acc = np.zeros((X.shape[0], c.shape[1]), dtype=np.result_type(c.dtype, float))
        for i in range(2, N):
            x2 = ((X - aa[i - 1]) * x1 - ap[i - 2] * x0) / am[i - 1]
            acc = acc + x2 * c[i]
            x0, x1 = x1, x2
        return acc + c[0] + c[1] * x1_first

    # Jacobi.py:165-202
    def eval_basis_functions(self, X):
        X = np.atleast_1d(np.asarray(X, dtype=float))
        am, ap, aa = self._rec(self.N)
        x0 = X * 0 + 1
        cols = [x0]
        if self.N > 1:
            x1 = (X - aa[0]) / am[0] * x0
            cols.append(x1)
            for n in range(2, self.N):
                x2 = ((X - aa[n - 1]) * x1 - ap[n - 2] * x0) / am[n - 1]
                cols.append(x2)
                x0, x1 = x1, x2
        return np.stack(cols, axis=-1)

    # Jacobi.py:204-247
    def _derivative1(self, c):
        N = c.shape[0] - 1
        out = np.zeros_like(c)
        if N <= 0:
            return out
        bm = _lamb(self.b(n_sym + 1, n_sym), N)
        bp = _lamb(self.b(n_sym + 1, n_sym + 2), N)
        bb = np.zeros_like(bm)
        if self.alpha != self.beta:
            bb = _lamb(self.b(n_sym + 1, n_sym + 1), N)
        x0 = np.zeros_like(c[0])
        x1 = c[-1] / bm[-1]
        out[N - 1] = x1
        for n in range(N - 2, -1, -1):
            x2 = (c[n + 1] - bb[n] * x1 - bp[n] * x0) / bm[n]
            out[n] = x2
            x0, x1 = x1, x2
        return out


class Legendre(Jacobi):
    """galerkin/Legendre.py:42-216."""

    def __init__(self, N, domain=None):
        super().__init__(N, domain, 0, 0)

    # Legendre.py:125-137
    def quad_points_and_weights(self, N=None):
        N = self.num_quad_points if N is None else N
        x, w = leggauss(N)
        return x, w

    # Legendre.py:162-183
    def eval_basis_functions(self, X):
        X = np.atleast_1d(np.asarray(X, dtype=float))
        x0 = X * 0 + 1
        cols = [x0]
        x1 = X
        for i in range(2, self.N + 1):
            cols.append(x1)
            x2 = (x1 * X * (2 * i - 1) - x0 * (i - 1)) / i
            x0, x1 = x1, x2
        return np.stack(cols[: self.N], axis=-1)

    # Legendre.py:185-216
    def _derivative1(self, c):
        N = c.shape[0] - 1
        out = np.zeros_like(c)
        if N <= 0:
            return out
        x0 = np.zeros_like(c[0])
        x1 = c[-1] * (2 * N - 1)
        out[N - 1] = x1
        for n in range(N - 2, -1, -1):
            x2 = (2 * n + 1) * c[n + 1] + (2 * n + 1) / (2 * n + 5) * x0
            out[n] = x2
            x0, x1 = x1, x2
        return out


class Chebyshev(Jacobi):
    """galerkin/Chebyshev.py:44-397."""

    def __init__(self, N, domain=None):
        super().__init__(N, domain, -sp.S.Half, -sp.S.Half)

    # Chebyshev.py:320-338
    def gn(self, n):
        return sp.S.One / sp.jacobi(n, self.alpha, self.beta, 1)

    def h(self, n, k):
        if k > 0:
            return sp.simplify(sp.pi * n * sp.gamma(n + k) / (2 * sp.factorial(n - k)))
        return sp.Piecewise((sp.pi, sp.Eq(n, 0)), (sp.pi / 2, True))

    # Chebyshev.py:352-384
    def a(self, i, j):
        if (i - j) == 1:
            return sp.Piecewise((1, sp.Eq(j, 0)), (sp.S.Half, True))
        if (j - i) == 1:
            return sp.S.Half
        return 0

    # Chebyshev.py:149-170
    def quad_points_and_weights(self, N=None):
        N = self.num_quad_points if N is None else N
        return (np.cos(np.pi + (2 * np.arange(N) + 1) * np.pi / (2 * N)), np.ones(N) * np.pi / N)

    # Chebyshev.py:199-223
    def eval_basis_functions(self, X):
        X = np.atleast_1d(np.asarray(X, dtype=float))
        x0 = X * 0 + 1
        cols = [x0]
        x1 = X
        for _ in range(self.N - 1):
            cols.append(x1)
            x0, x1 = x1, 2 * X * x1 - x0
        return np.stack(cols, axis=-1)

    # Chebyshev.py:225-241 (linear extension to complex data: SURVEY §8a note)
    def backward(self, c, N=None, axis=-1):
        n = self.num_quad_points if N is None else N

        def f(cm):
            if n > cm.shape[0]:
                cm = np.concatenate([cm, np.zeros((n - cm.shape[0],) + cm.shape[1:], dtype=cm.dtype)])
            sign = (-1) ** np.arange(n)
            uh = cm * sign[:, None]
            return 0.5 * uh[0] + n * _idct2(uh, n)
        return _along(f, c, axis)

    # Chebyshev.py:243-260
    def forward(self, u, axis=-1):
        def f(um):
            n = um.shape[0]
            assert n >= self.N
            sign = (-1) ** np.arange(n)
            uh = _dct2(um, n)
            uh[0] = uh[0] / 2
            uh = uh * sign[:, None] / n
            return uh[: self.N]
        return _along(f, u, axis)

    # Chebyshev.py:262-279
    def scalar_product(self, u, axis=-1):
        def f(um):
            n = um.shape[0]
            assert n >= self.N
            sign = (-1) ** np.arange(n)
            uh = _dct2(um, n)
            uh = uh * np.pi * sign[:, None] / n / 2 / self.domain_factor
            return uh[: self.N]
        return _along(f, u, axis)

    # Chebyshev.py:281-315
    def _derivative1(self, c):
        N = c.shape[0] - 1
        out = np.zeros_like(c)
        if N <= 0:
            return out
        x0 = np.zeros_like(c[0])
        x1 = c[-1] * N * 2
        out[N - 1] = x1
        for n in range(N - 2, -1, -1):
            x2 = 2 * (n + 1) * c[n + 1] + x0
            out[n] = x2
            x0, x1 = x1, x2
        out[0] = out[0] / 2
        return out


def _dct2(x, n):
    """jax.scipy.fft.dct(x, n=n) (type 2, norm=None) along axis 0; complex input = linear extension."""
    if np.iscomplexobj(x):
        return scipy.fft.dct(x.real, type=2, n=n, axis=0, workers=WORKERS) + 1j * scipy.fft.dct(x.imag, type=2, n=n, axis=0, workers=WORKERS)
    return scipy.fft.dct(x, type=2, n=n, axis=0, workers=WORKERS)


def _idct2(x, n):
    if np.iscomplexobj(x):
        return scipy.fft.idct(x.real, type=2, n=n, axis=0, workers=WORKERS) + 1j * scipy.fft.idct(x.imag, type=2, n=n, axis=0, workers=WORKERS)
    return scipy.fft.idct(x, type=2, n=n, axis=0, workers=WORKERS)


def dst(x, type=2, n=None):
    """utils/common.py:197-230 along axis 0."""
    N = x.shape[0] if n is None else n
    if x.shape[0] < N:
        x = np.concatenate([x, np.zeros((N - x.shape[0],) + x.shape[1:], dtype=x.dtype)])
    zeros = np.zeros((1,) + x.shape[1:], dtype=x.dtype)
    if type == 1:
        y = np.concatenate([zeros, x, zeros, -x[::-1]], axis=0)
        Y = np.fft.fft(y, axis=0)
        return -np.imag(Y[1 : N + 1])
    y = np.concatenate([x, -x[::-1]], axis=0)
    Y = np.fft.fft(y, axis=0)
    k = np.arange(N)
    tw = np.exp(-1j * np.pi * (k + 1) / (2 * N))
    return -np.imag(tw[:, None] * Y[1 : N + 1])


class ChebyshevU(Jacobi):
    """galerkin/ChebyshevU.py:14-221."""

    def __init__(self, N, domain=None):
        super().__init__(N, domain, sp.S.Half, sp.S.Half)

    def gn(self, n):
        return (n + 1) / sp.jacobi(n, self.alpha, self.beta, 1)

    def norm_squared(self):
        return np.full(self.N, np.pi / 2)

    # ChebyshevU.py:83-104
    def quad_points_and_weights(self, N=None):
        N = self.num_quad_points if N is None else N
        theta = (np.arange(N) + 1) * np.pi / (N + 1)
        points = np.cos(theta + np.pi)
        weights = np.full(N, np.pi / (N + 1)) * (1 - points**2)
        return points, weights

    # ChebyshevU.py:133-157
    def eval_basis_functions(self, X):
        X = np.atleast_1d(np.asarray(X, dtype=float))
        x0 = X * 0 + 1
        cols = [x0]
        x1 = 2 * X
        for _ in range(self.N - 1):
            cols.append(x1)
            x0, x1 = x1, 2 * X * x1 - x0
        return np.stack(cols, axis=-1)

    # ChebyshevU.py:162-177 (real data; complex handled part-wise)
    def backward(self, c, N=None, axis=-1):
        n = self.num_quad_points if N is None else N

        def g(cm):
            d = dst(cm, n=n, type=1)
            return (d / (2 * np.sin((np.arange(n) + 1) * np.pi / (n + 1)))[:, None])[::-1]

        def f(cm):
            return g(cm.real) + 1j * g(cm.imag) if np.iscomplexobj(cm) else g(cm)
        return _along(f, c, axis)

    # ChebyshevU.py:192-209
    def scalar_product(self, u, axis=-1):
        def g(um):
            n = um.shape[0]
            uh = um * np.sin(np.pi / (n + 1) * np.arange(1, n + 1))[:, None]
            uh = dst(uh, n=n, type=1)
            uh = uh * ((-1) ** np.arange(n) * np.pi / (2 * (n + 1) * self.domain_factor))[:, None]
            return uh[: self.N]

        def f(um):
            return g(um.real) + 1j * g(um.imag) if np.iscomplexobj(um) else g(um)
        return _along(f, u, axis)

    # ChebyshevU.py:179-190
    def forward(self, u, axis=-1):
        return self.scalar_product(u, axis) * (2 * self.domain_factor / np.pi)


class Ultraspherical(Jacobi):
    """galerkin/Ultraspherical.py:12-101."""

    def __init__(self, N, domain=None, lambda_=1):
        lam = sp.nsimplify(lambda_)
        super().__init__(N, domain, lam - sp.S.Half, lam - sp.S.Half)

    def gn(self, n):
        return sp.S.One / sp.jacobi(n, self.alpha, self.beta, 1)


def fourier_wavenumbers(N, eliminate_highest_freq=False):
    """galerkin/Fourier.py:14-20."""
    indices = np.arange(N)
    k = np.where(indices < (N + 1) // 2, indices, indices - N)
    if eliminate_highest_freq and N % 2 == 0:
        k[N // 2] = 0
    return k


class Fourier(OrthogonalSpace):
    """galerkin/Fourier.py:23-235."""

    @property
    def reference_domain(self):
        return (0, 2 * np.pi)

    def __init__(self, N, domain=None):
        assert N % 2 == 0
        super().__init__(N, domain)

    def wavenumbers(self, N=None, eliminate_highest_freq=False):
        return fourier_wavenumbers(self.N if N is None else N, eliminate_highest_freq)

    # Fourier.py:77-89
    def quad_points_and_weights(self, N=None):
        N = self.num_quad_points if N is None else N
        return np.arange(N, dtype=float) * 2 * np.pi / N, np.full(N, 2 * np.pi / N)

    # Fourier.py:221-235
    def mesh(self, kind="quadrature", N=None):
        a, b = self.domain
        N = self.num_quad_points if N is None else N
        return np.linspace(float(a), float(b), N, endpoint=False)

    # Fourier.py:105-116
    def eval_basis_functions(self, X):
        X = np.atleast_1d(np.asarray(X, dtype=float))
        return np.exp(1j * self.wavenumbers()[None, :] * X[:, None])

    def norm_squared(self):
        return np.ones(self.N) * 2 * np.pi

    # Fourier.py:126-148
    def backward(self, c, N=None, axis=-1):
        n = self.N if N is None else N

        def f(cm):
            Lc = cm.shape[0]
            assert n >= Lc
            if n > Lc:
                cm = np.concatenate([cm[: Lc // 2], np.zeros((n - Lc,) + cm.shape[1:], dtype=cm.dtype), cm[Lc // 2:]])
            return scipy.fft.ifft(cm, axis=0, norm="forward", workers=WORKERS)
        return _along(f, np.asarray(c, dtype=complex), axis)

    # Fourier.py:150-163
    def scalar_product(self, u, axis=-1):
        def f(um):
            out = scipy.fft.fft(um, axis=0, norm="forward", workers=WORKERS) * 2 * np.pi / float(self.domain_factor)
            return out[self.wavenumbers()] if um.shape[0] > self.N else out
        return _along(f, np.asarray(u, dtype=complex), axis)

    # Fourier.py:165-180
    def forward(self, u, axis=-1):
        def f(um):
            assert um.shape[0] >= self.N
            out = scipy.fft.fft(um, axis=0, norm="forward", workers=WORKERS)
            return out[self.wavenumbers()] if um.shape[0] > self.N else out
        return _along(f, np.asarray(u, dtype=complex), axis)

    # Fourier.py:206-219
    def derivative_coeffs(self, c, k=0, axis=-1):
        if k == 0:
            return c
        m = self.wavenumbers(eliminate_highest_freq=k % 2 == 1)
        shp = [1] * np.ndim(c)
        shp[axis] = -1
        return ((1j * m) ** k).reshape(shp) * c


# =================================================================================================
# tensor products  (galerkin/tensorproductspace.py:330-460, sharding.py:24-40)
# =================================================================================================
class TensorProductSpace:
    def __init__(self, *spaces, system=None):
        self.basespaces = list(spaces)
        self.system = system        # tensorproductspace.py:38-60; the factors keep sub-systems with sg == 1 (coordinates.py:1254)

    def __len__(self):
        return len(self.basespaces)

    def _axes(self, x):
        d = len(self)
        return [x.ndim - d + ax for ax in range(d)]

    def backward(self, c, N=None):
        c = np.asarray(c)
        for ax, axis in enumerate(self._axes(c)):
            c = self.basespaces[ax].backward(c, N=None if N is None else N[ax], axis=axis)
        return c

    def forward(self, u):
        u = np.asarray(u)
        for ax, axis in enumerate(self._axes(u)):
            u = self.basespaces[ax].forward(u, axis=axis)
        return u

    def scalar_product(self, u):
        u = np.asarray(u)
        if self.system is not None and sp.sympify(self.system.sg) != 1:      # tensorproductspace.py:376-379
            sg = sp.lambdify(self.system.base_scalars(), self.system.sg, modules="numpy")(*self.mesh())
            u = u * sg
        for ax, axis in enumerate(self._axes(u)):
            u = self.basespaces[ax].scalar_product(u, axis=axis)
        return u

    def backward_primitive(self, c, k, N=None):
        c = np.asarray(c)
        for ax, axis in enumerate(self._axes(c)):
            c = self.basespaces[ax].backward_primitive(c, k=k[ax], N=None if N is None else N[ax], axis=axis)
        return c

    # tensorproductspace.py:263-273: einsum("i,j,ij") / ("i,j,k,ijk") per point
    def evaluate(self, x, c):
        x = np.atleast_2d(np.asarray(x, dtype=float))
        Cs = [s.eval_basis_functions(np.asarray(s.map_reference_domain(x[:, i])))[:, : c.shape[i]] for i, s in enumerate(self.basespaces)]
        path = "pi,pj,ij->p" if len(self) == 2 else "pi,pj,pk,ijk->p"
        return np.einsum(path, *Cs, c)

    def mesh(self, N=None):
        d = len(self)
        out = []
        for ax, s in enumerate(self.basespaces):
            x = np.asarray(s.mesh("quadrature", None if N is None else N[ax]))
            shp = [1] * d
            shp[ax] = -1
            out.append(x.reshape(shp))
        return tuple(out)


# sharding.py:43-105: the slab algorithm on P simulated ranks (lists of local blocks)
def slab_transform(fns, blocks, sharded_axis, split_axis):
    """blocks[r]: local array of rank r, sharded along `sharded_axis`.  fns[ax] acts along axis ax.
    Phase 1 = unsharded axes, tiled all_to_all(split_axis -> concat sharded_axis), phase 2 = sharded axis."""
    P = len(blocks)
    d = blocks[0].ndim
    unsharded = [ax for ax in range(d) if ax != sharded_axis]
    loc = list(blocks)
    for ax in unsharded:
        loc = [fns[ax](b) for b in loc]
    pieces = [np.split(b, P, axis=split_axis) for b in loc]          # pieces[src][dst]
    loc = [np.concatenate([pieces[src][dst] for src in range(P)], axis=sharded_axis) for dst in range(P)]
    return [fns[sharded_axis](b) for b in loc]


# =================================================================================================
# nonlinear terms and integrators (integrators/base.py:230-260, nonlinear.py, etdrk4.py, rk4.py)
# =================================================================================================
def nonlinear_rhs(space, leaves, expr, uh, N=None, final="forward"):
    """testspace.forward(E(backward_primitive(uh, k_l)...)) — base.py:230-248 with the evaluator of
    nonlinear.py:89-176 given as a Python callable `expr(*leaf_values)`."""
    vals = []
    for k in leaves:
        if isinstance(space, TensorProductSpace):
            vals.append(space.backward_primitive(uh, k, N) if any(k) else space.backward(uh, N))
        else:
            vals.append(space.backward_primitive(uh, k, N) if k else space.backward(uh, N))
    e = expr(*vals)
    return space.forward(e) if final == "forward" else space.scalar_product(e)


def phi1(z):
    """etdrk4.py:21-25."""
    z = np.asarray(z, dtype=complex)
    small = np.abs(z) < 1e-7
    series = 1 + z / 2 + z**2 / 6 + z**3 / 24 + z**4 / 120
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(small, series, np.expm1(z) / z)


def phi2(z):
    """etdrk4.py:28-32."""
    z = np.asarray(z, dtype=complex)
    small = np.abs(z) < 1e-6
    series = 0.5 + z / 6 + z**2 / 24 + z**3 / 120 + z**4 / 720
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(small, series, (np.expm1(z) - z) / z**2)


def phi3(z):
    """etdrk4.py:35-39."""
    z = np.asarray(z, dtype=complex)
    small = np.abs(z) < 1e-5
    series = 1 / 6 + z / 24 + z**2 / 120 + z**3 / 720 + z**4 / 5040
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(small, series, (np.expm1(z) - z - z**2 / 2) / z**3)


def etdrk4_coefficients(dt, Ldiag):
    """etdrk4.py:108-121 + :42-52."""
    z = dt * np.asarray(Ldiag)
    E, E2 = np.exp(z), np.exp(z / 2)
    Q = 0.5 * phi1(z / 2)
    p1, p2, p3 = phi1(z), phi2(z), phi3(z)
    return E, E2, Q, p1 - 3 * p2 + 4 * p3, p2 - 2 * p3, 4 * p3 - p2


def etdrk4_step(u_hat, dt, coeffs, Nfun):
    """etdrk4.py:152-166 (diagonal operators: `@` is elementwise)."""
    E, E2, Q, f1, f2, f3 = coeffs
    dtQ = dt * Q
    n1 = Nfun(u_hat)
    a = E2 * u_hat + dtQ * n1
    n2 = Nfun(a)
    b = E2 * u_hat + dtQ * n2
    n3 = Nfun(b)
    c = E2 * a + dtQ * (2 * n3 - n1)
    n4 = Nfun(c)
    return E * u_hat + dt * ((f1 * n1) + 2 * (f2 * (n2 + n3)) + (f3 * n4))


def rk4_step(u_hat, dt, rhs):
    """rk4.py:14-20."""
    k1 = rhs(u_hat)
    k2 = rhs(u_hat + 0.5 * dt * k1)
    k3 = rhs(u_hat + 0.5 * dt * k2)
    k4 = rhs(u_hat + dt * k3)
    return u_hat + (dt / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)


def backward_euler_step(u_hat, dt, M, Lw, Nsp, forcing=None):
    """backward_euler.py:29-39 with diagonal mass M and weak-form linear operator Lw; Nsp = scalar-product
    nonlinear term (base.py:238-248)."""
    rhs = M * u_hat
    if forcing is not None:
        rhs = rhs + dt * forcing
    if Nsp is not None:
        rhs = rhs + dt * Nsp(u_hat)
    return rhs / (M - dt * Lw)


def imex_rk_step(u_hat, dt, tableau, M, Lw, Nsp, forcing=None):
    """imex_rk.py:54-159 for diagonal operators.  `tableau` has .explicit/.implicit with A, b, c and the
    stiff-accuracy properties (tableau.py:43-127)."""
    a_e, a_i = tableau.explicit.A, tableau.implicit.A
    b_e, b_i, c_i = tableau.explicit.b, tableau.implicit.b, tableau.implicit.c
    s = tableau.stages
    full = tableau.is_stiffly_accurate
    impl_only = (not full) and tableau.implicit_is_stiffly_accurate
    m_u = M * u_hat
    stages, nls, lins = [], [], []
    for i in range(s):
        rhs = m_u
        for j in range(i):
            if a_e[i][j] != 0.0:
                rhs = rhs + dt * a_e[i][j] * nls[j]
            if a_i[i][j] != 0.0:
                rhs = rhs + dt * a_i[i][j] * lins[j]
        if forcing is not None and c_i[i] != 0.0:
            rhs = rhs + dt * c_i[i] * forcing
        a_ii = a_i[i][i]
        st = rhs / M if a_ii == 0.0 else rhs / (M - dt * a_ii * Lw)
        stages.append(st)
        last = i == s - 1
        nls.append(None if (last and full) else Nsp(st))
        lins.append(None if (last and (full or impl_only)) else Lw * st)
    if full:
        return stages[-1]
    if impl_only:
        rhs = M * stages[-1]
        for j in range(s):
            w = b_e[j] - a_e[-1][j]
            if w != 0.0:
                rhs = rhs + dt * w * nls[j]
        return rhs / M
    rhs = m_u
    for j in range(s):
        if b_e[j] != 0.0:
            rhs = rhs + dt * b_e[j] * nls[j]
        if b_i[j] != 0.0:
            rhs = rhs + dt * b_i[j] * lins[j]
    if forcing is not None:
        rhs = rhs + dt * forcing
    return rhs / M


# =================================================================================================
# Composite (boundary-condition) bases  (galerkin/composite.py:121-349) — SURVEY §8(f) rank 1
# =================================================================================================
class Composite(OrthogonalSpace):
    """phi_i = sum_j S_ij P_j.  Each method applies the stencil around the orthogonal transform exactly as
    the reference does (composite.py:205-208, 277-284, 339-349); the mass solve is a dense solve here
    (the reference uses a banded LU of the same matrix, la/diamatrix.py:394-773)."""

    def __init__(self, N, orthogonal, stencil, scaling=1, domain=None, **kw):
        self.orthogonal = orthogonal(N, domain=domain, **kw)
        self.N = N
        self.num_quad_points = N
        self.domain = self.orthogonal.domain
        k = np.arange(N - 1)
        nsym = sp.Symbol("n", integer=True)
        shifts = sorted(int(s) for s in stencil)
        self.width = max(shifts) - min(shifts)
        rows = N - self.width
        S = np.zeros((rows, N))
        for sh in shifts:
            v = sp.lambdify(nsym, sp.sympify(stencil[sh]) / sp.sympify(scaling), modules="numpy")(k)
            v = np.atleast_1d(np.asarray(v, dtype=float))
            if v.shape[0] == 1:
                v = np.full(N - 1, float(v[0]))
            for i in range(rows):
                if 0 <= i + sh < N:
                    S[i, i + sh] = v[i]
        self.S = S
        h = np.asarray(self.orthogonal.norm_squared(), dtype=float) * np.ones(N) / float(self.orthogonal.domain_factor)
        self.mass = S @ np.diag(h) @ S.T                       # composite.py:310-313

    @property
    def reference_domain(self):
        return self.orthogonal.reference_domain

    @property
    def dim(self):
        return self.N - self.width

    def quad_points_and_weights(self, N=None):
        return self.orthogonal.quad_points_and_weights(N)

    def mesh(self, kind="quadrature", N=None):
        return self.orthogonal.mesh(kind, N)

    def to_orthogonal(self, a, axis=-1):                       # composite.py:277-280: a @ S
        return _along(lambda am: self.S.T @ am, a, axis)

    def from_orthogonal(self, a, axis=-1):                     # composite.py:282-284
        return _along(lambda am: np.linalg.solve(self.S @ self.S.T, self.S @ am), a, axis)

    def backward(self, c, N=None, axis=-1):                    # composite.py:205-208
        return self.orthogonal.backward(self.to_orthogonal(c, axis), N=N, axis=axis)

    def backward_primitive(self, c, k=0, N=None, axis=-1):     # composite.py:210-218
        return self.orthogonal.backward_primitive(self.to_orthogonal(c, axis), k=k, N=N, axis=axis)

    def scalar_product(self, u, axis=-1):                      # composite.py:346-349
        P = self.orthogonal.scalar_product(u, axis=axis)
        return _along(lambda pm: self.S @ pm, P, axis)

    def forward(self, u, axis=-1):                             # composite.py:339-344
        Lp = self.scalar_product(u, axis)
        return _along(lambda lm: np.linalg.solve(self.mass, lm), Lp, axis)

    def eval_basis_functions(self, X):
        return self.orthogonal.eval_basis_functions(X) @ self.S.T

    def evaluate(self, x, c, axis=-1):
        return self.orthogonal.evaluate(x, self.to_orthogonal(c, axis), axis)


# =================================================================================================
# Inhomogeneous boundary values: BCGeneric lifting basis and DirectSum  (galerkin/composite.py:40-118, 411-488,
# 502-638, 835-896) — SURVEY §8(f) rank 1
# =================================================================================================
_BC_DERIV = {"D": 0, "R": 0, "N": 1, "N2": 2, "N3": 3, "N4": 4}


class BoundaryConditions(dict):
    """composite.py:40-118.  Robin conditions ("R": u + alfa u', "W": u' + alfa u'') carry the tuple (alfa, value)."""

    def __init__(self, bc, domain=None):
        super().__init__({"left": dict(bc.get("left", {})), "right": dict(bc.get("right", {}))})
        if not isinstance(bc, BoundaryConditions) and domain is not None:   # composite.py:66-73: dict input is normalised
            df = 2 / (float(domain[1]) - float(domain[0]))
            for val in self.values():
                for k, v in val.items():
                    if k[0] == "N":
                        nd = int(k[1:]) if len(k) > 1 else 1
                        val[k] = v / df**nd

    def orderednames(self):                                     # composite.py:75-79
        return ["L" + k for k in sorted(self["left"])] + ["R" + k for k in sorted(self["right"])]

    def orderedvals(self):                                      # composite.py:81-88
        vals = [self[lr][k] for lr in ("left", "right") for k in sorted(self[lr])]
        return [v[1] if isinstance(v, (tuple, list)) else v for v in vals]

    def num_bcs(self):                                          # composite.py:90-92
        return len(self.orderedvals())

    def num_derivatives(self):                                  # composite.py:94-101
        return sum(_BC_DERIV[k] for v in self.values() for k in v)

    def get_homogeneous(self):                                  # composite.py:111-118
        return BoundaryConditions({lr: {k: 0 for k in v} for lr, v in self.items()})


def get_bc_basis(bcs, orthogonal):
    """composite.py:835-896: rows = lifting functions B_i = sum_j S_ij P_j with (boundary functional b)(B_i) = delta_bi;
    the first block of `nb` consecutive modes whose boundary matrix is invertible is used."""
    bcs = BoundaryConditions(bcs)
    nb = bcs.num_bcs()

    def computematrix(first):
        rows = []
        for key in bcs.orderednames():
            side, kind = key[0], key[1:]
            lr = 0 if side == "L" else 1
            if kind in "WR":                                     # composite.py:860-872
                k0 = 0 if kind == "R" else 1
                alfa = bcs["left" if side == "L" else "right"][kind][0]
                f0, f1 = orthogonal.bnd_values(k=k0)[lr], orthogonal.bnd_values(k=k0 + 1)[lr]
                rows.append([sp.simplify(f0(j) + alfa * f1(j)) for j in range(first, first + nb)])
                continue
            f = orthogonal.bnd_values(k=_BC_DERIV[kind])[lr]
            rows.append([sp.simplify(f(j)) for j in range(first, first + nb)])
        A = sp.Matrix(rows)
        return sp.simplify(A.solve(sp.eye(nb)).T)

    first, sol = 0, None
    for first in range(bcs.num_derivatives() + 1):
        try:
            sol = computematrix(first)
            break
        except Exception:                                       # sympy NonInvertibleMatrixError
            continue
    S = np.zeros((nb, first + nb))
    S[:, first:] = np.array(sol.tolist(), dtype=float)
    return S


class DirectSum:
    """V = Composite (+) BCGeneric (composite.py:502-638): the expansion is the homogeneous part plus a fixed boundary lift
    sum_b val_b B_b whose orthogonal coefficients are `bnd_vals @ S_bc`, zero padded (composite.py:583-597)."""

    def __init__(self, a, bcs):
        self.a = a
        self.bcs = BoundaryConditions(bcs, domain=tuple(a.domain))   # functionspace.py:134
        self.orthogonal = a.orthogonal
        self.N = a.N
        self.domain = a.domain
        small = type(a.orthogonal)(self.bcs.num_bcs() + self.bcs.num_derivatives(), domain=tuple(a.domain))
        self.S_bc = get_bc_basis(self.bcs, small)               # composite.py:440-450
        cb = np.asarray(self.bnd_vals(), dtype=float) @ self.S_bc   # BCGeneric.to_orthogonal: c @ S (composite.py:277-280)
        self.c_b = np.zeros(self.N)
        self.c_b[:cb.shape[0]] = cb

    def bnd_vals(self):                                         # composite.py:463-470
        return np.array([float(v) for v in self.bcs.orderedvals()])

    @property
    def dim(self):
        return self.a.dim

    def mesh(self, kind="quadrature", N=None):
        return self.a.mesh(kind, N)

    def _lift(self, x, axis, sign=1.0):
        shp = [1] * np.ndim(x)
        shp[axis] = -1
        return x + sign * self.c_b.reshape(shp)

    def to_orthogonal(self, c, axis=-1):                        # composite.py:583-589
        return self._lift(self.a.to_orthogonal(c, axis), axis)

    def from_orthogonal(self, x, axis=-1):                      # composite.py:591-597
        return self.a.from_orthogonal(self._lift(np.asarray(x), axis, -1.0), axis)

    def backward(self, c, N=None, axis=-1):                     # composite.py:611-614
        return self.orthogonal.backward(self.to_orthogonal(c, axis), N=N, axis=axis)

    def backward_primitive(self, c, k=0, N=None, axis=-1):      # composite.py:616-624
        return self.orthogonal.backward_primitive(self.to_orthogonal(c, axis), k=k, N=N, axis=axis)

    def forward(self, u, axis=-1):                              # composite.py:626-629
        return self.from_orthogonal(self.orthogonal.forward(u, axis=axis), axis)

    def scalar_product(self, u, axis=-1):                       # composite.py:631-634
        return self.a.scalar_product(u, axis)

    def evaluate(self, x, c, axis=-1):                          # composite.py:599-602
        return self.orthogonal.evaluate(x, self.to_orthogonal(c, axis), axis)


class DirectSumTPS:
    """tensorproductspace.py:575-851 for ONE DirectSum factor (the poisson2D_periodic case): the boundary data are
    projected onto the other factors (`project1D`, inner.py:1027-1046 = forward transform of their samples on the mesh),
    `to_orthogonal` adds `pad(v.to_orthogonal(bndvals))` for the tensor space v = (others x BCGeneric) (:817-830),
    backward / forward go through the orthogonal tensor product (:781-790)."""

    def __init__(self, spaces, bc_axis, bcs, boundary_samples):
        # spaces: factor spaces with the homogeneous Composite on `bc_axis`; boundary_samples[j]: datum j sampled on the
        # tensor mesh of the other axes (what lambdify(...)(V.mesh()) gives in project1D)
        self.spaces, self.b = list(spaces), bc_axis
        a = self.spaces[bc_axis]
        self.bcs = BoundaryConditions(bcs)
        small = type(a.orthogonal)(self.bcs.num_bcs() + self.bcs.num_derivatives(), domain=tuple(a.domain))
        self.S_bc = get_bc_basis(self.bcs, small)
        self.orth = TensorProductSpace(*[s.orthogonal if hasattr(s, "S") else s for s in self.spaces])   # get_orthogonal()
        others = [s for i, s in enumerate(self.spaces) if i != bc_axis]
        uh = []
        for g in boundary_samples:                               # project1D onto the other factors
            g = np.asarray(g)
            for ax, sp_ in enumerate(others):
                g = sp_.forward(g, axis=ax)
                if hasattr(sp_, "S"):
                    g = sp_.to_orthogonal(g, axis=ax)
            uh.append(g)
        bnd = np.stack(uh, axis=bc_axis)                         # the BCGeneric axis holds the boundary index (:705-708)
        ai = np.moveaxis(np.tensordot(self.S_bc.T, bnd, axes=(1, bc_axis)), 0, bc_axis)   # BCGeneric.to_orthogonal: @ S
        full = [s.orthogonal.N if hasattr(s, "S") else s.N for s in self.spaces]
        self.lift = np.pad(ai, [(0, full[i] - ai.shape[i]) for i in range(len(full))])     # jnp.pad (:826-829)

    def _hom(self, c, name):
        for ax, s in enumerate(self.spaces):
            if hasattr(s, "S"):
                c = getattr(s, name)(c, axis=ax)
        return c

    def to_orthogonal(self, c):                                  # :817-830
        return self._hom(np.asarray(c), "to_orthogonal") + self.lift

    def from_orthogonal(self, c):                                # :832-851
        return self._hom(np.asarray(c) - self.lift, "from_orthogonal")

    def backward(self, c, N=None):                               # :781-786
        return self.orth.backward(self.to_orthogonal(c), N=N)

    def forward(self, u):                                        # :788-790
        return self.from_orthogonal(self.orth.forward(u))


# =================================================================================================
# Wavenumber-batched banded solves of Fourier x polynomial systems  (la/tpmatrix.py, la/diamatrix.py)
# Pinned: tests/golden/reference_banded.npz is produced by executing the reference's own
# `_lu_banded_no_pivot_kernel` and `_make_wavenumber_vmap_solve` (function bodies taken from where they lie, on the numpy
# stand-in for jax) — tests/golden/make_golden_banded.py; tests/test_banded_host.py replays it against these functions.
# =================================================================================================
def dia_from_dense(A, tol=0.0):
    """(offsets, data) of a square matrix in the reference's column-aligned DIA form: data[d][j] = A[j - offsets[d], j]
    (the band convention of la/diamatrix.py:1944: band[center + off, j] = A[j - off, j])."""
    A = np.asarray(A)
    n = A.shape[0]
    scale = np.abs(A).max() if A.size else 0.0
    offs, rows = [], []
    for off in range(-(n - 1), n):
        dg = np.diagonal(A, off)
        if np.abs(dg).max(initial=0.0) > tol * scale:
            row = np.zeros(n, dtype=A.dtype)
            if off >= 0:
                row[off:] = dg           # entry (j - off, j) for j = off .. n-1
            else:
                row[:n + off] = dg       # entry (j - off, j) for j = 0 .. n-1+off
            offs.append(off)
            rows.append(row)
    return tuple(offs), np.array(rows)


def wavenumber_band_data(W, P):
    """B_data_batch[k, d, :] = sum_t W[t, k] * P[t, d, :]   (la/tpmatrix.py:1345-1347)."""
    return np.einsum("tf,tdp->fdp", np.asarray(W), np.asarray(P))


def banded_lu_no_pivot(data, offsets):
    """Batched LU without pivoting of banded matrices given as DIA data [n_sys, n_diags, n] on `offsets`
    (la/tpmatrix.py:743-768 -> la/diamatrix.py:1937-1973).  Returns (band_lu [n_sys, p + q + 1, n], p, q) with
    band[center + off, j] = entry (j - off, j), center = p: rows below center = multipliers of L, rows >= center = U."""
    data = np.asarray(data)
    n_sys, _, n = data.shape
    p = max((-o for o in offsets if o < 0), default=0)
    q = max((o for o in offsets if o > 0), default=0)
    center = p
    band = np.zeros((n_sys, p + q + 1, n), dtype=data.dtype)
    for d, off in enumerate(offsets):
        band[:, center + off, :] = data[:, d, :]
    for k in range(n):
        pivot = band[:, center, k]
        for s in range(1, p + 1):
            if k + s >= n:
                continue
            with np.errstate(divide="ignore", invalid="ignore"):
                f = band[:, center - s, k] / pivot        # a zero pivot is reported by the caller (WavenumberSolver)
            band[:, center - s, k] = f
            for u in range(1, q + 1):
                j = k + u
                if j < n:
                    band[:, center + u - s, j] -= f * band[:, center + u, j]
    return band, p, q


def banded_solve(band_lu, p, q, rhs):
    """Forward elimination and back substitution for every system (la/tpmatrix.py:637-680): band_lu [n_sys, p + q + 1, n],
    rhs [n_sys, n] -> x [n_sys, n]."""
    n_sys, _, n = band_lu.shape
    y = np.zeros((n_sys, n), dtype=np.result_type(band_lu.dtype, rhs.dtype))
    for i in range(n):
        acc = np.zeros(n_sys, dtype=y.dtype)
        for s in range(1, min(p, i) + 1):
            acc += band_lu[:, p - s, i - s] * y[:, i - s]        # L[i, i-s] at band[center - s][i - s]
        y[:, i] = rhs[:, i] - acc
    x = np.zeros_like(y)
    for i in range(n - 1, -1, -1):
        acc = np.zeros(n_sys, dtype=y.dtype)
        for s in range(1, min(q, n - 1 - i) + 1):
            acc += band_lu[:, p + s, i + s] * x[:, i + s]        # U[i, i+s] at band[center + s][i + s]
        x[:, i] = (y[:, i] - acc) / band_lu[:, p, i]
    return x


class WavenumberSolver:
    """`TPMatricesWavenumberSolver` (la/tpmatrix.py:686-1014) from the separable form of `tpmats_wavenumber_factor`
    (:1236-1354): systems flattened in C order over the Fourier axes, polynomial axis = `poly_axis` of `shape`."""

    def __init__(self, poly_axis, shape, W, P, offsets):
        self.poly_axis, self.shape, self.offsets = int(poly_axis), tuple(shape), tuple(int(o) for o in offsets)
        self.band_lu, self.p, self.q = banded_lu_no_pivot(wavenumber_band_data(W, P), self.offsets)
        d = np.abs(self.band_lu[:, self.p, :])
        if not np.all(np.isfinite(self.band_lu[:, self.p, :])) or d.min() == 0.0:
            raise ValueError("Matrix is singular or has a zero/non-finite pivot in LU factorisation without pivoting.")

    def solve(self, rhs):
        rhs = np.asarray(rhs)
        nd = rhs.ndim
        order = [a for a in range(nd) if a != self.poly_axis] + [self.poly_axis]      # tpmatrix.py:931-939
        inv = np.argsort(order)
        r2 = np.transpose(rhs, order)
        fshape = r2.shape[:-1]
        x = banded_solve(self.band_lu, self.p, self.q, r2.reshape(-1, rhs.shape[self.poly_axis]))
        return np.transpose(x.reshape(fshape + (rhs.shape[self.poly_axis],)), inv)
