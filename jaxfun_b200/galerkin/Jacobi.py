"""Jacobi polynomial spaces P_n^{(alpha,beta)} — host tables for the engine.

Mirrors `jaxfun.galerkin.Jacobi.Jacobi` (`src/jaxfun/galerkin/Jacobi.py:25-251`): three-term
recurrence for the (optionally g_n-scaled) basis, Gauss-Jacobi nodes from scipy's `roots_jacobi`
(the same third-party call the reference makes, Jacobi.py:112-124), norms h_n and the
derivative-coefficient recurrence.  The reference derives its recurrence coefficients through
SymPy (`_a`, `_b`, `h0`, Jacobi.py:306-391); here the same closed forms are evaluated directly in
float64 (removable singularities at n = 0 taken analytically).

`backward` is the reference's recurrence *series evaluation* (Jacobi.py:65-110) restated as a
contraction with the table built by that SAME recurrence (`series_table`: its coefficient arrays are the reference's own
sympy expressions a(n+1, n), a(n, n+1), a(n, n) lambdified over arange(N), so the table columns carry the rounding of the
reference's scan), and runs on the FP64 tensor cores instead of an N-step sequential scan.  `forward` / `scalar_product`
use the Vandermonde of `eval_basis_functions`, as the reference does (orthogonal.py:264-277).  The two recurrences agree to
k^2 ulp only: at n = 1024 a single high mode differs by 2e-11 of its own magnitude between them, so the backward table must
follow `_evaluate`, not `vandermonde`, to stay within the 1e-12 parity bar for such inputs.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.special import roots_jacobi

from .orthogonal import Domain, OrthogonalSpace


class Jacobi(OrthogonalSpace):
    def __init__(self, N: int, domain=None, system=None, name: str = "Jacobi", fun_str: str = "J",
                 alpha=0, beta=0, **kw) -> None:
        self.alpha = alpha
        self.beta = beta
        domain = Domain(-1, 1) if domain is None else domain
        OrthogonalSpace.__init__(self, N, domain=domain, system=system, name=name, fun_str=fun_str)

    @property
    def reference_domain(self) -> Domain:
        return Domain(-1, 1)

    # ---- scaling g_n (1 for plain Jacobi; subclasses override) -----------------------------------
    def gn_values(self, n: int) -> np.ndarray:
        return np.ones(n)

    def _inv_jacobi_at_one(self, n: int) -> np.ndarray:
        """1 / P_k^{(alpha,beta)}(1) = k! Gamma(alpha+1) / Gamma(k+alpha+1), k < n (ratio recurrence)."""
        a = float(self.alpha)
        g = np.ones(n)
        for k in range(1, n):
            g[k] = g[k - 1] * k / (k + a)
        return g

    # ---- recurrence coefficients (Jacobi.py:359-447) ------------------------------------------------
    def recurrence_coefficients(self, n: int):
        """(am, ap, aa) with am[k]=a(k+1,k), ap[k]=a(k,k+1), aa[k]=a(k,k) for k < n, g_n scaling applied."""
        a, b = float(self.alpha), float(self.beta)
        k = np.arange(n, dtype=float)
        s = a + b
        with np.errstate(divide="ignore", invalid="ignore"):
            am = 2 * (k + 1) * (k + s + 1) / ((2 * k + s + 2) * (2 * k + s + 1))
            j = k + 1
            ap = 2 * (j + a) * (j + b) / ((2 * j + s + 1) * (2 * j + s))
            aa = -(a * a - b * b) / ((2 * k + s + 2) * (2 * k + s))
        # removable singularities at k = 0
        am[0] = 2.0 / (s + 2)
        aa[0] = (b - a) / (s + 2)
        g = self.gn_values(n + 2)
        am = am * g[:n] / g[1:n + 1]          # a(k+1,k): factor gn(j=k)/gn(i=k+1)
        ap = ap * g[1:n + 1] / g[:n]          # a(k,k+1): factor gn(j=k+1)/gn(i=k)
        if a == b:
            aa = np.zeros(n)
        return am, ap, aa

    def derivative_recurrence_coefficients(self, n: int):
        """(bm, bp, bb) with bm[k]=b(k+1,k), bp[k]=b(k+1,k+2), bb[k]=b(k+1,k+1), k < n (Jacobi.py:376-419)."""
        a, b = float(self.alpha), float(self.beta)
        s = a + b
        k = np.arange(n, dtype=float)
        i = k + 1
        with np.errstate(divide="ignore", invalid="ignore"):
            bm = 2 * (i + s) / ((2 * i + s) * (2 * i + s - 1))
            bp = -(2 * (i + a + 1) * (i + b + 1)) / ((2 * i + s + 3) * (2 * i + s + 2) * (i + s + 1))
            bb = np.zeros(n)
            if a != b:
                bb = (2 * (a * a - b * b)) / (s * (2 * i + s + 2) * (2 * i + s)) if s != 0 else \
                    (2 * (a - b)) / ((2 * i + s + 2) * (2 * i + s))
        g = self.gn_values(n + 3)
        bm = bm * g[:n] / g[1:n + 1]            # b(i=k+1, j=k):   gn(j)/gn(i)
        bp = bp * g[2:n + 2] / g[1:n + 1]       # b(i=k+1, j=k+2)
        return bm, bp, bb

    # ---- the reference's series-evaluation recurrence (Jacobi.py:65-110) ------------------------------
    def gn_symbolic(self, n):
        """Scaling g_n as a SymPy expression (Jacobi.py:344-357; subclasses override like the reference's do)."""
        import sympy as sp
        return sp.S.One

    def _a_symbolic(self, i, j):
        """`Jacobi.a(i, j)` (Jacobi.py:359-374, 421-447): gn(j) / gn(i) * _a(i, j), simplified when symbolic."""
        import sympy as sp
        a, b = sp.nsimplify(self.alpha), sp.nsimplify(self.beta)
        delta = lambda m, n: sp.KroneckerDelta(m, n)  # noqa: E731
        f = (2 * (j + a) * (j + b) / ((2 * j + a + b + 1) * (2 * j + a + b)) * delta(i + 1, j)
             - (a**2 - b**2) / ((2 * j + a + b + 2) * (2 * j + a + b)) * delta(i, j)
             + 2 * (j + 1) * (j + a + b + 1) / ((2 * j + a + b + 2) * (2 * j + a + b + 1)) * delta(i - 1, j))
        return sp.simplify(self.gn_symbolic(j) / self.gn_symbolic(i) * f)

    def series_coefficients(self, N: int):
        """(am, ap, aa) exactly as `Jacobi._evaluate` builds them (Jacobi.py:81-89): the SymPy expressions a(n+1, n),
        a(n, n+1) and — only when alpha != beta — a(n, n), lambdified and evaluated over arange(N) in float64."""
        import sympy as sp
        key = (type(self).__name__, str(sp.nsimplify(self.alpha)), str(sp.nsimplify(self.beta)), int(N))
        v = _SERIES_CACHE.get(key)
        if v is None:
            v = _SERIES_CACHE[key] = _series_coefficients(self, int(N))
        return v

    def series_table(self, X, n_coeff: int) -> np.ndarray:
        """[len(X), n_coeff] table whose column k is `_evaluate(X, e_k)` of the reference (Jacobi.py:91-110):
        x0 = 1, x1 = (X - aa[0]) / am[0] * x0, x2 = ((X - aa[i-1]) x1 - ap[i-2] x0) / am[i-1]."""
        X = np.atleast_1d(np.asarray(X, dtype=float))
        V = np.empty((X.shape[0], n_coeff))
        x0 = np.ones_like(X)
        V[:, 0] = x0
        if n_coeff == 1:
            return V
        am, ap, aa = self.series_coefficients(n_coeff)
        x1 = (X - aa[0]) / am[0] * x0
        V[:, 1] = x1
        for i in range(2, n_coeff):
            x2 = ((X - aa[i - 1]) * x1 - ap[i - 2] * x0) / am[i - 1]
            V[:, i] = x2
            x0, x1 = x1, x2
        return V

    # ---- host tables ---------------------------------------------------------------------------------
    def quad_points_and_weights(self, N: int | None = None):
        N = self.num_quad_points if N is None else N
        x, w = roots_jacobi(N, float(self.alpha), float(self.beta))
        return np.asarray(x), np.asarray(w)

    def eval_basis_functions(self, X) -> np.ndarray:
        X = np.atleast_1d(np.asarray(X, dtype=float))
        N = self.N
        V = np.empty((X.shape[0], N))
        V[:, 0] = 1.0
        if N == 1:
            return V
        am, ap, aa = self.recurrence_coefficients(N)
        V[:, 1] = (X - aa[0]) / am[0]
        for n in range(2, N):
            V[:, n] = ((X - aa[n - 1]) * V[:, n - 1] - ap[n - 2] * V[:, n - 2]) / am[n - 1]
        return V

    def h0(self, n: int) -> np.ndarray:
        """Norms squared of the unscaled P_k^{(alpha,beta)}, k < n (Jacobi.py:306-326), by ratio recurrence."""
        a, b = float(self.alpha), float(self.beta)
        h = np.empty(n)
        h[0] = 2.0 ** (a + b + 1) * math.gamma(a + 1) * math.gamma(b + 1) / math.gamma(a + b + 2)
        for k in range(n - 1):
            if k == 0:  # (a+b+1) cancels analytically in the ratio
                r = (1 + a) * (1 + b) / (a + b + 3)
            else:
                r = ((k + 1 + a) * (k + 1 + b)) / ((k + 1) * (k + 1 + a + b)) * (2 * k + a + b + 1) / (2 * k + a + b + 3)
            h[k + 1] = h[k] * r
        return h

    def norm_squared(self) -> np.ndarray:
        g = self.gn_values(self.N)
        return g * g * self.h0(self.N)

    def _derivative_host(self, c: np.ndarray) -> np.ndarray:
        """Jacobi.derivative_coeffs (Jacobi.py:204-247) along axis 0, vectorised over columns."""
        n1 = c.shape[0]
        N = n1 - 1
        out = np.zeros_like(c)
        if N <= 0:
            return out
        bm, bp, bb = self.derivative_recurrence_coefficients(N)
        x0 = np.zeros_like(c[0])
        x1 = c[-1] / bm[-1]
        out[N - 1] = x1
        for n in range(N - 2, -1, -1):
            x2 = (c[n + 1] - bb[n] * x1 - bp[n] * x0) / bm[n]
            out[n] = x2
            x0, x1 = x1, x2
        return out


_SERIES_CACHE: dict = {}


def _series_coefficients(space, N: int):
    import sympy as sp
    n = sp.Symbol("n", integer=True, positive=True)

    def lamb(expr):
        v = sp.lambdify(n, expr, modules=["scipy", "numpy"])(np.arange(N))
        if np.ndim(v) == 0:                                   # Jacobi.py:83-86: scalar results are broadcast
            v = np.full(N, float(v))
        return np.asarray(v, dtype=float)

    am = lamb(space._a_symbolic(n + 1, n))
    ap = lamb(space._a_symbolic(n, n + 1))
    aa = np.zeros_like(am)
    if sp.nsimplify(space.alpha) != sp.nsimplify(space.beta):
        aa = lamb(space._a_symbolic(n, n))
    for v in (am, ap, aa):
        v.setflags(write=False)
    return am, ap, aa
