"""Chebyshev (first kind) space — mirrors `jaxfun.galerkin.Chebyshev.Chebyshev`
(`src/jaxfun/galerkin/Chebyshev.py:44-315`).

forward / scalar_product / backward are the reference's DCT-II / DCT-III formulas
(Chebyshev.py:225-279).  On the GPU they run as the engine's Chebyshev fast kernels when the
transform length has one; otherwise as dense tables holding exactly those cosine sums (exact
integer argument reduction), through the tensor-core contraction.
"""
from __future__ import annotations

import numpy as np

from .. import _lib as L
from .Jacobi import Jacobi


def _cos_pi_frac(num: np.ndarray, den: int) -> np.ndarray:
    """cos(pi * num / den) for integer arrays, argument reduced exactly to [0, 2*den)."""
    r = np.mod(num, 2 * den)
    return np.cos(np.pi * (r.astype(np.float64) / den))


class Chebyshev(Jacobi):
    fast_basis = L.BASIS_CHEBYSHEV

    def __init__(self, N: int, domain=None, system=None, name: str = "Chebyshev", fun_str: str = "T", **kw) -> None:
        Jacobi.__init__(self, N, domain=domain, system=system, name=name, fun_str=fun_str, alpha=-0.5, beta=-0.5)

    def gn_values(self, n: int) -> np.ndarray:
        return self._inv_jacobi_at_one(n)

    def quad_points_and_weights(self, N: int | None = None):
        # Chebyshev.py:149-170
        N = self.num_quad_points if N is None else N
        return (np.cos(np.pi + (2 * np.arange(N) + 1) * np.pi / (2 * N)), np.ones(N) * np.pi / N)

    def eval_basis_functions(self, X) -> np.ndarray:
        # T_{n+1} = 2 X T_n - T_{n-1}      (Chebyshev.py:199-223)
        X = np.atleast_1d(np.asarray(X, dtype=float))
        N = self.N
        V = np.empty((X.shape[0], N))
        V[:, 0] = X * 0 + 1
        if N > 1:
            V[:, 1] = X
        for i in range(2, N):
            V[:, i] = 2 * X * V[:, i - 1] - V[:, i - 2]
        return V

    def norm_squared(self) -> np.ndarray:
        h = np.full(self.N, np.pi / 2)
        h[0] = np.pi
        return h

    def _derivative_host(self, c: np.ndarray) -> np.ndarray:
        # Chebyshev.derivative_coeffs (Chebyshev.py:281-315)
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

    def _dense_table(self, op: int, n_coeff: int, n_quad: int, deriv: int) -> np.ndarray:
        key = ("T", op, n_coeff, n_quad, deriv)
        T = self._tables.get(key)
        if T is not None:
            return T
        n = n_quad
        j = np.arange(n)
        if op in (L.OP_FORWARD, L.OP_SCALAR_PRODUCT):
            k = np.arange(self.N)
            sign = np.where(k % 2 == 0, 1.0, -1.0)
            dct = 2.0 * _cos_pi_frac(np.outer(k, 2 * j + 1), 2 * n)        # DCT-II rows
            if op == L.OP_FORWARD:
                scale = sign / n
                scale[0] = scale[0] / 2
            else:
                scale = np.pi * sign / n / 2 / float(self.domain_factor)
            T = dct * scale[:, None]
        else:
            k = np.arange(n_coeff)
            sign = np.where(k % 2 == 0, 1.0, -1.0)
            T = _cos_pi_frac(np.outer(2 * j + 1, k), 2 * n) * sign[None, :]  # 0.5*uh0 + n*idct(uh)
            if deriv:
                T = (float(self.domain_factor) ** deriv) * (T @ self.derivative_matrix(deriv, n_coeff))
        T = np.ascontiguousarray(T)
        self._tables[key] = T
        return T
