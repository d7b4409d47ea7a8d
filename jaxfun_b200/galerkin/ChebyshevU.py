"""Chebyshev (second kind) space — mirrors `jaxfun.galerkin.ChebyshevU.ChebyshevU`
(`src/jaxfun/galerkin/ChebyshevU.py:84-209`).  The reference goes through a DST-I built from a
2(n+1)-point FFT (`utils/common.py:197-230`); here the same sine sums are dense tables run by
the tensor-core contraction (secondary basis: not part of the benchmark configurations)."""
from __future__ import annotations

import numpy as np

from .. import _lib as L
from .Jacobi import Jacobi


def _sin_pi_frac(num: np.ndarray, den: int) -> np.ndarray:
    r = np.mod(num, 2 * den)
    return np.sin(np.pi * (r.astype(np.float64) / den))


class ChebyshevU(Jacobi):
    def __init__(self, N: int, domain=None, system=None, name: str = "ChebyshevU", fun_str: str = "U", **kw) -> None:
        Jacobi.__init__(self, N, domain=domain, system=system, name=name, fun_str=fun_str, alpha=0.5, beta=0.5)

    def gn_values(self, n: int) -> np.ndarray:
        return (np.arange(n) + 1.0) * self._inv_jacobi_at_one(n)

    def quad_points_and_weights(self, N: int | None = None):
        N = self.num_quad_points if N is None else N
        theta = (np.arange(N) + 1) * np.pi / (N + 1)
        points = np.cos(theta + np.pi)
        weights = np.full(N, np.pi / (N + 1)) * (1 - points**2)
        return points, weights

    def eval_basis_functions(self, X) -> np.ndarray:
        X = np.atleast_1d(np.asarray(X, dtype=float))
        N = self.N
        V = np.empty((X.shape[0], N))
        V[:, 0] = X * 0 + 1
        if N > 1:
            V[:, 1] = 2 * X
        for i in range(2, N):
            V[:, i] = 2 * X * V[:, i - 1] - V[:, i - 2]
        return V

    def norm_squared(self) -> np.ndarray:
        return np.full(self.N, np.pi / 2)

    def _dense_table(self, op: int, n_coeff: int, n_quad: int, deriv: int) -> np.ndarray:
        key = ("T", op, n_coeff, n_quad, deriv)
        T = self._tables.get(key)
        if T is not None:
            return T
        n = n_quad
        j = np.arange(n)
        if op in (L.OP_FORWARD, L.OP_SCALAR_PRODUCT):
            # uh = dst1(u * sin(pi (j+1)/(n+1))) * (-1)^k pi / (2 (n+1) df), truncated   (ChebyshevU.py:192-209)
            k = np.arange(self.N)
            sign = np.where(k % 2 == 0, 1.0, -1.0)
            dst = 2.0 * _sin_pi_frac(np.outer(k + 1, j + 1), n + 1)
            T = dst * np.sin(np.pi / (n + 1) * np.arange(1, n + 1))[None, :]
            T = T * (sign * np.pi / (2 * (n + 1) * float(self.domain_factor)))[:, None]
            if op == L.OP_FORWARD:
                T = T * (2 * float(self.domain_factor) / np.pi)
        else:
            # u = (dst1(c, n) / (2 sin((m+1) pi/(n+1))))[::-1]                           (ChebyshevU.py:162-177)
            k = np.arange(n_coeff)
            m = n - 1 - j                                   # reversal
            dst = 2.0 * _sin_pi_frac(np.outer(m + 1, k + 1), n + 1)
            T = dst / (2 * np.sin((m + 1) * np.pi / (n + 1)))[:, None]
            if deriv:
                T = (float(self.domain_factor) ** deriv) * (T @ self.derivative_matrix(deriv, n_coeff))
        T = np.ascontiguousarray(T)
        self._tables[key] = T
        return T
