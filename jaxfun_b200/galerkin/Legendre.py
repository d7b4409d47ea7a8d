"""Legendre space (Jacobi alpha = beta = 0) — mirrors `jaxfun.galerkin.Legendre.Legendre`
(`src/jaxfun/galerkin/Legendre.py:42-216`): FastGL nodes, P_n recurrence, legder recurrence."""
from __future__ import annotations

import numpy as np

from ..utils.fastgl import leggauss
from .Jacobi import Jacobi


class Legendre(Jacobi):
    def __init__(self, N: int, domain=None, system=None, name: str = "Legendre", fun_str: str = "P", **kw) -> None:
        Jacobi.__init__(self, N, domain=domain, system=system, name=name, fun_str=fun_str, alpha=0, beta=0)

    def quad_points_and_weights(self, N: int | None = None):
        N = self.num_quad_points if N is None else N
        x, w = leggauss(N)
        return x, w

    def eval_basis_functions(self, X) -> np.ndarray:
        # P_i = (P_{i-1} X (2i-1) - P_{i-2} (i-1)) / i      (Legendre.py:162-183)
        X = np.atleast_1d(np.asarray(X, dtype=float))
        N = self.N
        V = np.empty((X.shape[0], N))
        V[:, 0] = X * 0 + 1
        if N > 1:
            V[:, 1] = X
        for i in range(2, N):
            V[:, i] = (V[:, i - 1] * X * (2 * i - 1) - V[:, i - 2] * (i - 1)) / i
        return V

    def norm_squared(self) -> np.ndarray:
        return 2.0 / (2.0 * np.arange(self.N) + 1.0)

    def _derivative_host(self, c: np.ndarray) -> np.ndarray:
        # Legendre.derivative_coeffs (Legendre.py:185-216)
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
