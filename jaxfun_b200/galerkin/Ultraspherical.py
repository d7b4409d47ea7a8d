"""Ultraspherical (Gegenbauer) space — mirrors `jaxfun.galerkin.Ultraspherical.Ultraspherical`
(`src/jaxfun/galerkin/Ultraspherical.py:12-101`): Jacobi(lambda-1/2, lambda-1/2) scaled so that
C_n(1) = 1; transforms are the generic Vandermonde contraction."""
from __future__ import annotations

import numpy as np

from .Jacobi import Jacobi


class Ultraspherical(Jacobi):
    def __init__(self, N: int, domain=None, system=None, name: str = "Ultraspherical", fun_str: str = "C",
                 lambda_=1, **kw) -> None:
        Jacobi.__init__(self, N, domain=domain, system=system, name=name, fun_str=fun_str,
                        alpha=float(lambda_) - 0.5, beta=float(lambda_) - 0.5)

    @property
    def lambda_(self):
        return self.alpha + 0.5

    def gn_values(self, n: int) -> np.ndarray:
        return self._inv_jacobi_at_one(n)

    def gn_symbolic(self, n):
        """1 / P_n^{(alpha,beta)}(1) (Ultraspherical.py:74-83)."""
        import sympy as sp
        return sp.S.One / sp.jacobi(n, sp.nsimplify(self.alpha), sp.nsimplify(self.beta), 1)
