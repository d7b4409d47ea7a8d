"""Tensor-product Galerkin solves by per-axis diagonalisation — the fast path of
`jaxfun.la.tpmatrix.TPMatrices.solve(method="lu")` / `tpmats_lu_factor`
(`src/jaxfun/la/tpmatrix.py:429-587`), SURVEY.md §8(f) rank 3.

A separable operator  sum_i (B_0 x .. x A_i x .. x B_{d-1})  (e.g. the weak Laplacian with stiffness A_i
and mass B_i per axis) is diagonalised by the generalised eigenvectors A_i V_i = B_i V_i Lambda_i:

        u = (V_0 x .. x V_{d-1})  [ ((B_0 V_0)^-1 x .. x (B_{d-1} V_{d-1})^-1) f  /  (lambda_0 (+) .. (+) lambda_{d-1}) ]

The factorisation (O(n^3) per axis) stays on the host; the solve is two multi-axis `JFX_OP_APPLY` plans —
the same FP64 tensor-core contraction as the Vandermonde transforms, with eigenvector tables — and one
diagonal scaling (`jfx_axpby_diag`).
"""
from __future__ import annotations

import numpy as np

from .. import _lib as L
from ..engine import AxisSpec, Plan, jfx_dtype


def stiffness_matrix(space, k: int = 2, nq: int | None = None) -> np.ndarray:
    """(phi_i, d^k phi_j / dx^k)_w by Gauss quadrature of the space (exact for polynomial bases with nq >= N):
    what `inner(v * Div(Grad(u)))` assembles per axis (`galerkin/inner.py:809-912`)."""
    o = space.get_orthogonal()
    nq = o.N if nq is None else nq
    xq, wq = o.quad_points_and_weights(nq)
    df = float(o.domain_factor)
    P = space.eval_basis_functions(xq)                    # [nq, dim]
    D = space.evaluate_basis_derivative(xq, k)            # [nq, dim]
    return (P * (wq / df)[:, None]).T @ D * df ** k


def mass_matrix(space, nq: int | None = None) -> np.ndarray:
    o = space.get_orthogonal()
    nq = o.N if nq is None else nq
    xq, wq = o.quad_points_and_weights(nq)
    P = space.eval_basis_functions(xq)
    return (P * (wq / float(o.domain_factor))[:, None]).T @ P


class KroneckerSumSolver:
    """Solve  sum_i (B_0 x .. x A_i x .. x B_{d-1}) u = f  for coefficient arrays on the device."""

    def __init__(self, mats):
        self.V, self.W, lams = [], [], []
        for A, B in mats:
            lam, V = np.linalg.eig(np.linalg.solve(B, A))
            if np.abs(lam.imag).max() > 1e-9 * np.abs(lam).max():
                raise ValueError("operator pencil has complex eigenvalues: not diagonalisable over the reals")
            order = np.argsort(lam.real)
            lam, V = lam.real[order], np.ascontiguousarray(V.real[:, order])
            self.V.append(V)
            self.W.append(np.ascontiguousarray(np.linalg.inv(B @ V)))
            lams.append(lam)
        d = len(mats)
        D = 0.0
        for i, lam in enumerate(lams):
            shp = [1] * d
            shp[i] = -1
            D = D + lam.reshape(shp)
        self.Dinv = np.ascontiguousarray(1.0 / D)
        self._plans = {}
        self._dinv_dev = {}

    def solve(self, f):
        import torch
        from ..integrators.base import axpby_diag
        dt = jfx_dtype(f.dtype)
        key = (tuple(f.shape), dt)
        if key not in self._plans:
            lead = f.ndim - len(self.V)
            pw = Plan(L.OP_APPLY, dt, tuple(f.shape), [None] * lead + [AxisSpec(L.BASIS_TABLE, table=W) for W in self.W])
            pv = Plan(L.OP_APPLY, dt, pw.shape_out, [None] * lead + [AxisSpec(L.BASIS_TABLE, table=V) for V in self.V])
            self._plans[key] = (pw, pv)
            dinv = torch.from_numpy(np.broadcast_to(self.Dinv, pw.shape_out).copy()).to(f.device)
            self._dinv_dev[key] = dinv
        pw, pv = self._plans[key]
        g = pw(f.contiguous())
        g = axpby_diag([(1.0, self._dinv_dev[key], g)])
        return pv(g)


def poisson_solver(T) -> KroneckerSumSolver:
    """Solver of the weak Laplace operator  (v, div grad u)_w  on the tensor-product space T."""
    return KroneckerSumSolver([(stiffness_matrix(s, 2), mass_matrix(s)) for s in T.basespaces])
