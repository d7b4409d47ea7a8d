"""Boundary-condition (composite) bases — host mirror of `jaxfun.galerkin.composite.Composite`
(`src/jaxfun/galerkin/composite.py:121-349`), SURVEY.md §8(f) rank 1.

phi_i = sum_j S_ij P_j with a banded stencil matrix S ([N - w, N]); all transforms are those of the
underlying orthogonal basis wrapped by S:

    to_orthogonal(a)   = a @ S                                   composite.py:277-280
    backward(c, N)     = orthogonal.backward(c @ S, N)           composite.py:205-208
    scalar_product(u)  = orthogonal.scalar_product(u) @ S^T      composite.py:346-349
    forward(u)         = M^-1 scalar_product(u),  M = S diag(h/df) S^T     composite.py:339-344, 310-313
    from_orthogonal(a) = (S S^T)^-1 (S a)                        composite.py:282-284

On the GPU nothing new runs: S (and M^-1) are folded into the dense per-axis tables on the host, so every
composite transform is ONE pass of the FP64 tensor-core contraction — the stencil costs nothing.  The
reference derives stencils symbolically from boundary conditions (`get_stencil_matrix`, composite.py:765+);
here the stencil is given explicitly ({shift: value or SymPy expression in n}) or selected for the common
homogeneous Dirichlet / Neumann cases on Chebyshev / Legendre by `FunctionSpace`.
"""
from __future__ import annotations

import numpy as np
import sympy as sp

from .. import _lib as L
from .orthogonal import OrthogonalSpace

n = sp.Symbol("n", integer=True)


def _stencil_rows(stencil: dict, scaling, N: int) -> tuple[list[int], list[np.ndarray]]:
    """Diagonal shifts and values val(k)/scaling(k), k = 0..N-2 (composite.py:263-275)."""
    k = np.arange(N - 1)
    shifts, vals = [], []
    for shift, val in sorted(stencil.items()):
        if isinstance(val, np.ndarray):     # numeric stencil (stencil_from_bcs): one value per basis index
            sc = np.atleast_1d(np.asarray(sp.lambdify(n, sp.sympify(scaling), modules="numpy")(k), dtype=float))
            v = np.zeros(N - 1)
            v[: val.shape[0]] = val[: N - 1]
            v = v / (sc if sc.shape[0] > 1 else float(sc[0]))
            shifts.append(int(shift))
            vals.append(v)
            continue
        expr = sp.sympify(val) / sp.sympify(scaling)
        f = sp.lambdify(n, expr, modules="numpy")
        v = np.atleast_1d(np.asarray(f(k), dtype=float))
        if v.shape[0] == 1:
            v = np.full(N - 1, float(v[0]))
        shifts.append(int(shift))
        vals.append(v)
    return shifts, vals


class Composite(OrthogonalSpace):
    is_orthogonal = False

    def __init__(self, N: int, orthogonal, bcs=None, domain=None, name: str = "Composite", fun_str: str = "phi",
                 system=None, stencil: dict | None = None, alpha=0, beta=0, scaling=None) -> None:
        if stencil is None:
            raise NotImplementedError("automatic stencil derivation from boundary conditions is not part of this "
                                      "build: pass `stencil={shift: value}` or use FunctionSpace for Dirichlet / Neumann")
        kw = {}
        if orthogonal.__name__ == "Jacobi":
            kw = dict(alpha=alpha, beta=beta)
        orth = orthogonal(N, domain=domain, system=system, **kw)
        self.orthogonal = orth             # needed by reference_domain during the base constructor ...
        super().__init__(N, domain=domain if domain is not None else tuple(orth.domain), system=system,
                         name=name, fun_str=fun_str)
        self.orthogonal = orth             # ... which resets it to `self`
        self.bcs = bcs
        self.scaling = sp.S.One if scaling is None else scaling
        self.stencil = dict(sorted(stencil.items()))
        shifts, vals = _stencil_rows(self.stencil, self.scaling, N)
        self._width = max(shifts) - min(shifts)
        rows = N - self._width
        S = np.zeros((rows, N))
        for sh, v in zip(shifts, vals):
            for i in range(rows):
                if 0 <= i + sh < N:
                    S[i, i + sh] = v[i]
        self.S = S
        h = np.asarray(self.orthogonal.norm_squared(), dtype=float) * np.ones(N) / float(self.orthogonal.domain_factor)
        self._mass = S @ np.diag(h) @ S.T                  # composite.py:310-313
        self._mass_inv = np.linalg.inv(self._mass)
        self._P_inv = np.linalg.inv(S @ S.T)

    # ---- delegated host tables -------------------------------------------------------------------
    @property
    def reference_domain(self):
        return self.orthogonal.reference_domain

    @property
    def dim(self) -> int:
        return self.orthogonal.N - self._width

    def stencil_width(self) -> int:
        return self._width

    def quad_points_and_weights(self, N=None):
        return self.orthogonal.quad_points_and_weights(self.num_quad_points if N is None else N)

    def eval_basis_functions(self, X):
        return self.orthogonal.eval_basis_functions(X) @ self.S.T

    def evaluate_basis_derivative(self, X, k: int = 0):
        return self.orthogonal.evaluate_basis_derivative(X, k) @ self.S.T

    def mass_matrix(self) -> np.ndarray:
        return self._mass

    def get_orthogonal(self):
        return self.orthogonal

    # ---- dense tables with the stencil folded in -------------------------------------------------
    def _dense_table(self, op: int, n_coeff: int, n_quad: int, deriv: int) -> np.ndarray:
        key = ("T", op, n_coeff, n_quad, deriv)
        T = self._tables.get(key)
        if T is not None:
            return T
        o = self.orthogonal
        if op in (L.OP_FORWARD, L.OP_SCALAR_PRODUCT):
            To = o._dense_table(L.OP_SCALAR_PRODUCT, o.N, n_quad, 0)     # [N, nq]
            T = self.S @ To
            if op == L.OP_FORWARD:
                T = self._mass_inv @ T
        else:
            assert n_coeff == self.dim, f"Coefficient length {n_coeff} != composite dimension {self.dim}"
            To = o._dense_table(op, o.N, n_quad, deriv)                   # [nq, N]
            T = To @ self.S.T
        T = np.ascontiguousarray(T)
        self._tables[key] = T
        return T

    def axis_spec(self, op: int, n_in: int, dtype: int, N=None, k: int = 0, inner: int = 1):
        from ..engine import AxisSpec
        if op in (L.OP_FORWARD, L.OP_SCALAR_PRODUCT):
            n_quad, n_coeff = n_in, self.dim
            assert n_quad >= self.orthogonal.N, "Only truncation supported for forward transform"
        else:
            n_quad = self.num_quad_points if N is None else int(N)
            n_coeff = n_in
        T = self._dense_table(op, n_coeff, n_quad, k)
        return AxisSpec(L.BASIS_TABLE, n_modes=n_coeff, n_quad=n_quad, deriv=k,
                        domain_factor=float(self.orthogonal.domain_factor), table=T)

    # ---- coefficient-space maps -------------------------------------------------------------------
    def to_orthogonal(self, a, axis: int = -1):
        return self._run(L.OP_APPLY, a, axis, table=np.ascontiguousarray(self.S.T))

    def from_orthogonal(self, a, axis: int = -1):
        return self._run(L.OP_APPLY, a, axis, table=np.ascontiguousarray(self._P_inv @ self.S))

    def _to_orthogonal_axis(self, c, axis):
        return self.to_orthogonal(c, axis)

    def _from_orthogonal_axis(self, c, axis):
        return self.from_orthogonal(c, axis)

    def derivative_matrix(self, k: int = 1, n: int | None = None):
        raise NotImplementedError("composite derivative coefficients live in the orthogonal basis: use to_orthogonal")

    def evaluate(self, x, c, axis: int = -1):
        X = np.atleast_1d(np.asarray(self.map_reference_domain(np.asarray(x, dtype=float))))
        T = np.ascontiguousarray(self.eval_basis_functions(X))
        return self._run(L.OP_APPLY, c, axis, table=T, cache=False)


_BC_ORDER = {"D": 0, "N": 1, "N2": 2, "N3": 3, "N4": 4}


def ordered_bc_names(bcs: dict) -> list[tuple[str, str]]:
    """[(side, kind)] in the reference's order: left before right, by derivative order
    (`BoundaryConditions.orderednames`, composite.py:40-118)."""
    out = []
    for side in ("left", "right"):
        for kind in sorted(bcs.get(side, {}), key=lambda v: _BC_ORDER[v]):
            out.append((side, kind))
    return out


def stencil_from_bcs(bcs: dict, orthogonal) -> dict:
    """Numeric counterpart of `get_stencil_matrix` (composite.py:765-838): phi_n = P_n + sum_{j=1..nb} d_j(n) P_{n+j}
    with d(n) solving  sum_j d_j f_b(n + j) = -f_b(n)  for every homogeneous boundary functional
    f_b(m) = d^k P_m / dX^k at X = -1 or +1.  The reference solves this system symbolically in n; here it is
    solved per basis index from the boundary values of the (derivative) Vandermonde.  Robin conditions are
    not covered."""
    names = ordered_bc_names(bcs)
    nb = len(names)
    N = orthogonal.N
    F = np.empty((nb, N))
    for b, (side, kind) in enumerate(names):
        if kind not in _BC_ORDER:
            raise NotImplementedError(f"boundary condition kind {kind!r} (Robin) has no numeric stencil here")
        if bcs[side][kind] != 0:
            raise NotImplementedError("inhomogeneous boundary values need the DirectSum lifting (composite.py:502-634)")
        X = np.array([-1.0 if side == "left" else 1.0])
        F[b] = orthogonal.evaluate_basis_derivative(X, _BC_ORDER[kind])[0]
    rows = N - nb
    d = np.zeros((nb, rows))
    for i in range(rows):
        A = F[:, i + 1: i + 1 + nb]
        d[:, i] = np.linalg.solve(A, -F[:, i])
    st = {0: np.ones(rows)}
    for j in range(nb):
        if np.abs(d[j]).max() > 1e-14:
            dj = d[j].copy()
            dj[np.abs(dj) < 1e-15] = 0.0
            st[j + 1] = dj
    if max(st) != nb:                       # keep the full width even when the last diagonal vanishes
        st[nb] = np.zeros(rows)
    return st


def FunctionSpace(N: int, space, bcs=None, domain=None, name: str = "fun", fun_str: str = "phi", scaling=None, **kw):
    """`jaxfun.galerkin.functionspace.FunctionSpace` (functionspace.py:63-173) for the cases whose stencil is
    known in closed form: no BCs -> the orthogonal space; homogeneous Dirichlet on both ends of a Chebyshev /
    Legendre space -> phi_k = P_k - P_{k+2} (both families have P_k(+-1) = (+-1)^k)."""
    if bcs is None:
        return space(N, domain=domain, name=name, fun_str=fun_str, **kw)
    left, right = bcs.get("left", {}), bcs.get("right", {})
    if set(left) == {"D"} and set(right) == {"D"} and left["D"] == 0 and right["D"] == 0 and \
            space.__name__ in ("Chebyshev", "Legendre"):
        return Composite(N, space, bcs=bcs, domain=domain, name=name, fun_str=fun_str, stencil={0: 1, 2: -1},
                         scaling=scaling)
    # any other set of homogeneous Dirichlet / Neumann / higher-derivative conditions: numeric stencil
    orth = space(N, domain=domain, **kw)
    return Composite(N, space, bcs=bcs, domain=domain, name=name, fun_str=fun_str, stencil=stencil_from_bcs(bcs, orth),
                     scaling=scaling, **kw)
