"""Diagonally-implicit IMEX Runge-Kutta and backward Euler for diagonal-operator spectral systems —
`jaxfun.integrators.imex_rk.IMEXRungeKutta` (`src/jaxfun/integrators/imex_rk.py:38-159`) and
`jaxfun.integrators.backward_euler.BackwardEuler` (`src/jaxfun/integrators/backward_euler.py:17-39`).

Semi-discrete system in coefficient space (weak form, mass NOT divided out, as in the reference):

        M d(uh)/dt = Lw uh + N_sp(uh) + f,       M = diag(h_k / df),  Lw = M * Ldiag,

with `N_sp = testspace.scalar_product(evaluator(uh))` (`integrators/base.py:238-248`).  Every stage is one
nonlinear evaluation (`jfx_nonlinear_execute` with a scalar_product final transform) plus fused diagonal
combinations (`jfx_axpby_diag`); the implicit solves are divisions by `M - dt a_ii Lw`."""
from __future__ import annotations

import numpy as np

from .base import BaseIntegrator, axpby_diag
from .tableau import IMEXTableau

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


class _WeakFormMixin:
    def _weak_setup(self):
        if getattr(self, "_M", None) is not None:
            return
        M = self.mass_diag_like(self.Ldiag)
        self._M = M
        self._Minv = 1.0 / M
        self._Lw = M * self.Ldiag
        self._nl_sp = None if self.nonlinear is None else self.nonlinear.with_final("scalar_product")

    def mass_diag_like(self, ref):
        """diag of the mass operator broadcast to the coefficient shape: prod_ax h_k / df (orthogonal.py:256-262)."""
        spaces = list(self.trialspace.basespaces) if hasattr(self.trialspace, "basespaces") else [self.trialspace]
        d = len(spaces)
        m = None
        for ax, sp in enumerate(spaces):
            h = np.asarray(sp.mass_diagonal(), dtype=float) * np.ones(sp.N)
            shp = [1] * d
            shp[ax] = sp.N
            h = h.reshape(shp)
            m = h if m is None else m * h
        m = np.broadcast_to(m, tuple(ref.shape[-d:])).copy()
        return torch.from_numpy(m).to(ref.device)

    def nonlinear_rhs_scalar_product(self, uh, N=None):
        if self._nl_sp is None:
            return torch.zeros_like(uh)
        if N is None:
            return self._nl_sp(uh)
        return self._nonlinear_for(N, final="scalar_product")(uh)      # base.py:238-248: N forwarded to the evaluator


class BackwardEuler(_WeakFormMixin, BaseIntegrator):
    def setup(self, dt: float) -> None:
        self._weak_setup()
        self._dt = dt
        self._sys_inv = 1.0 / (self._M - dt * self._Lw)          # (M - dt Lw)^-1, backward_euler.py:26-27

    def step(self, u_hat, dt: float, N=None):
        if getattr(self, "_dt", None) != dt:
            self.setup(dt)
        terms = [(1.0, self._M, u_hat)]
        if self.forcing is not None:
            terms.append((dt, None, self.forcing))
        if self.has_nonlinear:
            terms.append((dt, None, self.nonlinear_rhs_scalar_product(u_hat, N)))
        rhs = axpby_diag(terms)
        return axpby_diag([(1.0, self._sys_inv, rhs)])


class IMEXRungeKutta(_WeakFormMixin, BaseIntegrator):
    def __init__(self, space, linear_diag=None, nonlinear=None, forcing=None, *, tableau: IMEXTableau):
        super().__init__(space, linear_diag, nonlinear, forcing)
        self.tableau = tableau

    def setup(self, dt: float) -> None:
        self._weak_setup()
        self._dt = dt
        self._stage_inv = {a: 1.0 / (self._M - dt * a * self._Lw) for a in self.tableau.distinct_diagonal_coeffs}

    def stage(self, i, m_u, dt, nonlinear_stage, linear_stage, forcing):
        t = self.tableau
        a_e, a_i, c_i = t.explicit.A, t.implicit.A, t.implicit.c
        terms = [(1.0, None, m_u)]
        for j in range(i):
            if a_e[i][j] != 0.0:
                terms.append((dt * a_e[i][j], None, nonlinear_stage[j]))
            if a_i[i][j] != 0.0:
                terms.append((dt * a_i[i][j], None, linear_stage[j]))
        if forcing is not None and c_i[i] != 0.0:
            terms.append((dt * c_i[i], None, forcing))
        rhs = _sum(terms)
        a_ii = a_i[i][i]
        inv = self._Minv if a_ii == 0.0 else self._stage_inv[a_ii]
        return axpby_diag([(1.0, inv, rhs)])

    def step(self, u_hat, dt: float, N=None):
        if getattr(self, "_dt", None) != dt:
            self.setup(dt)
        t = self.tableau
        a_e, b_e, b_i = t.explicit.A, t.explicit.b, t.implicit.b
        full_gsa = t.is_stiffly_accurate
        implicit_only_sa = (not full_gsa) and t.implicit_is_stiffly_accurate
        last = t.stages - 1
        m_u = axpby_diag([(1.0, self._M, u_hat)])
        stages, nls, lins = [], [], []
        for i in range(t.stages):
            st = self.stage(i, m_u, dt, nls, lins, self.forcing)
            stages.append(st)
            is_last = i == last
            nls.append(None if (is_last and full_gsa) else self.nonlinear_rhs_scalar_product(st, N))
            lins.append(None if (is_last and (full_gsa or implicit_only_sa)) else axpby_diag([(1.0, self._Lw, st)]))
        if full_gsa:
            return stages[-1]
        if implicit_only_sa:
            terms = [(1.0, self._M, stages[-1])]
            for j in range(t.stages):
                w = b_e[j] - a_e[-1][j]
                if w != 0.0:
                    terms.append((dt * w, None, nls[j]))
            return axpby_diag([(1.0, self._Minv, _sum(terms))])
        terms = [(1.0, None, m_u)]
        for j in range(t.stages):
            if b_e[j] != 0.0:
                terms.append((dt * b_e[j], None, nls[j]))
            if b_i[j] != 0.0:
                terms.append((dt * b_i[j], None, lins[j]))
        if self.forcing is not None:
            terms.append((dt, None, self.forcing))
        return axpby_diag([(1.0, self._Minv, _sum(terms))])


def _sum(terms):
    """axpby_diag takes at most 8 terms per launch."""
    acc = None
    while terms:
        chunk, terms = terms[:7 if acc is not None else 8], terms[7 if acc is not None else 8:]
        if acc is not None:
            chunk = [(1.0, None, acc)] + chunk
        acc = axpby_diag(chunk)
    return acc
