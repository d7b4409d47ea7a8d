"""ETDRK4 (Cox-Matthews / Kassam-Trefethen) for diagonal linear operators —
`jaxfun.integrators.etdrk4.ETDRK4` (`src/jaxfun/integrators/etdrk4.py:21-52, 108-121, 152-166`).

`setup(dt)` computes the diagonal coefficient arrays E, E2, Q, f1, f2, f3 once (host float64 /
complex128, small-argument series as in the reference) and uploads them; `step` is four nonlinear
evaluations plus five fused diagonal combinations on the device."""
from __future__ import annotations

import numpy as np

from .base import BaseIntegrator, axpby_diag

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


def _phi1(z):
    small = np.abs(z) < 1e-7
    series = 1 + z / 2 + z**2 / 6 + z**3 / 24 + z**4 / 120
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(small, series, np.expm1(z) / z)


def _phi2(z):
    small = np.abs(z) < 1e-6
    series = 0.5 + z / 6 + z**2 / 24 + z**3 / 120 + z**4 / 720
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(small, series, (np.expm1(z) - z) / z**2)


def _phi3(z):
    small = np.abs(z) < 1e-5
    series = 1 / 6 + z / 24 + z**2 / 120 + z**3 / 720 + z**4 / 5040
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(small, series, (np.expm1(z) - z - z**2 / 2) / z**3)


def etd_coefficients(dt: float, Ldiag: np.ndarray):
    """(E, E2, Q, f1, f2, f3) of etdrk4.py:108-121."""
    z = dt * np.asarray(Ldiag)
    E, E2 = np.exp(z), np.exp(z / 2)
    Q = 0.5 * _phi1(z / 2)
    p1, p2, p3 = _phi1(z), _phi2(z), _phi3(z)
    return E, E2, Q, p1 - 3 * p2 + 4 * p3, p2 - 2 * p3, 4 * p3 - p2


class ETDRK4(BaseIntegrator):
    def setup(self, dt: float) -> None:
        L = self.Ldiag.detach().cpu().numpy() if torch is not None and isinstance(self.Ldiag, torch.Tensor) else np.asarray(self.Ldiag)
        dev = self.Ldiag.device if torch is not None and isinstance(self.Ldiag, torch.Tensor) else "cuda"
        self._dt = dt
        self.E, self.E2, self.Q, self.f1, self.f2, self.f3 = (
            torch.from_numpy(np.ascontiguousarray(c)).to(dev) for c in etd_coefficients(dt, L))

    def _N(self, u_hat, N=None):
        n = self.nonlinear_rhs(u_hat, N)
        if self.forcing is not None:
            n = axpby_diag([(1.0, None, n), (1.0, None, self.forcing)])
        return n

    def step(self, u_hat, dt: float, N=None):
        if getattr(self, "_dt", None) != dt:
            self.setup(dt)
        E, E2, Q, f1, f2, f3 = self.E, self.E2, self.Q, self.f1, self.f2, self.f3
        n1 = self._N(u_hat, N)
        a = axpby_diag([(1.0, E2, u_hat), (dt, Q, n1)])
        n2 = self._N(a, N)
        b = axpby_diag([(1.0, E2, u_hat), (dt, Q, n2)])
        n3 = self._N(b, N)
        c = axpby_diag([(1.0, E2, a), (2.0 * dt, Q, n3), (-dt, Q, n1)])
        n4 = self._N(c, N)
        return axpby_diag([(1.0, E, u_hat), (dt, f1, n1), (2.0 * dt, f2, n2), (2.0 * dt, f2, n3), (dt, f3, n4)])
