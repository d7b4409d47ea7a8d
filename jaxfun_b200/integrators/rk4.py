"""Classical RK4 in coefficient space — `jaxfun.integrators.rk4.RK4.step`
(`src/jaxfun/integrators/rk4.py:14-20`) with the stage combinations fused into single
`jfx_axpby_diag` launches."""
from .base import BaseIntegrator, axpby_diag


class RK4(BaseIntegrator):
    def step(self, u_hat, dt: float, N=None):
        k1 = self.total_rhs(u_hat, N)
        k2 = self.total_rhs(axpby_diag([(1.0, None, u_hat), (0.5 * dt, None, k1)]), N)
        k3 = self.total_rhs(axpby_diag([(1.0, None, u_hat), (0.5 * dt, None, k2)]), N)
        k4 = self.total_rhs(axpby_diag([(1.0, None, u_hat), (dt, None, k3)]), N)
        return axpby_diag([(1.0, None, u_hat), (dt / 6.0, None, k1), (dt / 3.0, None, k2),
                           (dt / 3.0, None, k3), (dt / 6.0, None, k4)])
