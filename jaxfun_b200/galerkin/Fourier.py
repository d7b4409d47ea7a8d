"""Fourier space exp(i k x) — mirrors `jaxfun.galerkin.Fourier.Fourier`
(`src/jaxfun/galerkin/Fourier.py:14-235`): wavenumber ordering, mid-spectrum zero padding in
`backward`, wavenumber-gather truncation in `forward` / `scalar_product`, (i k)^m derivatives.
Runs as the engine's FFT kernels at supported lengths, otherwise as a dense complex DFT table
(exact integer argument reduction)."""
from __future__ import annotations

import numpy as np

from .. import _lib as L
from .orthogonal import Domain, OrthogonalSpace


def fourier_wavenumbers(N: int, eliminate_highest_freq: bool = False) -> np.ndarray:
    """Integer wavenumbers in FFT order (Fourier.py:14-20)."""
    indices = np.arange(N)
    k = np.where(indices < (N + 1) // 2, indices, indices - N)
    if eliminate_highest_freq and N % 2 == 0:
        k[N // 2] = 0
    return k


class Fourier(OrthogonalSpace):
    fast_basis = L.BASIS_FOURIER
    complex_data = True

    def __init__(self, N: int, domain=None, system=None, name: str = "Fourier", fun_str: str = "E") -> None:
        assert N % 2 == 0, "Fourier must use an even number of modes"
        domain = Domain(0, 2 * np.pi) if domain is None else domain
        OrthogonalSpace.__init__(self, N, domain=domain, system=system, name=name, fun_str=fun_str)

    @property
    def reference_domain(self) -> Domain:
        return Domain(0, 2 * np.pi)

    def _table_dtype(self):
        return np.complex128

    def wavenumbers(self, N: int | None = None, eliminate_highest_freq: bool = False) -> np.ndarray:
        N = self.N if N is None else N
        return fourier_wavenumbers(N, bool(eliminate_highest_freq))

    def quad_points_and_weights(self, N: int | None = None):
        N = self.num_quad_points if N is None else N
        points = np.arange(N, dtype=float) * 2 * np.pi / N
        return points, np.full(N, 2 * np.pi / N)

    def mesh(self, kind: str = "quadrature", N: int | None = None) -> np.ndarray:
        a, b = self.domain
        N = self.num_quad_points if N is None else N
        return np.linspace(float(a), float(b), N, endpoint=False)

    def eval_basis_functions(self, X) -> np.ndarray:
        X = np.atleast_1d(np.asarray(X, dtype=float))
        return np.exp(1j * self.wavenumbers()[None, :] * X[:, None])

    def evaluate_basis_derivative(self, X, k: int = 0) -> np.ndarray:
        v = self.wavenumbers(eliminate_highest_freq=bool(k % 2 and self.N % 2 == 0))
        return (1j * v) ** k * self.eval_basis_functions(X)

    def norm_squared(self) -> np.ndarray:
        return np.ones(self.N) * 2 * np.pi

    def derivative_scale(self, k: int, n: int | None = None) -> np.ndarray:
        """(i m)^k with the Nyquist mode removed for odd k (Fourier.py:206-219)."""
        m = self.wavenumbers(n, eliminate_highest_freq=(k % 2 == 1))
        return (1j * m) ** k

    def _derivative_host(self, c: np.ndarray) -> np.ndarray:
        return self.derivative_scale(1, c.shape[0]).reshape((-1,) + (1,) * (c.ndim - 1)) * c

    def derivative_matrix(self, k: int = 1, n: int | None = None) -> np.ndarray:
        n = self.N if n is None else n
        return np.diag(self.derivative_scale(k, n))

    def _dense_table(self, op: int, n_coeff: int, n_quad: int, deriv: int) -> np.ndarray:
        key = ("T", op, n_coeff, n_quad, deriv)
        T = self._tables.get(key)
        if T is not None:
            return T
        n = n_quad
        j = np.arange(n)

        def expi(num):  # exp(2 pi i num / n) with exact reduction
            r = np.mod(num, n).astype(np.float64)
            return np.cos(2 * np.pi * r / n) + 1j * np.sin(2 * np.pi * r / n)

        if op in (L.OP_FORWARD, L.OP_SCALAR_PRODUCT):
            # out = fft(u)/n, truncated by out[wavenumbers()] when n > N     (Fourier.py:150-180)
            m = self.wavenumbers() if n > self.N else np.arange(n)
            T = expi(-np.outer(m, j)) / n
            if op == L.OP_SCALAR_PRODUCT:
                T = T * (2 * np.pi / float(self.domain_factor))
        else:
            # pad [c[:L/2], 0.., c[L/2:]] then ifft(norm="forward")          (Fourier.py:126-148)
            Lc = n_coeff
            p = np.arange(Lc)
            m = np.where(p < Lc // 2, p, p - Lc) if n > Lc else p
            T = expi(np.outer(j, m))
            if deriv:
                T = T * ((float(self.domain_factor) ** deriv) * self.derivative_scale(deriv, Lc))[None, :]
        T = np.ascontiguousarray(T)
        self._tables[key] = T
        return T
