"""The reference's own known-answer and physics pins, replayed against this implementation
(SURVEY.md §4 / §8c):

  * Vandermonde and derivative Vandermonde vs numpy.polynomial        tests/galerkin/test_approximations.py:48-87
  * backward_primitive vs analytic derivatives on the quadrature mesh tests/galerkin/test_backward_primitive_eval.py:20-60
  * Fourier padding shape + round trip                                tests/galerkin/test_fourier.py:7-16
  * KdV soliton tracked by ETDRK4 (rel. error < 5e-3)                 tests/integrators/test_etdrk4.py:152-198
  * composite <-> orthogonal coefficient maps in 2-D / 3-D            tests/galerkin/test_to_from_orthogonal*.py

Tolerances are the reference's (`ulp(k)` of float64, utils/common.py:103-104)."""
import numpy as np
import pytest
import sympy as sp

import jaxfun_oracle as O


def ulp(x):
    return np.nextafter(x, x + 1) - x


# ---- host tables: CPU ----------------------------------------------------------------------------------
@pytest.mark.parametrize("mod", ["oracle", "product"])
@pytest.mark.parametrize("name", ["Legendre", "Chebyshev"])
@pytest.mark.parametrize("k", [0, 1, 2, 3])
def test_vandermonde_and_derivatives_vs_numpy_polynomial(mod, name, k):
    N = 12
    x = np.linspace(-1, 1, 23)
    vander = {"Legendre": np.polynomial.legendre.legvander, "Chebyshev": np.polynomial.chebyshev.chebvander}[name]
    der = {"Legendre": np.polynomial.legendre.legder, "Chebyshev": np.polynomial.chebyshev.chebder}[name]
    ref = vander(x, N - 1)
    if k > 0:
        D = np.zeros((N, N))
        D[:-k] = der(np.eye(N, N), k)
        ref = ref @ D
    if mod == "oracle":
        V = getattr(O, name)(N)
        got = V.vandermonde(x)
        if k > 0:
            Dm = np.eye(N)
            for _ in range(k):
                Dm = V._derivative1(Dm)
            got = got @ Dm
    else:
        import jaxfun_b200 as jf
        got = getattr(jf, name)(N).evaluate_basis_derivative(x, k)
    assert np.linalg.norm(ref - got) < ulp(10.0 ** (k + 2))


# ---- device ------------------------------------------------------------------------------------------
def _dev(x, cuda):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x)).to(cuda)


@pytest.mark.gpu
def test_fourier_backward_primitive_matches_analytical(cuda):
    import jaxfun_b200 as jf
    N = 32
    V = jf.Fourier(N, domain=(-1, 1))
    xj = np.asarray(V.mesh(), dtype=float)
    uh = V.forward(_dev(np.sin(np.pi * xj) + 0j, cuda))             # project1D = forward of the samples
    for k, f, tol in [(0, np.sin(np.pi * xj), ulp(100.0)), (1, np.pi * np.cos(np.pi * xj), ulp(100.0)),
                      (3, -np.pi**3 * np.cos(np.pi * xj), ulp(10000.0))]:
        got = V.backward_primitive(uh, k=k).cpu().numpy()
        assert np.linalg.norm(got - f) / np.linalg.norm(f) < tol, k


@pytest.mark.gpu
def test_chebyshev_backward_primitive_matches_analytical(cuda):
    import jaxfun_b200 as jf
    N = 36
    V = jf.Chebyshev(N, domain=(0, 2))
    xj = np.asarray(V.mesh(), dtype=float)
    ue = xj**4 - 3 * xj**2 + xj
    uh = V.forward(_dev(ue, cuda))
    for k, f, tol in [(0, ue, ulp(1000.0)), (1, 4 * xj**3 - 6 * xj + 1, ulp(1000.0)), (2, 12 * xj**2 - 6, np.sqrt(ulp(1.0)))]:
        got = V.backward_primitive(uh, k=k).cpu().numpy()
        assert np.linalg.norm(got - f) / np.linalg.norm(f) < tol, k


@pytest.mark.gpu
def test_fourier_padding_shape_and_roundtrip(cuda):
    import jaxfun_b200 as jf
    rng = np.random.default_rng(0)
    V = jf.Fourier(8)
    c = rng.standard_normal(8) + 1j * rng.standard_normal(8)
    u = V.backward(_dev(c, cuda), N=12)
    assert tuple(u.shape) == (12,)
    assert np.abs(V.forward(u).cpu().numpy() - c).max() < ulp(100.0)
    T = jf.TensorProduct(jf.Fourier(8), jf.Fourier(8))
    c2 = rng.standard_normal((8, 8)) + 1j * rng.standard_normal((8, 8))
    u2 = T.backward(_dev(c2, cuda), N=(12, 8))
    assert tuple(u2.shape) == (12, 8)
    assert np.abs(T.forward(u2).cpu().numpy() - c2).max() < ulp(100.0)


@pytest.mark.gpu
def test_etdrk4_kdv_soliton_tracks_exact_short_time(cuda):
    """u_t + u u_x + mu^2 u_xxx = 0 on (-20, 20), N = 64, 20 ETDRK4 steps of 2.5e-4 (test_etdrk4.py:152-198)."""
    import jaxfun_b200 as jf
    from jaxfun_b200.integrators import ETDRK4, NonlinearTerm, field
    N, Lh, mu, cspeed, x0, steps, dt = 64, 20.0, 0.4, 0.5, -5.0, 20, 2.5e-4
    V = jf.Fourier(N, domain=(-Lh, Lh))
    u, (x,) = field(V)
    term = NonlinearTerm(V, -u * u.diff(x))                         # the integrator stores -(nonlinear part)
    k = np.asarray(V.wavenumbers(eliminate_highest_freq=True), dtype=float) * float(V.domain_factor)   # odd order
    Ldiag = -(mu**2) * (1j * k) ** 3
    xj = np.asarray(V.mesh(), dtype=float)
    u0 = 3 * cspeed / np.cosh(0.5 * np.sqrt(cspeed) / mu * (xj - x0)) ** 2
    uh0 = V.forward(_dev(u0 + 0j, cuda))
    integ = ETDRK4(V, linear_diag=_dev(Ldiag, cuda), nonlinear=term)
    uh = integ.solve(uh0, dt, steps)
    u_num = V.backward(uh).cpu().numpy().real
    T = steps * dt
    u_exact = 3 * cspeed / np.cosh(0.5 * np.sqrt(cspeed) / mu * (xj - cspeed * T - x0)) ** 2
    assert np.linalg.norm(u_num - u_exact) / np.linalg.norm(u_exact) < 5e-3


@pytest.mark.gpu
@pytest.mark.parametrize("dims", [2, 3])
def test_to_from_orthogonal_tensor_product(cuda, dims):
    import jaxfun_b200 as jf
    n = sp.Symbol("n", integer=True)
    bcs = {"left": {"D": 0}, "right": {"D": 0}}
    bases = [jf.Chebyshev, jf.Legendre, jf.Chebyshev][:dims]
    obases = [O.Chebyshev, O.Legendre, O.Chebyshev][:dims]
    Ns = [10, 12, 9][:dims]
    T = jf.TensorProduct(*[jf.FunctionSpace(N, b, bcs, scaling=n + 1) for N, b in zip(Ns, bases)])
    Co = [O.Composite(N, b, {0: 1, 2: -1}, scaling=n + 1) for N, b in zip(Ns, obases)]
    rng = np.random.default_rng(dims)
    c = rng.standard_normal([N - 2 for N in Ns])
    ref = c
    for ax, C in enumerate(Co):
        ref = C.to_orthogonal(ref, axis=ax)
    got = T.to_orthogonal(_dev(c, cuda))
    assert tuple(got.shape) == tuple(Ns)
    assert np.abs(got.cpu().numpy() - ref).max() < ulp(100.0)
    back = T.from_orthogonal(got)
    assert np.abs(back.cpu().numpy() - c).max() < 1e-12
