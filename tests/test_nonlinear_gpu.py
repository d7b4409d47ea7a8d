"""Nonlinear-term evaluation and integrator stages on the GPU against the oracle
(reference pins: tests/integrators/test_backward_euler.py:118-159, 298-319; test_etdrk4.py:262-315)."""
import numpy as np
import pytest
import sympy as sp
import torch

import jaxfun_oracle as O
import jaxfun_b200 as jf
from jaxfun_b200.integrators import ETDRK4, RK4, NonlinearTerm, field

pytestmark = pytest.mark.gpu


def relerr(a, b):
    a = a.detach().cpu().numpy()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def dev(x, cuda):
    return torch.from_numpy(np.ascontiguousarray(x)).to(cuda)


def crand(rng, shape, scale=1.0):
    return scale * (rng.standard_normal(shape) + 1j * rng.standard_normal(shape))


@pytest.mark.parametrize("N", [16, 64, 30])
def test_burgers_1d_fourier(cuda, N):
    """nonlinear_rhs(uh) == V.forward(-(u * u_x))   (test_backward_euler.py:118-143)."""
    rng = np.random.default_rng(N)
    dom = (0.0, 2.0)
    V, Vo = jf.Fourier(N, domain=dom), O.Fourier(N, domain=dom)
    u, (x,) = field(V)
    term = NonlinearTerm(V, -u * u.diff(x))
    uh = crand(rng, (N,), 0.1)
    ref = O.nonlinear_rhs(Vo, [0, 1], lambda a, ax: -(a * ax), uh)
    assert relerr(term(dev(uh, cuda)), ref) < 1e-12
    # explicit composition through the public transforms (the reference's call-through points)
    comp = V.forward(-(V.backward(dev(uh, cuda)) * V.backward_primitive(dev(uh, cuda), k=1)))
    assert relerr(comp, ref) < 1e-12


def test_square_of_sum_and_padding(cuda):
    """(u + u_x)^2 (test_backward_euler.py:159) with 3/2 padding forwarded as N (ibid :207-243)."""
    rng = np.random.default_rng(1)
    N, M = 32, 48
    V, Vo = jf.Fourier(N), O.Fourier(N)
    u, (x,) = field(V)
    term = NonlinearTerm(V, (u + u.diff(x)) ** 2, N=M)
    uh = crand(rng, (N,), 0.1)
    ref = O.nonlinear_rhs(Vo, [0, 1], lambda a, ax: (a + ax) ** 2, uh, N=M)
    assert relerr(term(dev(uh, cuda)), ref) < 1e-12


def test_2d_advection(cuda):
    """u u_x + u u_y on Fourier x Fourier (test_backward_euler.py:298-319)."""
    rng = np.random.default_rng(2)
    N = (16, 32)
    T = jf.TensorProduct(jf.Fourier(N[0]), jf.Fourier(N[1]))
    To = O.TensorProductSpace(O.Fourier(N[0]), O.Fourier(N[1]))
    u, (x, y) = field(T)
    term = NonlinearTerm(T, u * u.diff(x) + u * u.diff(y))
    uh = crand(rng, N, 0.1)
    ref = O.nonlinear_rhs(To, [(0, 0), (1, 0), (0, 1)], lambda a, ax, ay: a * ax + a * ay, uh)
    assert relerr(term(dev(uh, cuda)), ref) < 1e-12


def test_cahn_hilliard_compact_equals_expanded(cuda):
    """-Laplace(u^3) is expanded by the product rule into 6u(ux^2+uy^2)+3u^2(uxx+uyy)
    (test_etdrk4.py:262-315, examples/cahn_hilliard2D_etdrk4.py:86)."""
    rng = np.random.default_rng(3)
    N = (32, 32)
    dom = (0.0, 1.0)
    T = jf.TensorProduct(jf.Fourier(N[0], domain=dom), jf.Fourier(N[1], domain=dom))
    To = O.TensorProductSpace(O.Fourier(N[0], domain=dom), O.Fourier(N[1], domain=dom))
    u, (x, y) = field(T)
    compact = NonlinearTerm(T, -((u**3).diff(x, 2) + (u**3).diff(y, 2)))
    expanded = NonlinearTerm(T, -(6 * u * (u.diff(x) ** 2 + u.diff(y) ** 2) + 3 * u**2 * (u.diff(x, 2) + u.diff(y, 2))))
    uh = crand(rng, N, 1e-2)
    ref = O.nonlinear_rhs(To, [(0, 0), (1, 0), (0, 1), (2, 0), (0, 2)],
                          lambda a, ax, ay, axx, ayy: -(6 * a * (ax**2 + ay**2) + 3 * a**2 * (axx + ayy)), uh)
    r1, r2 = compact(dev(uh, cuda)), expanded(dev(uh, cuda))
    assert relerr(r1, ref) < 1e-11 and relerr(r2, ref) < 1e-11
    assert sorted(compact.compiled.leaves) == sorted(expanded.compiled.leaves)
    assert len(compact.compiled.leaves) == 5   # u, u_x, u_y, u_xx, u_yy evaluated once each


def test_complex_cubic_and_abs(cuda):
    """(1+1.5j) u |u|^2  (examples/ginzburg_landau_imex.py:50) and -i |psi|^2 psi (examples/nls1D_etdrk4.py:76)."""
    rng = np.random.default_rng(4)
    N = 64
    V, Vo = jf.Fourier(N), O.Fourier(N)
    u, (x,) = field(V)
    uh = crand(rng, (N,), 0.2)
    for expr, fn in [((1 + 1.5j) * u * sp.Abs(u) ** 2, lambda a: (1 + 1.5j) * a * np.abs(a) ** 2),
                     (-sp.I * sp.Abs(u) ** 2 * u, lambda a: -1j * np.abs(a) ** 2 * a)]:
        term = NonlinearTerm(V, expr)
        assert relerr(term(dev(uh, cuda)), O.nonlinear_rhs(Vo, [0], fn, uh)) < 1e-12


@pytest.mark.parametrize("name", ["Chebyshev", "Legendre"])
def test_polynomial_space_scalar_product_leg(cuda, name):
    """u u_x on a polynomial space with the scalar_product final leg (IMEX / Petrov-Galerkin route,
    integrators/base.py:238-248; tests/integrators/test_petrov_galerkin.py:49)."""
    rng = np.random.default_rng(5)
    N = 24
    V = getattr(jf, name)(N, domain=(-1.0, 2.0))
    Vo = getattr(O, name)(N, domain=(-1.0, 2.0))
    u, (x,) = field(V)
    uh = rng.standard_normal((N,)) / (1 + np.arange(N)) ** 2
    for final in ("forward", "scalar_product"):
        term = NonlinearTerm(V, -u * u.diff(x), final=final)
        ref = O.nonlinear_rhs(Vo, [0, 1], lambda a, ax: -(a * ax), uh, final=final)
        assert relerr(term(dev(uh, cuda)), ref) < 1e-11


def test_functions_and_static_coefficient(cuda):
    """General functions (utils/common.py:36-55) and mesh-sampled static factors (nonlinear.py:219-242)."""
    rng = np.random.default_rng(6)
    N = 32
    V, Vo = jf.Chebyshev(N), O.Chebyshev(N)
    u, (x,) = field(V)
    uh = rng.standard_normal((N,)) / (1 + np.arange(N)) ** 2
    term = NonlinearTerm(V, sp.sin(x) * u**2 + sp.exp(u) - 2 * sp.cos(u) / (1 + u**2))
    xj = Vo.mesh()
    ref = Vo.forward(np.sin(xj) * Vo.backward(uh) ** 2 + np.exp(Vo.backward(uh)) - 2 * np.cos(Vo.backward(uh)) / (1 + Vo.backward(uh) ** 2))
    assert relerr(term(dev(uh, cuda)), ref) < 1e-12


def test_leading_batch_of_fields(cuda):
    rng = np.random.default_rng(7)
    N = 32
    V, Vo = jf.Fourier(N), O.Fourier(N)
    u, (x,) = field(V)
    term = NonlinearTerm(V, -u * u.diff(x))
    uh = crand(rng, (5, N), 0.1)
    ref = O.nonlinear_rhs(Vo, [0, 1], lambda a, ax: -(a * ax), uh)
    assert relerr(term(dev(uh, cuda)), ref) < 1e-12


def _kdv_setup(N, cuda):
    dom = (-np.pi, np.pi)
    V, Vo = jf.Fourier(N, domain=dom), O.Fourier(N, domain=dom)
    k = Vo.wavenumbers().astype(float) * float(Vo.domain_factor)
    Ldiag = 1j * k**3                               # u_t + u u_x + u_xxx = 0  ->  L = -(ik)^3 = i k^3
    u, (x,) = field(V)
    term = NonlinearTerm(V, -u * u.diff(x))
    Nfun = lambda uh: O.nonlinear_rhs(Vo, [0, 1], lambda a, ax: -(a * ax), uh)  # noqa: E731
    xj = Vo.mesh()
    u0 = Vo.forward(0.5 / np.cosh(0.5 * xj) ** 2 + 0j)
    return V, Vo, Ldiag, term, Nfun, u0


def test_etdrk4_step_matches_oracle(cuda):
    """One and several ETDRK4 steps of KdV (etdrk4.py:152-166 stage arithmetic)."""
    N = 64
    V, Vo, Ldiag, term, Nfun, u0 = _kdv_setup(N, cuda)
    dt = 1e-3
    integ = ETDRK4(V, linear_diag=dev(Ldiag, cuda), nonlinear=term)
    coeffs = O.etdrk4_coefficients(dt, Ldiag)
    u_ref = u0
    for _ in range(5):
        u_ref = O.etdrk4_step(u_ref, dt, coeffs, Nfun)
    u_gpu = integ.solve(dev(u0, cuda), dt, 5)
    assert relerr(u_gpu, u_ref) < 1e-11


def test_rk4_step_matches_oracle(cuda):
    N = 32
    V, Vo, Ldiag, term, Nfun, u0 = _kdv_setup(N, cuda)
    dt = 1e-4
    integ = RK4(V, linear_diag=dev(Ldiag, cuda), nonlinear=term)
    u_ref = u0
    for _ in range(4):
        u_ref = O.rk4_step(u_ref, dt, lambda uh: Ldiag * uh + Nfun(uh))
    assert relerr(integ.solve(dev(u0, cuda), dt, 4), u_ref) < 1e-11


def test_etdrk4_cahn_hilliard_mass_conservation(cuda):
    """Mean of the Cahn-Hilliard field is conserved (examples/cahn_hilliard2D_etdrk4.py:125-129: drift < 1e-5)."""
    rng = np.random.default_rng(8)
    N = 32
    dom = (0.0, 1.0)
    T = jf.TensorProduct(jf.Fourier(N, domain=dom), jf.Fourier(N, domain=dom))
    To = O.TensorProductSpace(O.Fourier(N, domain=dom), O.Fourier(N, domain=dom))
    kx = To.basespaces[0].wavenumbers().astype(float) * float(To.basespaces[0].domain_factor)
    K2 = kx[:, None] ** 2 + kx[None, :] ** 2
    gamma = 1e-4
    Ldiag = (K2 - gamma * K2**2 * 1.0).astype(complex) * 0 + (K2 - gamma * K2**2)  # u_t = lap(u^3 - u - gamma lap u)
    u, (x, y) = field(T)
    term = NonlinearTerm(T, (u**3).diff(x, 2) + (u**3).diff(y, 2))
    u0 = To.forward(1e-2 * rng.standard_normal((N, N)) + 0j)
    integ = ETDRK4(T, linear_diag=dev(Ldiag.astype(complex), cuda), nonlinear=term)
    u1 = integ.solve(dev(u0, cuda), 1e-6, 8)
    assert abs(complex(u1[0, 0].cpu()) - u0[0, 0]) < 1e-5
    assert bool(torch.isfinite(torch.view_as_real(u1)).all())


def _kdv_weak_setup(N, cuda):
    """KdV-like semilinear system on Fourier(N): u_t = -u_xxx - (u^2/2)_x (weak form, diagonal operators)."""
    dom = (0.0, 2 * np.pi)
    V, Vo = jf.Fourier(N, domain=dom), O.Fourier(N, domain=dom)
    u, (x,) = field(V)
    term = NonlinearTerm(V, -u * u.diff(x))
    k = np.asarray(Vo.wavenumbers(), dtype=float)
    Ldiag = -(1j * k) ** 3
    M = np.full(N, 2 * np.pi)                      # Fourier mass: h_k / df
    Nsp = lambda a: O.nonlinear_rhs(Vo, [0, 1], lambda p, q: -(p * q), a, final="scalar_product")
    return V, term, Ldiag, M, Nsp


@pytest.mark.parametrize("tab", ["IMEX_EULER", "ARS222", "ARS443"])
def test_imex_rk_step_matches_oracle(cuda, tab):
    """IMEXRungeKutta.step (imex_rk.py:100-159) on the GPU against the oracle's restatement."""
    from jaxfun_b200.integrators import IMEXRungeKutta
    import jaxfun_b200.integrators as I
    rng = np.random.default_rng(7)
    N = 32
    V, term, Ldiag, M, Nsp = _kdv_weak_setup(N, cuda)
    tableau = getattr(I, tab)
    integ = IMEXRungeKutta(V, dev(Ldiag, cuda), term, tableau=tableau)
    uh = crand(rng, (N,), 0.05)
    dt = 1e-3
    got = integ.step(dev(uh, cuda), dt)
    ref = O.imex_rk_step(uh, dt, tableau, M, M * Ldiag, Nsp)
    assert relerr(got, ref) < 1e-12
    # a short solve stays finite and equals repeated oracle steps
    got3 = integ.solve(dev(uh, cuda), dt, 3)
    r = uh
    for _ in range(3):
        r = O.imex_rk_step(r, dt, tableau, M, M * Ldiag, Nsp)
    assert relerr(got3, r) < 1e-11


def test_backward_euler_step_matches_oracle(cuda):
    """BackwardEuler.step (backward_euler.py:29-39)."""
    from jaxfun_b200.integrators import BackwardEuler
    rng = np.random.default_rng(8)
    N = 32
    V, term, Ldiag, M, Nsp = _kdv_weak_setup(N, cuda)
    integ = BackwardEuler(V, dev(Ldiag, cuda), term)
    uh = crand(rng, (N,), 0.05)
    dt = 1e-3
    assert relerr(integ.step(dev(uh, cuda), dt), O.backward_euler_step(uh, dt, M, M * Ldiag, Nsp)) < 1e-12


@pytest.mark.parametrize("N,M", [(32, 48), (64, 96), (128, 192)])
def test_three_halves_rule_padding_is_row_fused(cuda, N, M):
    """3/2-rule de-aliasing: N modes evaluated on M = 3N/2 points (a 3 * 2^m transform length) — the padded
    nonlinear term runs as ONE row-fused launch and matches the oracle (nonlinear_rhs(uh, N=M), base.py:230-236)."""
    rng = np.random.default_rng(N)
    V, Vo = jf.Fourier(N), O.Fourier(N)
    u, (x,) = field(V)
    term = NonlinearTerm(V, -u * u.diff(x), N=(M,))
    uh = crand(rng, (6, N), 0.1)
    got = term(dev(uh, cuda))
    assert term.launches(dev(uh, cuda)) == 1
    ref = np.stack([O.nonlinear_rhs(Vo, [0, 1], lambda a, ax: -(a * ax), r, N=M) for r in uh])
    assert relerr(got, ref) < 1e-12


def test_integrator_forwards_N_to_the_evaluator(cuda):
    """`nonlinear_rhs(uh, N)` / `nonlinear_rhs_scalar_product(uh, N)` evaluate at the padded resolution N
    (base.py:230-248, pinned by tests/integrators/test_backward_euler.py:207-243); N = None is the space's own shape."""
    from jaxfun_b200.integrators.imex_rk import BackwardEuler
    rng = np.random.default_rng(8)
    N, M = 32, 48
    V, Vo = jf.Fourier(N), O.Fourier(N)
    u, (x,) = field(V)
    term = NonlinearTerm(V, -u * u.diff(x))
    uh = crand(rng, (N,), 0.1)
    k = Vo.wavenumbers().astype(float)
    Ld = dev(-(k**2) + 0j, cuda)
    for integ in (RK4(V, linear_diag=Ld, nonlinear=term), ETDRK4(V, linear_diag=Ld, nonlinear=term)):
        ref_pad = O.nonlinear_rhs(Vo, [0, 1], lambda a, ax: -(a * ax), uh, N=M)
        ref = O.nonlinear_rhs(Vo, [0, 1], lambda a, ax: -(a * ax), uh)
        assert relerr(integ.nonlinear_rhs(dev(uh, cuda), M), ref_pad) < 1e-12
        assert relerr(integ.nonlinear_rhs(dev(uh, cuda)), ref) < 1e-12
        assert np.abs(ref_pad - ref).max() > 1e-6 * np.abs(ref).max()      # padding really changes the (aliased) result
        tot = integ.total_rhs(dev(uh, cuda), M).cpu().numpy()
        assert np.abs(tot - (ref_pad + (-(k**2)) * uh)).max() < 1e-12 * np.abs(tot).max()
    be = BackwardEuler(V, linear_diag=Ld, nonlinear=term)
    be.setup(1e-3)
    ref_sp = O.nonlinear_rhs(Vo, [0, 1], lambda a, ax: -(a * ax), uh, N=M, final="scalar_product")
    assert relerr(be.nonlinear_rhs_scalar_product(dev(uh, cuda), M), ref_sp) < 1e-12


def test_stage_arithmetic_in_single_precision_casts_its_coefficients(cuda):
    """ETDRK4 coefficients are built in float64 / complex128; a complex64 state must see them as complex64
    (jfx_axpby_diag reads coefficients in the state's precision)."""
    from jaxfun_b200.integrators.base import axpby_diag
    rng = np.random.default_rng(9)
    n = 1000
    x = crand(rng, (n,)).astype(np.complex64)
    y = crand(rng, (n,)).astype(np.complex64)
    cz = crand(rng, (n,))                               # complex128 coefficient
    cr = rng.standard_normal(n)                         # float64 coefficient
    got = axpby_diag([(0.5, dev(cz, cuda), dev(x, cuda)), (2.0, dev(cr, cuda), dev(y, cuda))]).cpu().numpy()
    ref = 0.5 * cz * x + 2.0 * cr * y
    assert got.dtype == np.complex64 and np.abs(got - ref).max() < 1e-5 * np.abs(ref).max()
    xr, yr = rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)
    got = axpby_diag([(1.0, dev(cr, cuda), dev(xr, cuda)), (-1.0, None, dev(yr, cuda))]).cpu().numpy()
    assert np.abs(got - (cr * xr - yr)).max() < 1e-5
    with pytest.raises(TypeError):
        axpby_diag([(1.0, dev(cz, cuda), dev(xr, cuda))])
