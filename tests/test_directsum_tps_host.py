"""DirectSumTPS (tensor product with one DirectSum factor whose boundary data depend on the other coordinates,
tensorproductspace.py:575-851 of the reference; the case of examples/poisson2D_periodic.py) WITHOUT a GPU: the product's
host-built lift and its transform algebra, run on a numpy stand-in for the engine plans, against the oracle's restatement,
plus the property that defines the space: the expansion takes the prescribed boundary functions whatever the free
coefficients are.  (The reference's own DirectSumTPS needs its flax-based `la` package and cannot run here: parity of this
class is pinned through `get_bc_basis` — tests/test_directsum_host.py — and these properties only.)"""
import numpy as np
import pytest
import sympy as sp

import jaxfun_oracle as O


@pytest.fixture()
def numpy_engine(monkeypatch):
    """Replace the device plans by dense numpy table applications (same tables the engine would get)."""
    import jaxfun_b200 as jf  # noqa: F401
    from jaxfun_b200 import _lib as L
    from jaxfun_b200.galerkin import tensorproductspace as TP
    from jaxfun_b200.galerkin.orthogonal import OrthogonalSpace

    def run(self, op, x, axis, N=None, k=0, table=None, cache=True, name=None):
        x = np.asarray(x)
        axis = axis % x.ndim
        if table is None:
            if op in (L.OP_FORWARD, L.OP_SCALAR_PRODUCT):
                n_quad, n_coeff = x.shape[axis], self.dim
            else:
                n_quad, n_coeff = (self.num_quad_points if N is None else int(N)), x.shape[axis]
            table = self._dense_table(op, n_coeff, n_quad, k)
        return np.moveaxis(np.tensordot(table, x, axes=(1, axis)), 0, axis)

    def tp(self, op, x, N=None, k=None):
        x = np.asarray(x)
        if self.complex_data and not np.iscomplexobj(x):
            x = x.astype(complex)
        for ax, s in enumerate(self.basespaces):
            x = run(s, op, x, ax, None if N is None else N[ax], 0 if k is None else k[ax])
        return x

    monkeypatch.setattr(OrthogonalSpace, "_run", run)
    monkeypatch.setattr(TP.TensorProductSpace, "backward", lambda self, c, N=None: tp(self, L.OP_BACKWARD, c, N))
    monkeypatch.setattr(TP.TensorProductSpace, "forward", lambda self, u: tp(self, L.OP_FORWARD, u))
    monkeypatch.setattr(TP, "as_jfx_array", lambda x, cplx=False: (
        (np.asarray(x).astype(complex) if cplx and not np.iscomplexobj(x) else np.asarray(x)), True))
    return L


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def test_fourier_x_legendre_directsum(numpy_engine):
    import jaxfun_b200 as jf
    x, y = sp.symbols("x y", real=True)
    ue = sp.cos(2 * x) * (1 - y**2) + sp.sin(x) * y + 0.3          # boundary data depend on x
    bcs = {"left": {"D": ue.subs(y, -1)}, "right": {"D": ue.subs(y, 1)}}
    NF, NL = 16, 12
    T = jf.TensorProduct(jf.Fourier(NF), jf.FunctionSpace(NL, jf.Legendre, bcs))
    assert type(T).__name__ == "DirectSumTPS" and T.num_dofs == (NF, NL - 2)
    Fo, Co = O.Fourier(NF), O.Composite(NL, O.Legendre, {0: 1, 2: -1})
    xm = np.asarray(Fo.mesh())
    samples = [sp.lambdify(x, ue.subs(y, s))(xm) + 0 * xm for s in (-1, 1)]
    To = O.DirectSumTPS([Fo, Co], 1, {"left": {"D": 1}, "right": {"D": 1}}, samples)
    assert rel(T.lift, To.lift) < 1e-14
    rng = np.random.default_rng(0)
    c = rng.standard_normal((NF, NL - 2)) + 1j * rng.standard_normal((NF, NL - 2))
    assert rel(T.to_orthogonal(c), To.to_orthogonal(c)) < 1e-14
    assert rel(T.backward(c), To.backward(c)) < 1e-13
    u = To.backward(c)
    assert rel(T.forward(u), To.forward(u)) < 1e-12
    assert rel(T.forward(T.backward(c)), c) < 1e-12
    # the boundary functions are reproduced: evaluate the orthogonal expansion at y = -1 / +1 (P_k(-1) = (-1)^k, P_k(1) = 1)
    a = T.to_orthogonal(c)
    k = np.arange(NL)
    assert np.abs(Fo.backward(a @ ((-1.0) ** k), axis=0) - samples[0]).max() < 1e-12
    assert np.abs(Fo.backward(a @ np.ones(NL), axis=0) - samples[1]).max() < 1e-12
    with pytest.raises(RuntimeError):
        T.scalar_product(u)


def test_three_factors_with_a_homogeneous_composite(numpy_engine):
    import jaxfun_b200 as jf
    x, y = sp.symbols("x y", real=True)
    g0, g1 = sp.sin(x) * (1 - y**2), sp.cos(x) + y
    T = jf.TensorProduct(jf.Fourier(8), jf.FunctionSpace(10, jf.Chebyshev, {"left": {"D": 0}, "right": {"D": 0}}),
                         jf.FunctionSpace(12, jf.Legendre, {"left": {"D": g0}, "right": {"D": g1}}))
    F3, C3, L3 = O.Fourier(8), O.Composite(10, O.Chebyshev, {0: 1, 2: -1}), O.Composite(12, O.Legendre, {0: 1, 2: -1})
    X, Y = np.meshgrid(np.asarray(F3.mesh()), np.asarray(C3.mesh()), indexing="ij")
    samples = [sp.lambdify((x, y), g)(X, Y) + 0 * X for g in (g0, g1)]
    To = O.DirectSumTPS([F3, C3, L3], 2, {"left": {"D": 1}, "right": {"D": 1}}, samples)
    assert rel(T.lift, To.lift) < 1e-14
    rng = np.random.default_rng(1)
    c = rng.standard_normal((8, 8, 10)) + 1j * rng.standard_normal((8, 8, 10))
    assert rel(T.backward(c), To.backward(c)) < 1e-13
    assert rel(T.forward(T.backward(c)), c) < 1e-12


def test_unsupported_layouts_are_refused():
    import jaxfun_b200 as jf
    x = sp.Symbol("x", real=True)
    D = jf.FunctionSpace(8, jf.Legendre, {"left": {"D": sp.sin(x)}, "right": {"D": 0}})
    with pytest.raises(ValueError):
        D.backward(np.zeros(6))                                   # function-valued data need the tensor product
    Dc = jf.FunctionSpace(8, jf.Legendre, {"left": {"D": 1.0}, "right": {"D": 1.0}})
    with pytest.raises(ValueError):                               # tensorproductspace.py:612-615, with two DirectSums as well
        jf.TensorProduct(Dc, jf.Fourier(8), Dc)
    with pytest.raises(NotImplementedError):
        jf.TensorProduct(Dc, Dc, Dc)
    with pytest.raises(ValueError):                               # tensorproductspace.py:612-615
        jf.TensorProduct(jf.FunctionSpace(8, jf.Legendre, {"left": {"D": 1.0}, "right": {"D": 0}}), jf.Fourier(8), jf.Fourier(8))


@pytest.mark.parametrize("space", ["Legendre", "Chebyshev"])
def test_two_inhomogeneous_directions(numpy_engine, space):
    """Both directions carry boundary data (tensorproductspace.py:620-668, 689-747: corner-compatible projected boundary
    conditions): the lift takes the prescribed Dirichlet / Neumann data on all four sides, a low-degree polynomial is
    reproduced exactly, and free coefficients do not move the boundary values."""
    import jaxfun_b200 as jf
    x, y = sp.symbols("x y", real=True)
    ue = (x**2 + 2 * x) * (y**3 - y) + 3 + x - 2 * y + sp.Rational(1, 2) * x * y + y**2 * x**3
    domx, domy = (0.0, 2.0), (-1.0, 1.0)
    bcx = {"left": {"D": ue.subs(x, domx[0])}, "right": {"N": ue.diff(x).subs(x, domx[1])}}          # physical-unit Neumann
    bcy = {"left": {"D": ue.subs(y, domy[0])}, "right": {"D": ue.subs(y, domy[1])}}
    Sp = getattr(jf, space)
    N = 12
    T = jf.TensorProduct(jf.FunctionSpace(N, Sp, bcx, domain=domx), jf.FunctionSpace(N, Sp, bcy, domain=domy))
    assert type(T).__name__ == "DirectSumTPS" and T.num_dofs == (N - 2, N - 2)
    To = T.orthogonal
    # (1) the lift satisfies the four boundary conditions
    ys = np.linspace(-1, 1, 7); xs = np.linspace(0, 2, 7)
    Vx = lambda pts, k=0: np.asarray(To.basespaces[0].evaluate_basis_derivative(np.asarray(To.basespaces[0].map_reference_domain(pts)), k))
    Vy = lambda pts, k=0: np.asarray(To.basespaces[1].evaluate_basis_derivative(np.asarray(To.basespaces[1].map_reference_domain(pts)), k))
    f = sp.lambdify((x, y), ue, "numpy"); fx = sp.lambdify((x, y), ue.diff(x), "numpy")
    left = Vx(np.array([0.0])) @ T.lift @ Vy(ys).T
    assert np.abs(left[0] - f(0.0, ys)).max() < 1e-11
    dfx = float(To.basespaces[0].domain_factor)
    right = dfx * (Vx(np.array([2.0]), 1) @ T.lift @ Vy(ys).T)
    assert np.abs(right[0] - fx(2.0, ys)).max() < 1e-10
    for yb in (-1.0, 1.0):
        row = Vx(xs) @ T.lift @ Vy(np.array([yb])).T
        assert np.abs(row[:, 0] - f(xs, yb)).max() < 1e-11
    # (2) exact reproduction of the polynomial
    X, Y = T.mesh()
    u = f(X, Y)
    c = T.forward(u)
    assert c.shape == (N - 2, N - 2)
    assert rel(T.backward(c), u) < 1e-12
    # (3) free coefficients keep the boundary values: evaluate the full expansion on the boundary
    rng = np.random.default_rng(0)
    cr = rng.standard_normal(c.shape)
    a = T.to_orthogonal(cr)
    assert np.abs((Vx(np.array([0.0])) @ a @ Vy(ys).T)[0] - f(0.0, ys)).max() < 1e-10
    assert np.abs((Vx(xs) @ a @ Vy(np.array([1.0])).T)[:, 0] - f(xs, 1.0)).max() < 1e-10
    assert rel(T.from_orthogonal(a), cr) < 1e-11
    # inconsistent corner data are refused
    bad = {"left": {"D": ue.subs(y, domy[0]) + 1}, "right": {"D": ue.subs(y, domy[1])}}
    with pytest.raises(ValueError):
        jf.TensorProduct(jf.FunctionSpace(N, Sp, bcx, domain=domx), jf.FunctionSpace(N, Sp, bad, domain=domy))


@pytest.mark.parametrize("first", ["Fourier", "Legendre", "Composite"])
def test_two_inhomogeneous_directions_in_3d(numpy_engine, first):
    """(plain, DirectSum, DirectSum) — the 3-D layout the reference builds for two inhomogeneous directions
    (tensorproductspace.py:704-747): the data of both directions depend on the first coordinate as well.  The lift takes the
    prescribed values on all four faces for every x, a function of the space is reproduced, free coefficients keep the faces."""
    import jaxfun_b200 as jf
    x, y, z = sp.symbols("x y z", real=True)
    if first == "Fourier":
        fx, S0 = sp.cos(x) + sp.Rational(1, 2) * sp.sin(2 * x) + 2, jf.Fourier(8)
    elif first == "Legendre":
        fx, S0 = 1 + x - x**3, jf.Legendre(8)
    else:
        fx, S0 = (1 - x**2) * (2 + x), jf.FunctionSpace(8, jf.Legendre, {"left": {"D": 0}, "right": {"D": 0}})
    ue = fx * ((y**2 + y) * (z**3 - z) + 3 - 2 * z + sp.Rational(1, 2) * y * z + z**2 * y**3 + y)
    domy, domz = (0.0, 2.0), (-1.0, 1.0)
    bcy = {"left": {"D": ue.subs(y, domy[0])}, "right": {"N": ue.diff(y).subs(y, domy[1])}}
    bcz = {"left": {"D": ue.subs(z, domz[0])}, "right": {"D": ue.subs(z, domz[1])}}
    N = 10
    T = jf.TensorProduct(S0, jf.FunctionSpace(N, jf.Legendre, bcy, domain=domy), jf.FunctionSpace(N, jf.Chebyshev, bcz, domain=domz))
    assert type(T).__name__ == "DirectSumTPS" and T.num_dofs == (S0.dim, N - 2, N - 2) and T.bc_axis == (1, 2)
    To = T.orthogonal
    V = [lambda pts, k=0, s=s: np.asarray(s.evaluate_basis_derivative(np.asarray(s.map_reference_domain(pts)), k)) for s in To.basespaces]
    xs = np.linspace(0.3, 2.9, 5) if first == "Fourier" else np.linspace(-0.9, 0.8, 5)
    ys, zs = np.linspace(0, 2, 6), np.linspace(-1, 1, 7)
    f = sp.lambdify((x, y, z), ue, "numpy")
    fy = sp.lambdify((x, y, z), ue.diff(y), "numpy")

    def expand(a, X, Y, Z, ky=0):
        return np.einsum("ijk,pi,qj,rk->pqr", a, V[0](X), V[1](Y, ky), V[2](Z))

    def faces_ok(a, tol):
        X, Z = np.meshgrid(xs, zs, indexing="ij")
        assert np.abs(expand(a, xs, np.array([0.0]), zs)[:, 0, :] - f(X, 0.0, Z)).max() < tol
        dfy = float(To.basespaces[1].domain_factor)
        assert np.abs(dfy * expand(a, xs, np.array([2.0]), zs, ky=1)[:, 0, :] - fy(X, 2.0, Z)).max() < 10 * tol
        X, Y = np.meshgrid(xs, ys, indexing="ij")
        for zb in (-1.0, 1.0):
            assert np.abs(expand(a, xs, ys, np.array([zb]))[:, :, 0] - f(X, Y, zb)).max() < tol

    faces_ok(T.lift, 1e-10)
    Xm, Ym, Zm = T.mesh()
    u = f(Xm, Ym, Zm) + 0 * Xm * Ym * Zm
    c = T.forward(u)
    assert c.shape == (S0.dim, N - 2, N - 2)
    assert rel(T.backward(c), u) < 1e-11                           # ue lies in the space: reproduced
    rng = np.random.default_rng(1)
    cr = rng.standard_normal(c.shape) + (1j * rng.standard_normal(c.shape) if first == "Fourier" else 0)
    a = T.to_orthogonal(cr)
    faces_ok(a, 1e-9)                                              # the homogeneous part vanishes on the four faces
    assert rel(T.from_orthogonal(a), cr) < 1e-10
    bad = {"left": {"D": ue.subs(z, domz[0]) + fx}, "right": {"D": ue.subs(z, domz[1])}}
    with pytest.raises(ValueError):
        jf.TensorProduct(S0, jf.FunctionSpace(N, jf.Legendre, bcy, domain=domy), jf.FunctionSpace(N, jf.Chebyshev, bad, domain=domz))


def test_directsumtps_3d_complex_data_as_the_reference_test(numpy_engine):
    """The space of tests/galerkin/test_tensorproductspace_extra.py:48-81 (`test_directsumtps_poisson_3d`): Fourier x Chebyshev x
    Chebyshev with COMPLEX Dirichlet data ue = sin(2x) exp(2y + i z) on both inhomogeneous directions.  Its transform assertions:
    the projection of ue evaluates back to ue (< sqrt(ulp(1))) and forward(backward(uh)) returns uh (< ulp(1000))."""
    import jaxfun_b200 as jf
    x, y, z = sp.symbols("x y z", real=True)
    N = 20
    ue = sp.sin(2 * x) * sp.exp(2 * y + z * sp.I)
    bcsy = {"left": {"D": ue.subs(y, -1)}, "right": {"D": ue.subs(y, 1)}}
    bcsz = {"left": {"D": ue.subs(z, -1)}, "right": {"D": ue.subs(z, 1)}}
    T = jf.TensorProduct(jf.Fourier(N), jf.FunctionSpace(N, jf.Chebyshev, bcsy), jf.FunctionSpace(N + 2, jf.Chebyshev, bcsz))
    assert type(T).__name__ == "DirectSumTPS" and np.iscomplexobj(T.lift)
    X, Y, Z = T.mesh()
    uej = sp.lambdify((x, y, z), ue, "numpy")(X, Y, Z)
    uh = T.forward(uej)
    uj = T.backward(uh)
    eps = np.finfo(float).eps
    assert np.linalg.norm(uj - uej) < np.sqrt(eps)
    assert np.linalg.norm(T.forward(uj) - uh) < 1000 * eps
    # the lift alone already carries the boundary values: on z = +-1 the homogeneous part vanishes
    a = T.to_orthogonal(np.zeros_like(uh))
    k = np.arange(N + 2)
    for zb, vals in ((-1.0, (-1.0) ** k), (1.0, np.ones(N + 2))):
        face = T.orthogonal.basespaces[0].backward(T.orthogonal.basespaces[1].backward(a @ vals, axis=1), axis=0)
        want = sp.lambdify((x, y), ue.subs(z, zb), "numpy")(X[:, :, 0], Y[:, :, 0])
        assert np.abs(face - want).max() < 1e-11


def test_directsumtps_3d_clamped_direction_on_a_mapped_domain(numpy_engine):
    """The space of `test_directsumtps_biharmonic_dirichlet_3d` (test_tensorproductspace_extra.py:84-135): Dirichlet data in y,
    Dirichlet + Neumann data (physical units) on both ends of z, both on the domain (-1/2, 1/2), complex values."""
    import jaxfun_b200 as jf
    x, y, z = sp.symbols("x y z", real=True)
    N = 20
    ue = sp.sin(2 * x) * sp.exp(4 * y + z * sp.I)
    lo, hi = -0.5, 0.5
    bcsy = {"left": {"D": ue.subs(y, lo)}, "right": {"D": ue.subs(y, hi)}}
    bcsz = {"left": {"D": ue.subs(z, lo), "N": ue.diff(z, 1).subs(z, lo)}, "right": {"D": ue.subs(z, hi), "N": ue.diff(z, 1).subs(z, hi)}}
    T = jf.TensorProduct(jf.Fourier(N), jf.FunctionSpace(N, jf.Chebyshev, bcsy, domain=(lo, hi)),
                         jf.FunctionSpace(N + 2, jf.Chebyshev, bcsz, domain=(lo, hi)))
    assert T.num_dofs == (N, N - 2, N - 2)
    X, Y, Z = T.mesh()
    uej = sp.lambdify((x, y, z), ue, "numpy")(X, Y, Z)
    uh = T.forward(uej)
    uj = T.backward(uh)
    eps = np.finfo(float).eps
    assert np.linalg.norm(uj - uej) < np.sqrt(eps)
    assert np.linalg.norm(T.forward(uj) - uh) < 1000 * eps


def test_directsumtps_evaluate_has_no_host_route():
    """`DirectSumTPS.evaluate` = the orthogonal product's scattered evaluation of the lifted coefficients; like every
    compute entry point it needs the device."""
    import torch
    import jaxfun_b200 as jf
    from jaxfun_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    x, y = sp.symbols("x y", real=True)
    ue = sp.exp(-(x**2 + y**2))                                   # test_tensorproductspace_more.py:83-99
    bcsx = {"left": {"D": ue.subs(x, 0)}, "right": {"D": ue.subs(x, 1)}}
    bcsy = {"left": {"D": ue.subs(y, 0)}, "right": {"D": ue.subs(y, 1)}}
    T = jf.TensorProduct(jf.FunctionSpace(16, jf.Legendre, bcsx, domain=(0, 1)), jf.FunctionSpace(16, jf.Legendre, bcsy, domain=(0, 1)))
    with pytest.raises(_lib.JfxError) as e:
        T.evaluate(np.array([0.5, 0.5]), np.zeros(T.num_dofs))
    assert e.value.code == -3


def test_directsum_two_inhomogeneous_point_value(numpy_engine):
    """tests/galerkin/test_tensorproductspace_more.py:83-99: the projection of exp(-(x^2 + y^2)) onto a space with Dirichlet data on
    all four sides of (0, 1)^2, evaluated at (1/2, 1/2) through the orthogonal expansion of the lifted coefficients, < ulp(100)."""
    import jaxfun_b200 as jf
    x, y = sp.symbols("x y", real=True)
    ue = sp.exp(-(x**2 + y**2))
    N = 16
    bcsx = {"left": {"D": ue.subs(x, 0)}, "right": {"D": ue.subs(x, 1)}}
    bcsy = {"left": {"D": ue.subs(y, 0)}, "right": {"D": ue.subs(y, 1)}}
    T = jf.TensorProduct(jf.FunctionSpace(N, jf.Legendre, bcsx, domain=(0, 1)), jf.FunctionSpace(N, jf.Legendre, bcsy, domain=(0, 1)))
    X, Y = T.mesh()
    uf = T.forward(sp.lambdify((x, y), ue, "numpy")(X, Y))
    a = T.to_orthogonal(uf)
    V = [np.asarray(s.eval_basis_functions(np.asarray(s.map_reference_domain(np.array([0.5]))))) for s in T.orthogonal.basespaces]
    u0 = (V[0] @ a @ V[1].T)[0, 0]
    assert abs(u0 - float(ue.subs({x: 0.5, y: 0.5}))) < 100 * np.finfo(float).eps
