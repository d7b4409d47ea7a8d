"""Composite (boundary-condition) bases and the tensor-product Poisson solve on the GPU (SURVEY §8(f) ranks 1
and 3): parity with the oracle's restatement of composite.py, and the reference's own acceptance criteria
(`examples/poisson2D.py:18-42`, `examples/poisson3D.py:14-35`: error of the manufactured solution < ulp(1000))."""
import numpy as np
import pytest
import sympy as sp
import torch

import jaxfun_oracle as O
import jaxfun_b200 as jf
from jaxfun_b200.galerkin.tpsolve import poisson_solver

pytestmark = pytest.mark.gpu
n = sp.Symbol("n", integer=True)
ULP1000 = 1000 * np.finfo(float).eps
BCS = {"left": {"D": 0}, "right": {"D": 0}}


def dev(x, cuda):
    return torch.from_numpy(np.ascontiguousarray(x)).to(cuda)


def rel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("base", ["Chebyshev", "Legendre"])
@pytest.mark.parametrize("N,dom", [(8, None), (20, None), (64, None), (33, (-2.0, 3.0))])
def test_composite_1d_matches_oracle(cuda, base, N, dom):
    Cp = jf.FunctionSpace(N, getattr(jf, base), BCS, domain=dom, scaling=n + 1)
    Co = O.Composite(N, getattr(O, base), {0: 1, 2: -1}, scaling=n + 1, domain=dom)
    assert Cp.dim == N - 2 == Co.dim
    rng = np.random.default_rng(N)
    c = rng.standard_normal((3, N - 2))
    u_ref = Co.backward(c, axis=-1)
    assert rel(Cp.backward(dev(c, cuda)), u_ref) < 1e-12
    assert rel(Cp.backward(dev(c, cuda), N=N + 6), Co.backward(c, N=N + 6, axis=-1)) < 1e-12
    assert rel(Cp.scalar_product(dev(u_ref, cuda)), Co.scalar_product(u_ref, axis=-1)) < 1e-12
    assert rel(Cp.forward(dev(u_ref, cuda)), Co.forward(u_ref, axis=-1)) < 1e-11
    assert rel(Cp.forward(Cp.backward(dev(c, cuda))), c) < 1e-11          # round trip (test_forward_backward.py:60-74)
    assert rel(Cp.to_orthogonal(dev(c, cuda)), Co.to_orthogonal(c, axis=-1)) < 1e-13
    a = rng.standard_normal((3, N))
    assert rel(Cp.from_orthogonal(dev(a, cuda)), Co.from_orthogonal(a, axis=-1)) < 1e-11
    assert rel(Cp.backward_primitive(dev(c, cuda), 1), Co.backward_primitive(c, 1, axis=-1)) < 1e-11
    # the basis satisfies the boundary conditions
    lo, hi = Cp.domain
    ends = Cp.evaluate(np.array([float(lo), float(hi)]), dev(c, cuda))
    assert float(ends.abs().max()) < 1e-12 * float(np.abs(u_ref).max())


def test_composite_tensor_product_mixed(cuda):
    names = [("Chebyshev", 16), ("Legendre", 12)]
    Tp = jf.TensorProduct(*[jf.FunctionSpace(N, getattr(jf, b), BCS, scaling=n + 1) for b, N in names], jf.Fourier(8))
    To = O.TensorProductSpace(*[O.Composite(N, getattr(O, b), {0: 1, 2: -1}, scaling=n + 1) for b, N in names], O.Fourier(8))
    rng = np.random.default_rng(5)
    c = rng.standard_normal((14, 10, 8)) + 1j * rng.standard_normal((14, 10, 8))
    u_ref = To.backward(c)
    assert rel(Tp.backward(dev(c, cuda)), u_ref) < 1e-12
    assert rel(Tp.forward(dev(u_ref, cuda)), c) < 1e-11
    assert rel(Tp.scalar_product(dev(u_ref, cuda)), To.scalar_product(u_ref)) < 1e-12


@pytest.mark.parametrize("base,M", [("Chebyshev", 20), ("Legendre", 20), ("Chebyshev", 64)])
def test_poisson2d_example(cuda, base, M):
    """examples/poisson2D.py: div grad u = div grad ue on (-1,1)^2, homogeneous Dirichlet, ue = (1-x^2)(1-y^2)
    (C1 of BASELINE.json at M = 64)."""
    D = jf.FunctionSpace(M, getattr(jf, base), BCS, scaling=n + 1, name="D", fun_str="psi")
    T = jf.TensorProduct(D, D, name="T")
    x, y = sp.symbols("x y", real=True)
    ue = (1 - x**2) * (1 - y**2)
    lap = sp.lambdify((x, y), sp.diff(ue, x, 2) + sp.diff(ue, y, 2), "numpy")
    xq = T.mesh()
    b = T.scalar_product(dev(lap(*xq) + 0 * xq[0] * xq[1], cuda))     # (v, div grad ue)_w
    uh = poisson_solver(T).solve(b)
    N = 100
    uj = T.evaluate_mesh(uh, kind="uniform", N=(N, N))
    xj = T.mesh(kind="uniform", N=(N, N))
    uej = sp.lambdify((x, y), ue, "numpy")(*xj)
    error = np.linalg.norm(uj.cpu().numpy() - uej) / N
    assert error < ULP1000, error
    # backward / forward round trip of the solution (BASELINE config 0)
    assert rel(T.forward(T.backward(uh)), uh.cpu().numpy()) < 1e-11


def test_poisson3d_example(cuda):
    """examples/poisson3D.py: Legendre^3, M = 20."""
    M = 20
    D = jf.FunctionSpace(M, jf.Legendre, BCS, scaling=n + 1)
    T = jf.TensorProduct(D, D, D)
    x, y, z = sp.symbols("x y z", real=True)
    ue = (1 - x**2) * (1 - y**2) * (1 - z**2)
    lap = sp.lambdify((x, y, z), sum(sp.diff(ue, s, 2) for s in (x, y, z)), "numpy")
    xq = T.mesh()
    b = T.scalar_product(dev(lap(*xq) + 0 * xq[0] * xq[1] * xq[2], cuda))
    uh = poisson_solver(T).solve(b)
    uj = T.evaluate_mesh(uh, kind="uniform", N=(20, 20, 20))
    xj = T.mesh(kind="uniform", N=(20, 20, 20))
    uej = sp.lambdify((x, y, z), ue, "numpy")(*xj)
    error = np.linalg.norm(uj.cpu().numpy() - uej) / np.sqrt(T.dim)
    assert error < ULP1000, error


@pytest.mark.parametrize("base", ["Chebyshev", "Legendre"])
@pytest.mark.parametrize("bcs", [{"left": {"D": 0, "N": 0}, "right": {"D": 0, "N": 0}}, {"left": {"N": 0}, "right": {"N": 0}},
                                 {"left": {"D": 0}, "right": {"N": 0}}])
def test_composite_general_boundary_conditions(cuda, base, bcs):
    """Numeric stencils (stencil_from_bcs == get_stencil_matrix, composite.py:765-838): round trip and the
    boundary conditions themselves, on the device."""
    N = 24
    C = jf.FunctionSpace(N, getattr(jf, base), bcs)
    nb = sum(len(v) for v in bcs.values())
    assert C.dim == N - nb
    rng = np.random.default_rng(nb)
    c = dev(rng.standard_normal((4, C.dim)), cuda)
    u = C.backward(c)
    assert rel(C.forward(u), c.cpu().numpy()) < 1e-10
    ends = {"left": float(C.domain[0]), "right": float(C.domain[1])}
    for side, kinds in bcs.items():
        for kind in kinds:
            k = {"D": 0, "N": 1}[kind]
            X = np.array([ends[side]])
            T = np.ascontiguousarray(C.evaluate_basis_derivative(X, k))
            val = C._run(jf._lib.OP_APPLY, c, -1, table=T, cache=False)
            assert float(val.abs().max()) < 1e-8 * max(1.0, float(u.abs().max()))
