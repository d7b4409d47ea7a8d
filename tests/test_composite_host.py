"""CPU checks of the composite-basis restatement (oracle) and of the host side of the tensor-product
Poisson solve (stiffness / mass assembly, per-axis diagonalisation) — no device needed."""
import numpy as np
import pytest
import sympy as sp

import jaxfun_oracle as O

n = sp.Symbol("n", integer=True)


@pytest.mark.parametrize("base", [O.Chebyshev, O.Legendre])
@pytest.mark.parametrize("N", [8, 20, 33])
def test_oracle_composite_identities(base, N):
    C = O.Composite(N, base, {0: 1, 2: -1}, scaling=n + 1)
    rng = np.random.default_rng(N)
    c = rng.standard_normal(C.dim)
    u = C.backward(c)
    assert np.abs(C.forward(u) - c).max() < 1e-12                      # round trip
    assert np.abs(C.evaluate(np.array([-1.0, 1.0]), c)).max() < 1e-13   # homogeneous Dirichlet
    # scalar_product == S @ orthogonal.scalar_product ; forward == mass^-1 scalar_product
    assert np.allclose(C.scalar_product(u), C.S @ C.orthogonal.scalar_product(u), rtol=0, atol=1e-14)
    assert np.allclose(C.mass @ C.forward(u), C.scalar_product(u), rtol=0, atol=1e-13)
    # to/from orthogonal are inverse on the range of S^T
    assert np.abs(C.from_orthogonal(C.to_orthogonal(c)) - c).max() < 1e-12


def test_host_poisson_assembly_and_diagonalisation():
    import jaxfun_b200 as jf
    from jaxfun_b200.galerkin.tpsolve import KroneckerSumSolver, mass_matrix, stiffness_matrix
    for base in (jf.Chebyshev, jf.Legendre):
        D = jf.FunctionSpace(20, base, {"left": {"D": 0}, "right": {"D": 0}}, scaling=n + 1)
        A, B = stiffness_matrix(D, 2), mass_matrix(D)
        assert np.abs(B - D.mass_matrix()).max() < 1e-13
        # (phi_i, phi_j'')_w against the oracle's basis: u = sum_j c_j phi_j  ->  (phi_i, u'')_w = (A c)_i
        Co = O.Composite(20, getattr(O, base.__name__), {0: 1, 2: -1}, scaling=n + 1)
        rng = np.random.default_rng(1)
        c = rng.standard_normal(D.dim)
        upp = Co.backward_primitive(c, 2)
        assert np.abs(Co.scalar_product(upp) - A @ c).max() < 1e-9 * np.abs(A @ c).max()
        S = KroneckerSumSolver([(A, B), (A, B)])
        u = rng.standard_normal((D.dim, D.dim))
        f = A @ u @ B.T + B @ u @ A.T
        u2 = S.V[0] @ ((S.W[0] @ f @ S.W[1].T) * S.Dinv) @ S.V[1].T
        assert np.abs(u2 - u).max() < 1e-9 * np.abs(u).max()


@pytest.mark.parametrize("base", ["Chebyshev", "Legendre"])
def test_numeric_stencils_match_the_reference_closed_forms(base):
    """get_stencil_matrix special cases (composite.py:783-797) and the Neumann example of its docstring."""
    import jaxfun_b200 as jf
    from jaxfun_b200.galerkin.composite import stencil_from_bcs
    N = 24
    orth = getattr(jf, base)(N)
    k = np.arange(N)
    st = stencil_from_bcs({"left": {"D": 0}, "right": {"D": 0}}, orth)
    assert np.allclose(st[0], 1) and np.allclose(st[2], -1) and (1 not in st or np.allclose(st[1], 0))
    st = stencil_from_bcs({"left": {"D": 0, "N": 0}, "right": {"D": 0, "N": 0}}, orth)
    kk = k[: N - 4].astype(float)
    if base == "Legendre":
        d2, d4 = 2 * (-2 * kk - 5) / (2 * kk + 7), (2 * kk + 3) / (2 * kk + 7)
    else:
        d2, d4 = 2 * (-kk - 2) / (kk + 3), (kk + 1) / (kk + 3)
    assert np.abs(st[2] - d2).max() < 1e-10 and np.abs(st[4] - d4).max() < 1e-10
    assert all(np.abs(st.get(j, 0)).max() < 1e-10 for j in (1, 3))
    if base == "Chebyshev":
        st = stencil_from_bcs({"left": {"N": 0}, "right": {"N": 0}}, orth)
        kk = k[: N - 2].astype(float)
        assert np.abs(st[2] + kk**2 / (kk + 2) ** 2).max() < 1e-10
    # the composite basis built from a numeric stencil satisfies its boundary conditions
    C = jf.FunctionSpace(N, getattr(jf, base), {"left": {"D": 0, "N": 0}, "right": {"D": 0, "N": 0}})
    assert C.dim == N - 4
    ends = np.array([-1.0, 1.0])
    assert np.abs(C.eval_basis_functions(ends)).max() < 1e-10
    assert np.abs(C.evaluate_basis_derivative(ends, 1)).max() < 1e-7
