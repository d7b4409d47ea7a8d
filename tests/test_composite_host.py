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
