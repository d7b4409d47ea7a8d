"""Host-side tables of the product (nodes, weights, wavenumbers, transform tables) against the
oracle restatement: bit-exact for index / quadrature-node tables, 1e-13 for derived tables."""
import numpy as np
import pytest

import jaxfun_oracle as O
import jaxfun_b200 as jf
from jaxfun_b200 import _lib as L

CASES = [
    ("Legendre", O.Legendre, jf.Legendre, {}),
    ("Chebyshev", O.Chebyshev, jf.Chebyshev, {}),
    ("ChebyshevU", O.ChebyshevU, jf.ChebyshevU, {}),
    ("Fourier", O.Fourier, jf.Fourier, {}),
    ("Jacobi(1,2)", O.Jacobi, jf.Jacobi, dict(alpha=1, beta=2)),
    ("Ultraspherical(1.5)", O.Ultraspherical, jf.Ultraspherical, dict(lambda_=1.5)),
]


@pytest.mark.parametrize("N", [1, 2, 3, 8, 9, 33, 64, 100, 101, 128, 200, 257, 1024])
def test_leggauss_bit_exact(N):
    from jaxfun_b200.utils.fastgl import leggauss
    a = leggauss(N)
    b = O.leggauss(N)
    assert a.shape == (2, N)
    assert np.array_equal(a, b)  # bit-exact
    x, w = np.polynomial.legendre.leggauss(N)
    assert np.abs(a[0] - x).max() < 1e-15 * 4 and np.abs(a[1] - w).max() < 4e-14


@pytest.mark.parametrize("name,OC,PC,kw", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("n", [8, 13, 64])
def test_quadrature_tables_bit_exact(name, OC, PC, kw, n):
    N = 8
    o, p = OC(N, **kw), PC(N, **kw)
    xo, wo = o.quad_points_and_weights(n)
    xp, wp = p.quad_points_and_weights(n)
    assert np.array_equal(xo, xp) and np.array_equal(wo, wp)
    for dom in ((-2.0, 3.0), (0.0, 1.0)):
        o, p = OC(N, domain=dom, **kw), PC(N, domain=dom, **kw)
        assert np.array_equal(o.mesh("quadrature", n), p.mesh("quadrature", n))
        assert float(o.domain_factor) == float(p.domain_factor)


@pytest.mark.parametrize("N", [2, 8, 12, 64])
def test_fourier_wavenumbers_bit_exact(N):
    from jaxfun_b200.galerkin.Fourier import fourier_wavenumbers
    for elim in (False, True):
        a = fourier_wavenumbers(N, elim)
        assert np.array_equal(a, O.fourier_wavenumbers(N, elim))
        ref = np.fft.fftfreq(N, 1.0 / N).astype(int)
        if elim:
            ref[N // 2] = 0
        assert np.array_equal(a, ref)


@pytest.mark.parametrize("name,OC,PC,kw", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("dom", [None, (-2.0, 3.0)])
def test_dense_tables_match_oracle_transforms(name, OC, PC, kw, dom):
    rng = np.random.default_rng(7)
    N, n = 10, 14
    o, p = OC(N, domain=dom, **kw), PC(N, domain=dom, **kw)
    c = rng.standard_normal((N, 3)) + (1j * rng.standard_normal((N, 3)) if name == "Fourier" else 0)
    assert np.allclose(o.norm_squared(), p.norm_squared(), rtol=1e-13, atol=0)
    for nn in (N, n):
        ub = o.backward(c, N=nn, axis=0)
        assert np.abs(p._dense_table(L.OP_BACKWARD, N, nn, 0) @ c - ub).max() <= 1e-13 * np.abs(ub).max()
        uf = o.forward(ub, axis=0)
        assert np.abs(p._dense_table(L.OP_FORWARD, N, nn, 0) @ ub - uf).max() <= 1e-13 * np.abs(uf).max()
        us = o.scalar_product(ub, axis=0)
        assert np.abs(p._dense_table(L.OP_SCALAR_PRODUCT, N, nn, 0) @ ub - us).max() <= 1e-13 * np.abs(us).max()
        if name == "ChebyshevU":
            continue
        for k in (1, 2, 3):
            up = o.backward_primitive(c, k=k, N=nn, axis=0)
            T = p._dense_table(L.OP_BACKWARD_PRIMITIVE, N, nn, k)
            assert np.abs(T @ c - up).max() <= 1e-12 * np.abs(up).max()


def test_device_scope_and_key_are_neutral_for_host_arrays():
    """Plans are per device: caches key on the array's device and creation runs under `device_scope` (a no-op for host
    arrays and for the device that is current already)."""
    import numpy as np
    from jaxfun_b200.engine import device_key, device_scope
    x = np.zeros(3)
    assert device_key(x) == "host"
    with device_scope(x) as sc:
        assert sc._ctx is None
