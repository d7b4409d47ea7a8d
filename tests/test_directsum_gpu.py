"""DirectSum spaces (inhomogeneous boundary values: homogeneous Composite + fixed boundary lift, composite.py:502-638 of
the reference) on the GPU against the oracle's restatement; the lifting basis itself is pinned on the host against the
reference's own get_bc_basis (tests/test_directsum_host.py)."""
import numpy as np
import pytest
import torch

import jaxfun_oracle as O
import jaxfun_b200 as jf

pytestmark = pytest.mark.gpu


def dev(x, cuda):
    return torch.from_numpy(np.ascontiguousarray(x)).to(cuda)


def rel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else a
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("base", ["Chebyshev", "Legendre"])
@pytest.mark.parametrize("N,dom", [(12, None), (32, None), (64, (-2.0, 3.0))])
def test_directsum_1d_matches_oracle(cuda, base, N, dom):
    bcs = {"left": {"D": 1.0}, "right": {"D": -2.0}}
    Dp = jf.FunctionSpace(N, getattr(jf, base), bcs, domain=dom)
    Do = O.DirectSum(O.Composite(N, getattr(O, base), {0: 1, 2: -1}, domain=dom), bcs)
    assert isinstance(Dp, jf.DirectSum) and Dp.dim == N - 2 == Do.dim
    assert np.abs(Dp.c_b - Do.c_b).max() < 1e-14
    rng = np.random.default_rng(N)
    c = rng.standard_normal((5, N - 2))
    u_ref = Do.backward(c, axis=-1)
    assert rel(Dp.backward(dev(c, cuda)), u_ref) < 1e-12
    assert rel(Dp.backward(dev(c, cuda), N=N + 8), Do.backward(c, N=N + 8, axis=-1)) < 1e-12
    assert rel(Dp.backward_primitive(dev(c, cuda), 1), Do.backward_primitive(c, 1, axis=-1)) < 1e-11
    assert rel(Dp.to_orthogonal(dev(c, cuda)), Do.to_orthogonal(c, axis=-1)) < 1e-13
    a = rng.standard_normal((5, N))
    assert rel(Dp.from_orthogonal(dev(a, cuda)), Do.from_orthogonal(a, axis=-1)) < 1e-11
    assert rel(Dp.forward(dev(u_ref, cuda)), Do.forward(u_ref, axis=-1)) < 1e-11
    assert rel(Dp.forward(Dp.backward(dev(c, cuda))), c) < 1e-11
    assert rel(Dp.scalar_product(dev(u_ref, cuda)), Do.scalar_product(u_ref, axis=-1)) < 1e-12
    # the expansion takes the prescribed boundary values whatever the free coefficients are
    lo, hi = (float(v) for v in Dp.domain)
    ends = Dp.evaluate(np.array([lo, hi]), dev(c, cuda)).cpu().numpy()
    assert np.abs(ends - np.array([1.0, -2.0])).max() < 1e-12 * max(1.0, np.abs(u_ref).max())


def test_directsum_along_a_leading_axis(cuda):
    N = 24
    bcs = {"left": {"D": 0.5}, "right": {"D": 0.25}}
    Dp = jf.FunctionSpace(N, jf.Legendre, bcs)
    Do = O.DirectSum(O.Composite(N, O.Legendre, {0: 1, 2: -1}), bcs)
    rng = np.random.default_rng(1)
    c = rng.standard_normal((N - 2, 6))
    u_ref = Do.backward(c, axis=0)
    assert rel(Dp.backward(dev(c, cuda), axis=0), u_ref) < 1e-12
    assert rel(Dp.forward(dev(u_ref, cuda), axis=0), c) < 1e-11


def test_directsum_tensor_product_fourier_legendre(cuda):
    """Fourier x (Legendre Dirichlet with boundary data depending on x) — the space of examples/poisson2D_periodic.py —
    against the oracle's DirectSumTPS; the boundary functions are reproduced for arbitrary free coefficients."""
    import sympy as sp
    x, y = sp.symbols("x y", real=True)
    ue = sp.cos(2 * x) * (1 - y**2) + sp.sin(x) * y + 0.3
    bcs = {"left": {"D": ue.subs(y, -1)}, "right": {"D": ue.subs(y, 1)}}
    NF, NL = 32, 24
    T = jf.TensorProduct(jf.Fourier(NF), jf.FunctionSpace(NL, jf.Legendre, bcs))
    Fo, Co = O.Fourier(NF), O.Composite(NL, O.Legendre, {0: 1, 2: -1})
    xm = np.asarray(Fo.mesh())
    samples = [sp.lambdify(x, ue.subs(y, s))(xm) + 0 * xm for s in (-1, 1)]
    To = O.DirectSumTPS([Fo, Co], 1, {"left": {"D": 1}, "right": {"D": 1}}, samples)
    assert np.abs(T.lift - To.lift).max() < 1e-14
    rng = np.random.default_rng(5)
    c = rng.standard_normal((NF, NL - 2)) + 1j * rng.standard_normal((NF, NL - 2))
    assert rel(T.to_orthogonal(dev(c, cuda)), To.to_orthogonal(c)) < 1e-13
    u_ref = To.backward(c)
    assert rel(T.backward(dev(c, cuda)), u_ref) < 1e-12
    assert rel(T.forward(dev(u_ref, cuda)), To.forward(u_ref)) < 1e-11
    assert rel(T.forward(T.backward(dev(c, cuda))), c) < 1e-11
    a = T.to_orthogonal(dev(c, cuda)).cpu().numpy()
    k = np.arange(NL)
    assert np.abs(Fo.backward(a @ ((-1.0) ** k), axis=0) - samples[0]).max() < 1e-12
    assert np.abs(Fo.backward(a @ np.ones(NL), axis=0) - samples[1]).max() < 1e-12


def test_two_inhomogeneous_directions_on_the_gpu(cuda):
    """Both directions inhomogeneous (tensorproductspace.py:620-668): the transforms are the homogeneous product's engine plans
    plus the host-built transfinite lift; checked against the oracle's ORTHOGONAL tensor product applied to the lifted
    coefficients, and by exact reproduction of a polynomial with Dirichlet / Neumann data on the four sides."""
    import sympy as sp
    x, y = sp.symbols("x y", real=True)
    ue = (x**2 + 2 * x) * (y**3 - y) + 3 + x - 2 * y + sp.Rational(1, 2) * x * y + y**2 * x**3
    domx, domy = (0.0, 2.0), (-1.0, 1.0)
    bcx = {"left": {"D": ue.subs(x, domx[0])}, "right": {"N": ue.diff(x).subs(x, domx[1])}}
    bcy = {"left": {"D": ue.subs(y, domy[0])}, "right": {"D": ue.subs(y, domy[1])}}
    N = 32
    T = jf.TensorProduct(jf.FunctionSpace(N, jf.Legendre, bcx, domain=domx), jf.FunctionSpace(N, jf.Legendre, bcy, domain=domy))
    To = O.TensorProductSpace(O.Legendre(N, domain=domx), O.Legendre(N, domain=domy))
    X, Y = T.mesh()
    u = sp.lambdify((x, y), ue, "numpy")(X, Y)
    c = T.forward(dev(u, cuda))
    assert tuple(c.shape) == (N - 2, N - 2)
    assert rel(T.backward(c), u) < 1e-12
    rng = np.random.default_rng(6)
    cr = rng.standard_normal((N - 2, N - 2))
    a = T.to_orthogonal(dev(cr, cuda)).cpu().numpy()
    assert rel(T.backward(dev(cr, cuda)), To.backward(a)) < 1e-12
    assert rel(T.backward_primitive(dev(cr, cuda), (1, 0)), To.backward_primitive(a, (1, 0))) < 1e-11
    assert rel(T.forward(dev(To.backward(a), cuda)), cr) < 1e-10


def test_two_inhomogeneous_directions_in_3d_on_the_gpu(cuda):
    """(Fourier, DirectSum, DirectSum): the 3-D layout of tensorproductspace.py:704-747.  The transforms are the homogeneous
    3-D product's engine plans plus the host-built lift (tests/test_directsum_tps_host.py checks the lift itself)."""
    import sympy as sp
    x, y, z = sp.symbols("x y z", real=True)
    ue = (sp.cos(x) + 2) * ((y**2 + y) * (z**3 - z) + 3 - 2 * z + sp.Rational(1, 2) * y * z + y)
    domy, domz = (0.0, 2.0), (-1.0, 1.0)
    bcy = {"left": {"D": ue.subs(y, domy[0])}, "right": {"N": ue.diff(y).subs(y, domy[1])}}
    bcz = {"left": {"D": ue.subs(z, domz[0])}, "right": {"D": ue.subs(z, domz[1])}}
    N = 16
    T = jf.TensorProduct(jf.Fourier(8), jf.FunctionSpace(N, jf.Legendre, bcy, domain=domy),
                         jf.FunctionSpace(N, jf.Chebyshev, bcz, domain=domz))
    To = O.TensorProductSpace(O.Fourier(8), O.Legendre(N, domain=domy), O.Chebyshev(N, domain=domz))
    X, Y, Z = T.mesh()
    u = sp.lambdify((x, y, z), ue, "numpy")(X, Y, Z) + 0j
    c = T.forward(dev(u, cuda))
    assert tuple(c.shape) == (8, N - 2, N - 2)
    assert rel(T.backward(c), u) < 1e-12
    rng = np.random.default_rng(7)
    cr = rng.standard_normal((8, N - 2, N - 2)) + 1j * rng.standard_normal((8, N - 2, N - 2))
    a = T.to_orthogonal(dev(cr, cuda)).cpu().numpy()
    assert rel(T.backward(dev(cr, cuda)), To.backward(a)) < 1e-12
    assert rel(T.forward(dev(To.backward(a), cuda)), cr) < 1e-10
