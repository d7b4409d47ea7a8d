"""Wavenumber-batched banded solver on the GPU through the C ABI (`jfx_banded_*`): the cases of test_banded_emul.py against
the oracle, the reference-made vectors, and the reference's acceptance tests for Fourier x polynomial Poisson problems
(tests/la/test_tpmatrices_solvers.py:68-77, 125-146, 305-330: agreement with the Kronecker solve, L2 error of the
manufactured solution)."""
import os

import numpy as np
import pytest
import sympy as sp
import torch

import jaxfun_oracle as O
import jaxfun_b200 as jf
from banded_cases import CASES, make_case, tolerance
from jaxfun_b200 import _lib as L
from jaxfun_b200.galerkin import tpsolve as S

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n_ = sp.Symbol("n", integer=True)
BCS = {"left": {"D": 0}, "right": {"D": 0}}


def dev(x, cuda):
    return torch.from_numpy(np.ascontiguousarray(x)).to(cuda)


@pytest.mark.parametrize("name", sorted(CASES))
def test_banded_matches_oracle(cuda, name):
    shape, pa, offsets, W, P, rhs = make_case(name)
    So = O.WavenumberSolver(pa, shape, W, P, offsets)
    want = So.solve(rhs.astype(np.complex128 if np.iscomplexobj(rhs) else np.float64))
    tol = tolerance(str(rhs.dtype))
    Sp = S.WavenumberBandedSolver(pa, shape, W, P, offsets)
    r = dev(rhs, cuda)
    x = Sp.solve(r)
    assert x.shape == r.shape and x.dtype == r.dtype
    assert np.array_equal(r.cpu().numpy(), rhs)                                   # the right-hand side is not written
    assert np.abs(x.cpu().numpy() - want).max() <= tol * np.abs(want).max()
    dt = jf.engine.jfx_dtype(r.dtype)
    assert Sp.bandwidths(dt)[:2] == (So.p, So.q)
    lu = Sp.factors(dt)
    assert np.abs(lu - So.band_lu).max() <= tol * np.abs(So.band_lu).max()
    x2 = Sp.solve(r, out=r)                                                       # in place
    assert x2 is r and torch.equal(x2, x)


def test_banded_matches_reference_vectors(cuda):
    gold = np.load(os.path.join(ROOT, "tests", "golden", "reference_banded.npz"))
    for name in sorted({k.split("/")[0] for k in gold.files}):
        W, P, rhs = gold[name + "/W"], gold[name + "/P"], gold[name + "/rhs"]
        offsets = tuple(int(o) for o in gold[name + "/offsets"])
        Sp = S.WavenumberBandedSolver(1, rhs.shape, W, P, offsets)
        x = Sp.solve(dev(rhs, cuda)).cpu().numpy()
        assert np.abs(x - gold[name + "/x"]).max() <= 1e-12 * np.abs(gold[name + "/x"]).max(), name
        lu = Sp.factors(L.C128 if np.iscomplexobj(rhs) else L.F64)
        assert np.abs(lu - gold[name + "/band_lu"]).max() <= 1e-13 * np.abs(gold[name + "/band_lu"]).max(), name


def test_banded_zero_pivot_is_refused(cuda):
    P = np.ones((1, 3, 6))
    P[0, 1] = 4.0
    W = np.ones((1, 3))
    W[0, 1] = 0.0
    Sp = S.WavenumberBandedSolver(1, (3, 6), W, P, (-1, 0, 1))
    with pytest.raises(ValueError, match="pivot"):                                # la/diamatrix.py:461-471 raises ValueError
        Sp.solve(torch.ones(3, 6, dtype=torch.float64, device=cuda))
    with pytest.raises(ValueError):
        S.WavenumberBandedSolver(1, (3, 6), np.ones((1, 3)), P, (-1, 0, 1)).solve(torch.ones(6, 3, dtype=torch.float64, device=cuda))
    with pytest.raises(TypeError):
        S.WavenumberBandedSolver(1, (3, 6), np.ones((1, 3)) * 1j, P, (-1, 0, 1)).solve(
            torch.ones(3, 6, dtype=torch.float64, device=cuda))


def test_banded_sharded_blocks_reproduce_the_global_solve(cuda):
    """la/tpmatrix.py:985-1013: every rank solves its own block of axis 0 with its own factors, no communication."""
    shape, pa, offsets, W, P, rhs = make_case("3d_last")
    S0 = S.WavenumberBandedSolver(pa, shape, W, P, offsets)
    full = S0.solve(dev(rhs, cuda))
    size = 3
    parts = [S0.shard(r, size).solve(dev(rhs[r * 2:(r + 1) * 2], cuda)) for r in range(size)]
    assert torch.equal(torch.cat(parts, dim=0), full)


def test_banded_in_cuda_graph(cuda):
    """jfx_banded_solve only enqueues one launch: it can be captured and replayed."""
    shape, pa, offsets, W, P, rhs = make_case("2d_last_penta")
    Sp = S.WavenumberBandedSolver(pa, shape, W, P, offsets)
    r = dev(rhs, cuda)
    want = Sp.solve(r).clone()
    out = torch.empty_like(r)
    side = torch.cuda.Stream(device=cuda)
    side.wait_stream(torch.cuda.current_stream(cuda))
    with torch.cuda.stream(side):
        Sp.solve(r, out=out)
    torch.cuda.current_stream(cuda).wait_stream(side)
    torch.cuda.synchronize(cuda)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        Sp.solve(r, out=out)
    out.zero_()
    g.replay()
    torch.cuda.synchronize(cuda)
    assert torch.equal(out, want)


def _poisson_fourier_poly(cuda, spaces, ue, syms):
    T = jf.TensorProduct(*spaces)
    lap = sp.lambdify(syms, sum(sp.diff(ue, s, 2) for s in syms), "numpy")
    xq = T.mesh()
    f = np.asarray(lap(*xq) + sum(0 * v for v in xq), dtype=complex)
    b = T.scalar_product(dev(f, cuda))                                            # (v, div grad ue)_w
    solver = S.poisson_solver(T)
    assert isinstance(solver, S.WavenumberBandedSolver)
    uh = solver.solve(b)
    return T, b, uh, solver


@pytest.mark.parametrize("base", ["Legendre", "Chebyshev"])
@pytest.mark.parametrize("order", ["FD", "DF"])
def test_poisson_fourier_poly_2d(cuda, base, order):
    """tests/la/test_tpmatrices_solvers.py:125-146: ue = cos(2x)(1 - y^2), N = 16; the wavenumber solve agrees with the dense
    Kronecker solve and the manufactured solution is reproduced."""
    N = 16
    F = jf.Fourier(N)
    D = jf.FunctionSpace(N, getattr(jf, base), BCS, scaling=n_ + 1)
    x, y = sp.symbols("x y", real=True)
    if order == "FD":
        spaces, ue, syms = [F, D], sp.cos(2 * x) * (1 - y**2), (x, y)
    else:
        spaces, ue, syms = [D, F], (1 - x**2) * sp.cos(2 * y), (x, y)
    T, b, uh, solver = _poisson_fourier_poly(cuda, spaces, ue, syms)
    K = 0
    for sc, mats in S.laplace_terms(T):
        k = np.array([[sc]])
        for m in mats:
            k = np.kron(k, np.diag(m) if np.ndim(m) == 1 else m)
        K = K + k
    ref = np.linalg.solve(K, b.cpu().numpy().ravel()).reshape(b.shape)
    assert np.abs(uh.cpu().numpy() - ref).max() < 1e-11 * np.abs(ref).max()
    M = 40
    uj = T.backward(uh, N=(M, M)).cpu().numpy()
    xj = T.mesh(N=(M, M))
    uej = sp.lambdify(syms, ue, "numpy")(*xj)
    l2 = np.linalg.norm(uj - uej) / M
    assert l2 < np.sqrt(10 * np.finfo(float).eps), l2


def test_poisson_fourier_fourier_legendre_3d(cuda):
    """tests/la/test_tpmatrices_solvers.py:305-330."""
    N = 8
    D = jf.FunctionSpace(N, jf.Legendre, BCS, scaling=n_ + 1)
    x, y, z = sp.symbols("x y z", real=True)
    ue = sp.cos(2 * x) * sp.cos(2 * y) * (1 - z**2)
    T, b, uh, _ = _poisson_fourier_poly(cuda, [jf.Fourier(N), jf.Fourier(N), D], ue, (x, y, z))
    M = 20
    uj = T.backward(uh, N=(M, M, M)).cpu().numpy()
    uej = sp.lambdify((x, y, z), ue, "numpy")(*T.mesh(N=(M, M, M)))
    assert np.linalg.norm(uj - uej) / M**1.5 < np.sqrt(10 * np.finfo(float).eps)


def test_tpmatrices_solve_dispatch(cuda):
    """`TPMatrices.solve` (la/tpmatrix.py:429-545) through both factored solvers, against the dense Kronecker solve; Helmholtz
    (alpha != 0) keeps the Fourier x polynomial structure and goes to the wavenumber solver as well."""
    D = jf.FunctionSpace(12, jf.Legendre, BCS, scaling=n_ + 1)
    Cb = jf.FunctionSpace(10, jf.Chebyshev, BCS, scaling=n_ + 1)
    rng = np.random.default_rng(8)
    for spaces, alpha in (([D, Cb], 0.0), ([jf.Fourier(8), D], 0.0), ([jf.Fourier(8), D], -3.0), ([D, jf.Fourier(6), jf.Fourier(4)], -1.5)):
        T = jf.TensorProduct(*spaces)
        terms = S.laplace_terms(T, alpha)
        A = S.TPMatrices(terms)
        shape = tuple(s.dim for s in T.basespaces)
        cplx = any(s.complex_data for s in T.basespaces)
        b = rng.standard_normal(shape) + (1j * rng.standard_normal(shape) if cplx else 0)
        K = 0
        for sc, mats in terms:
            k = np.array([[sc]])
            for m in mats:
                k = np.kron(k, np.diag(m) if np.ndim(m) == 1 else m)
            K = K + k
        ref = np.linalg.solve(K, b.ravel()).reshape(shape)
        got = A.solve(dev(b, cuda)).cpu().numpy()
        assert np.abs(got - ref).max() < 1e-10 * np.abs(ref).max(), (spaces, alpha)
        assert isinstance(A.lu_factor(), S.WavenumberBandedSolver if cplx else S.KroneckerSumSolver)


def test_banded_at_size(cuda):
    """Fourier 1024 x Legendre-Dirichlet 1022 Helmholtz-type systems (offsets -2, 0, 2): residual of the solve and the time of
    one launch, printed for the record (-s)."""
    nF, n = 1024, 1022
    rng = np.random.default_rng(3)
    k = np.fft.fftfreq(nF, 1.0 / nF)
    main = 4.0 + np.arange(n) * 0.01
    P = np.zeros((2, 3, n))
    P[0, 1] = main
    P[1, 0, :n - 2] = -0.2
    P[1, 1] = 1.0
    P[1, 2, 2:] = -0.2
    W = np.stack([np.ones(nF), 1.0 + k**2 / nF])
    Sp = S.WavenumberBandedSolver(1, (nF, n), W, P, (-2, 0, 2))
    rhs = rng.standard_normal((nF, n)) + 1j * rng.standard_normal((nF, n))
    r = dev(rhs, cuda)
    x = Sp.solve(r)
    xs = x.cpu().numpy()
    B = np.einsum("tf,tdp->fdp", W, P)
    res = B[:, 1, :] * xs
    res[:, 2:] += B[:, 0, :n - 2] * xs[:, :-2]          # sub-diagonal -2: entry (j + 2, j) stored at column j
    res[:, :-2] += B[:, 2, 2:] * xs[:, 2:]              # super-diagonal +2: entry (j - 2, j) stored at column j
    assert np.abs(res - rhs).max() < 1e-12 * np.abs(rhs).max()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = torch.empty_like(r)
    for _ in range(3):
        Sp.solve(r, out=out)
    ev0.record()
    for _ in range(10):
        Sp.solve(r, out=out)
    ev1.record()
    torch.cuda.synchronize(cuda)
    ms = ev0.elapsed_time(ev1) / 10
    nbytes = 2 * rhs.nbytes + 5 * 8 * n * nF
    print(f"banded solve {nF} x {n} c128, p = q = 2: {ms * 1e3:.1f} us, {nbytes / ms / 1e6:.0f} GB/s of compulsory traffic")
