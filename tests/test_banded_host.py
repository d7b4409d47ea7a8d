"""Wavenumber-batched banded solver (la/tpmatrix.py:590-1014, la/diamatrix.py:1937-1973 of the reference), host side:
the oracle against vectors made by the reference's own functions (tests/golden/make_golden_banded.py) and against dense
solves; the assembly logic of `tpmats_wavenumber_factor`; argument checks of the C ABI on a CPU-only host."""
import ctypes as C
import os

import numpy as np
import pytest
import sympy as sp

import jaxfun_oracle as O

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_banded.npz"))
CASES = sorted({k.split("/")[0] for k in GOLD.files})


def dense_from_dia(offsets, data):
    n = data.shape[-1]
    A = np.zeros((n, n), dtype=data.dtype)
    for d, off in enumerate(offsets):
        for j in range(n):
            i = j - off
            if 0 <= i < n:
                A[i, j] = data[d, j]
    return A


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_functions(name):
    W, P, rhs = GOLD[name + "/W"], GOLD[name + "/P"], GOLD[name + "/rhs"]
    offsets = tuple(int(o) for o in GOLD[name + "/offsets"])
    band, p, q = O.banded_lu_no_pivot(O.wavenumber_band_data(W, P), offsets)
    ref = GOLD[name + "/band_lu"]
    assert band.shape == ref.shape
    assert np.abs(band - ref).max() <= 1e-14 * np.abs(ref).max()
    x = O.banded_solve(band, p, q, rhs)
    assert np.abs(x - GOLD[name + "/x"]).max() <= 1e-13 * np.abs(GOLD[name + "/x"]).max()


@pytest.mark.parametrize("name", CASES)
def test_golden_solution_solves_the_dense_systems(name):
    """Known-answer check of the golden data themselves: B_s x_s = b_s with B_s rebuilt densely."""
    W, P, rhs, x = (GOLD[name + "/" + k] for k in ("W", "P", "rhs", "x"))
    offsets = tuple(int(o) for o in GOLD[name + "/offsets"])
    data = O.wavenumber_band_data(W, P)
    for s in range(data.shape[0]):
        A = dense_from_dia(offsets, data[s])
        assert np.abs(A @ x[s] - rhs[s]).max() < 1e-12 * np.abs(rhs[s]).max()
        assert np.abs(np.linalg.solve(A, rhs[s]) - x[s]).max() < 1e-12 * np.abs(x[s]).max()


def test_oracle_solver_layouts():
    """`WavenumberSolver.solve` follows the transposition of la/tpmatrix.py:931-977 for every position of the polynomial axis."""
    rng = np.random.default_rng(0)
    offsets = (-2, 0, 2)
    for shape, pa in [((6, 10), 1), ((10, 6), 0), ((4, 9, 5), 1), ((3, 4, 8), 2)]:
        n = shape[pa]
        n_sys = int(np.prod(shape)) // n
        P = rng.standard_normal((2, 3, n))
        P[0, 1] += 10
        W = np.abs(rng.standard_normal((2, n_sys))) + 1
        S = O.WavenumberSolver(pa, shape, W, P, offsets)
        rhs = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
        x = S.solve(rhs)
        data = O.wavenumber_band_data(W, P)
        xm = np.moveaxis(x, pa, -1).reshape(n_sys, n)
        rm = np.moveaxis(rhs, pa, -1).reshape(n_sys, n)
        for s in range(n_sys):
            assert np.allclose(dense_from_dia(offsets, data[s]) @ xm[s], rm[s], rtol=0, atol=1e-12)


def test_oracle_rejects_zero_pivot():
    P = np.ones((1, 3, 6))
    P[0, 1, 2] = 0.0
    P[0, 0] = P[0, 2] = 0.0
    with pytest.raises(ValueError):
        O.WavenumberSolver(1, (2, 6), np.ones((1, 2)), P, (-1, 0, 1))


def test_wavenumber_factor_assembly_matches_kron():
    """`tpmats_wavenumber_factor` of the product (host part): the systems it hands to the device are the diagonal blocks of the
    Kronecker sum, for Fourier x Legendre and Fourier x Fourier x Legendre Poisson operators (tests/la/test_tpmatrices_solvers.py:
    68-77, 305-320 of the reference build the same operators through `inner`)."""
    import jaxfun_b200 as jf
    from jaxfun_b200.galerkin import tpsolve as S
    n = sp.Symbol("n", integer=True)
    D = jf.FunctionSpace(10, jf.Legendre, {"left": {"D": 0}, "right": {"D": 0}}, scaling=n + 1)
    for spaces in ([jf.Fourier(6), D], [D, jf.Fourier(6)], [jf.Fourier(4), jf.Fourier(6), D]):
        T = jf.TensorProduct(*spaces)
        terms = S.laplace_terms(T)
        w = S.tpmats_wavenumber_factor(terms)
        pa = [i for i, s in enumerate(T.basespaces) if not s.complex_data][0]
        assert w.poly_axis == pa and w.offsets == (-2, 0, 2) and not w.band_complex
        K = 0
        for sc, mats in terms:
            k = np.array([[sc]])
            for m in mats:
                k = np.kron(k, np.diag(m) if np.ndim(m) == 1 else m)
            K = K + k
        shape = w.shape
        data = O.wavenumber_band_data(w.weights, w.diags)
        idx = np.arange(int(np.prod(shape))).reshape(shape)
        lines = np.moveaxis(idx, pa, -1).reshape(-1, shape[pa])
        for s, line in enumerate(lines):
            assert np.allclose(K[np.ix_(line, line)], dense_from_dia(w.offsets, data[s]), rtol=0, atol=1e-10 * np.abs(K).max())
        off_block = K.copy()
        for line in lines:
            off_block[np.ix_(line, line)] = 0
        assert np.abs(off_block).max() == 0
    with pytest.raises(TypeError):
        S.tpmats_wavenumber_factor("not a valid input")           # test_tpmatrices_solvers.py:158-160
    with pytest.raises(S.SolverNotApplicable):
        S.tpmats_wavenumber_factor(S.laplace_terms(jf.TensorProduct(D, D)))
    assert isinstance(S.poisson_solver(jf.TensorProduct(D, D)), S.KroneckerSumSolver)
    assert isinstance(S.poisson_solver(jf.TensorProduct(jf.Fourier(6), D)), S.WavenumberBandedSolver)


def test_tpmatrices_dispatch_and_caching():
    """`TPMatrices.lu_factor` (la/tpmatrix.py:386-427; tests/la/test_tpmatrices_solvers.py:245-270): wavenumber solver for
    Fourier x polynomial operators, diagonalisation for all-polynomial Kronecker sums, cached; other structures are refused."""
    import jaxfun_b200 as jf
    from jaxfun_b200.galerkin import tpsolve as S
    n = sp.Symbol("n", integer=True)
    bcs = {"left": {"D": 0}, "right": {"D": 0}}
    D = jf.FunctionSpace(10, jf.Legendre, bcs, scaling=n + 1)
    Cb = jf.FunctionSpace(12, jf.Chebyshev, bcs, scaling=n + 1)
    A = S.TPMatrices(S.laplace_terms(jf.TensorProduct(jf.Fourier(8), D)))
    assert isinstance(A.lu_factor(), S.WavenumberBandedSolver) and A.lu_factor() is A.lu_factor()
    B = S.TPMatrices(S.laplace_terms(jf.TensorProduct(D, Cb)))
    lu = B.lu_factor()
    assert isinstance(lu, S.KroneckerSumSolver) and lu is B.lu_factor()
    assert isinstance(S.TPMatrices(S.laplace_terms(jf.TensorProduct(D, Cb, D))).lu_factor(), S.KroneckerSumSolver)
    # the factors solve the Kronecker system (host arithmetic with the solver's own tables)
    K = 0
    for sc, mats in S.laplace_terms(jf.TensorProduct(D, Cb)):
        K = K + sc * np.kron(mats[0], mats[1])
    f = np.random.default_rng(0).standard_normal((8, 10))
    u = np.einsum("ia,jb,ab->ij", lu.V[0], lu.V[1], np.einsum("ia,jb,ab->ij", lu.W[0], lu.W[1], f) * lu.Dinv)
    assert np.abs(K @ u.ravel() - f.ravel()).max() < 1e-11 * np.abs(f).max()
    # the order of the terms does not matter; a sum that is not a Kronecker sum is refused
    assert isinstance(S.tpmats_lu_factor(list(reversed(S.laplace_terms(jf.TensorProduct(D, Cb))))), S.KroneckerSumSolver)
    with pytest.raises(S.SolverNotApplicable):
        S.TPMatrices(S.laplace_terms(jf.TensorProduct(D, Cb), alpha=2.0)).lu_factor()           # three terms on two axes
    with pytest.raises(S.SolverNotApplicable):
        e3, e4, e2 = np.eye(3), np.eye(4), np.eye(2)                           # three different matrices on axis 0
        S.tpmats_lu_factor([(1.0, [e3 + 1, e4, e2]), (1.0, [e3 * 2 + 1, e4 + 3, e2]), (1.0, [e3 * 3 + 1, e4, e2 + 1])])


def test_sharded_solver_blocks():
    """`shard(rank, size)`: the blocks of the reference's multi-device mode (la/tpmatrix.py:786-812) — contiguous wavenumber
    ranges of axis 0; together they reproduce the global solve (checked with the oracle); poly_axis = 0 is refused."""
    from jaxfun_b200.galerkin.tpsolve import WavenumberBandedSolver
    rng = np.random.default_rng(4)
    shape, pa, offsets = (8, 6, 11), 2, (-2, 0, 2)
    P = rng.standard_normal((2, 3, 11))
    P[0, 1] += 12
    W = np.abs(rng.standard_normal((2, 48))) + 1
    S = WavenumberBandedSolver(pa, shape, W, P, offsets)
    rhs = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    full = O.WavenumberSolver(pa, shape, W, P, offsets).solve(rhs)
    for size in (2, 4):
        parts = []
        for r in range(size):
            Sr = S.shard(r, size)
            assert Sr.shape == (8 // size, 6, 11) and Sr.n_sys == 48 // size
            blk = slice(r * 8 // size, (r + 1) * 8 // size)
            parts.append(O.WavenumberSolver(pa, Sr.shape, Sr.weights, Sr.diags, Sr.offsets).solve(rhs[blk]))
        assert np.allclose(np.concatenate(parts, axis=0), full, rtol=0, atol=1e-14)
    with pytest.raises(ValueError, match="axis 0"):
        WavenumberBandedSolver(0, (11, 6), W[:, :6], P, offsets).shard(0, 2)
    with pytest.raises(ValueError):
        S.shard(0, 3)


def test_banded_cabi_argument_checks_and_no_cpu_fallback():
    import torch
    from jaxfun_b200 import _lib
    lib = _lib.load()
    offs = (C.c_int32 * 3)(-1, 0, 1)
    W = np.ones((1, 4))
    P = np.ones((1, 3, 8))
    P[0, 1] = 4

    def desc(**kw):
        d = _lib.BandedDesc()
        d.abi_version, d.dtype, d.band_complex, d.n_terms = _lib.JFX_ABI_VERSION, _lib.C128, 0, 1
        d.n, d.n_sys, d.n_diags = 8, 4, 3
        d.offsets, d.weights, d.diags = offs, W.ctypes.data, P.ctypes.data
        for k, v in kw.items():
            setattr(d, k, v)
        return d

    h = C.c_void_p()
    assert lib.jfx_banded_create(None, C.byref(h)) == -1
    for bad in (dict(abi_version=99), dict(dtype=9), dict(n=0), dict(n_sys=0), dict(n_terms=0), dict(n_terms=9),
                dict(n_diags=0), dict(weights=None), dict(diags=None), dict(band_complex=1, dtype=_lib.F64)):
        d = desc(**bad)
        assert lib.jfx_banded_create(C.byref(d), C.byref(h)) == -1 and not h.value, bad
    for bad_offs in ((0, -1, 1), (-1, 1, 2), (-9, 0, 1)):          # not increasing / no main diagonal / outside the matrix
        d = desc(offsets=(C.c_int32 * 3)(*bad_offs))
        assert lib.jfx_banded_create(C.byref(d), C.byref(h)) == -1 and not h.value, bad_offs
    assert lib.jfx_banded_solve(None, None, None, None, 1, 1) == -1
    assert lib.jfx_banded_info(None, None, None, None) == -1
    lib.jfx_banded_destroy(None)
    if not torch.cuda.is_available():
        d = desc()
        assert lib.jfx_banded_create(C.byref(d), C.byref(h)) == -3 and not h.value
        from jaxfun_b200.galerkin.tpsolve import WavenumberBandedSolver
        S = WavenumberBandedSolver(1, (4, 8), W, P, (-1, 0, 1))
        with pytest.raises(_lib.JfxError) as e:
            S.solve(torch.zeros(4, 8, dtype=torch.complex128))
        assert e.value.code == -3


def test_axis_matrices_known_answers():
    """The per-axis operators `laplace_terms` hands to the solvers, against closed forms: Shen's Legendre-Dirichlet basis
    phi_k = P_k - P_{k+2} has (phi_j, phi_k'') = -(4k + 6) delta_jk and the pentadiagonal mass 2/(2k+1) + 2/(2k+5), -2/(2k+5);
    a Fourier axis on (0, 2 pi) has mass 2 pi and second-derivative symbol -k^2 2 pi; on (0, 1) the factors df = 2 pi enter as
    (1 / df) and df^2 / df."""
    import jaxfun_b200 as jf
    from jaxfun_b200.galerkin import tpsolve as S
    N = 14
    D = jf.FunctionSpace(N, jf.Legendre, {"left": {"D": 0}, "right": {"D": 0}})
    A, M = S.stiffness_matrix(D, 2), S.mass_matrix(D)
    k = np.arange(N - 2)
    assert np.abs(A - np.diag(-(4.0 * k + 6))).max() < 1e-12
    Mx = np.diag(2 / (2 * k + 1) + 2 / (2 * k + 5)) + np.diag(-2 / (2 * k[:-2] + 5), 2) + np.diag(-2 / (2 * k[:-2] + 5), -2)
    assert np.abs(M - Mx).max() < 1e-14
    offs, data = S.dia_from_dense(M)
    assert offs == (-2, 0, 2) and np.allclose(data[1], np.diagonal(M)) and np.allclose(data[2][2:], np.diagonal(M, 2))
    assert np.allclose(data[0][:-2], np.diagonal(M, -2)) and data[0][-2:].tolist() == [0.0, 0.0] and data[2][:2].tolist() == [0.0, 0.0]
    F = jf.Fourier(8)
    a, m = S._axis_matrices(F)
    kk = np.asarray(F.wavenumbers(), dtype=float)
    assert np.allclose(m, 2 * np.pi) and np.allclose(a, -(kk**2) * 2 * np.pi)
    F1 = jf.Fourier(8, domain=(0.0, 1.0))
    a1, m1 = S._axis_matrices(F1)
    assert np.allclose(m1, 1.0) and np.allclose(a1, -((2 * np.pi * kk) ** 2))


def test_plain_c_client_builds_and_refuses_without_a_device(tmp_path):
    """tools/banded_client.c is a C99 program against include/jfx.h alone (no Python, no torch): it compiles with gcc, links
    libjfx.so, and on a host without a GPU stops with the library's JFX_ERR_CUDA (exit code 3) instead of computing anything."""
    import subprocess
    import torch
    cuda_home = "/usr/local/cuda"
    if not os.path.exists(os.path.join(cuda_home, "include", "cuda_runtime_api.h")):
        pytest.skip("CUDA toolkit headers not found")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "banded_client")
    subprocess.check_call(["gcc", "-O2", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), "-I", os.path.join(cuda_home, "include"),
                           os.path.join(root, "tools", "banded_client.c"), "-L", os.path.join(root, "jaxfun_b200"), "-ljfx",
                           "-L", os.path.join(cuda_home, "lib64"), "-lcudart", "-lm",
                           "-Wl,-rpath," + os.path.join(root, "jaxfun_b200"), "-o", exe])
    if torch.cuda.is_available():
        pytest.skip("GPU present: the device run of the client is not part of the CPU suite")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 3 and "no CUDA device" in r.stdout
