"""The device code of the banded solver, executed on the host: jaxfun_b200/csrc/banded.cuh holds the per-system bodies of
the CUDA kernels as __host__ __device__ functions; tools/banded_emul.cpp runs them with a loop in place of the thread grid.
Compared with the oracle (itself pinned to the reference's functions, test_banded_host.py) for every window variant, layout,
dtype and for the golden vectors.  This is how the kernels are checked where there is no GPU; test_banded_gpu.py repeats the
same cases on the device through the C ABI."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import jaxfun_oracle as O
from banded_cases import CASES, make_case, tolerance

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DT = {"float32": 0, "float64": 1, "complex64": 2, "complex128": 3}


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("banded") / "libbanded_emul.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas",
                           os.path.join(ROOT, "tools", "banded_emul.cpp"), "-o", so])
    lib = C.CDLL(so)
    lib.banded_emul.restype = C.c_int
    lib.banded_emul.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int, C.POINTER(C.c_int32), C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
    return lib


def run_emul(lib, shape, pa, offsets, W, P, rhs, inplace=False, rows=0):
    n = shape[pa]
    n_sys = int(np.prod(shape)) // n
    inner = int(np.prod(shape[pa + 1:], dtype=np.int64))
    cband = np.iscomplexobj(W) or np.iscomplexobj(P)
    Wc = np.ascontiguousarray(W, dtype=np.complex128 if cband else np.float64)
    Pc = np.ascontiguousarray(P, dtype=np.complex128 if cband else np.float64)
    p = max((-o for o in offsets if o < 0), default=0)
    q = max((o for o in offsets if o > 0), default=0)
    real = np.float64 if rhs.dtype in (np.float64, np.complex128) else np.float32
    et = (np.complex128 if real is np.float64 else np.complex64) if cband else real
    lu = np.zeros((p + q + 1, n, n_sys), dtype=et)
    rhs = np.ascontiguousarray(rhs)
    out = rhs if inplace else np.empty_like(rhs)
    offs = (C.c_int32 * len(offsets))(*offsets)
    rc = lib.banded_emul(DT[str(rhs.dtype)], int(cband), W.shape[0], n, n_sys, len(offsets), offs, Wc.ctypes.data, Pc.ctypes.data,
                         rhs.ctypes.data, out.ctypes.data, inner, lu.ctypes.data, int(rows))
    return rc, out, np.transpose(lu, (2, 0, 1))


@pytest.mark.parametrize("rows", [0, 1], ids=["thread-per-system", "row-tiles"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_device_code_on_host_matches_oracle(emul, name, rows):
    shape, pa, offsets, W, P, rhs = make_case(name)
    if rows == 1 and (pa != len(shape) - 1 or max(abs(o) for o in offsets) > 4):
        pytest.skip("the row-tile variant serves a last polynomial axis with bandwidth <= 4")
    S = O.WavenumberSolver(pa, shape, W, P, offsets)
    want = S.solve(rhs.astype(np.complex128 if np.iscomplexobj(rhs) else np.float64))
    tol = tolerance(str(rhs.dtype))
    rc, x, lu = run_emul(emul, shape, pa, offsets, W, P, rhs, rows=rows)
    assert rc == 0
    assert np.abs(lu - S.band_lu).max() <= tol * np.abs(S.band_lu).max()
    assert np.abs(x - want).max() <= tol * np.abs(want).max()
    rc, x2, _ = run_emul(emul, shape, pa, offsets, W, P, rhs.copy(), inplace=True, rows=rows)        # rhs == out
    assert rc == 0 and np.array_equal(x2, x)


def test_device_code_on_host_matches_reference_vectors(emul):
    gold = np.load(os.path.join(ROOT, "tests", "golden", "reference_banded.npz"))
    for name in sorted({k.split("/")[0] for k in gold.files}):
        W, P, rhs = gold[name + "/W"], gold[name + "/P"], gold[name + "/rhs"]
        offsets = tuple(int(o) for o in gold[name + "/offsets"])
        rc, x, lu = run_emul(emul, rhs.shape, 1, offsets, W, P, rhs)
        assert rc == 0
        if max(abs(o) for o in offsets) <= 4:
            assert np.array_equal(run_emul(emul, rhs.shape, 1, offsets, W, P, rhs, rows=1)[1], x), name
        assert np.abs(lu - gold[name + "/band_lu"]).max() <= 1e-13 * np.abs(gold[name + "/band_lu"]).max(), name
        assert np.abs(x - gold[name + "/x"]).max() <= 1e-12 * np.abs(gold[name + "/x"]).max(), name


def test_device_code_flags_zero_and_nonfinite_pivots(emul):
    P = np.ones((1, 3, 6))
    P[0, 1] = 4.0
    W = np.ones((1, 3))
    rhs = np.ones((3, 6))
    assert run_emul(emul, (3, 6), 1, (-1, 0, 1), W, P, rhs)[0] == 0
    P0 = P.copy()
    P0[0, 1, 0] = 0.0
    assert run_emul(emul, (3, 6), 1, (-1, 0, 1), W, P0, rhs)[0] == 1
    Pn = P.copy()
    Pn[0, 1, 3] = np.inf
    assert run_emul(emul, (3, 6), 1, (-1, 0, 1), W, Pn, rhs)[0] == 1
    W0 = W.copy()
    W0[0, 1] = 0.0                                   # one singular system among healthy ones
    assert run_emul(emul, (3, 6), 1, (-1, 0, 1), W0, P, rhs)[0] == 1


def test_device_code_on_host_fuzz(emul):
    """Seeded random systems: shapes, position of the polynomial axis, one- and two-sided bands up to the generic variant,
    real / complex bands and all four dtypes, both kernel variants where they apply."""
    rng = np.random.default_rng(20261018)
    for trial in range(60):
        nd = int(rng.integers(2, 4))
        shape = tuple(int(v) for v in rng.integers(1, 9, size=nd))
        pa = int(rng.integers(0, nd))
        n = int(rng.integers(1, 40))
        shape = shape[:pa] + (n,) + shape[pa + 1:]
        n_sys = int(np.prod(shape)) // n
        wmax = int(rng.choice([1, 2, 4, 8, 12]))
        lo = int(rng.integers(0, min(wmax, n - 1) + 1)) if n > 1 else 0
        hi = int(rng.integers(0, min(wmax, n - 1) + 1)) if n > 1 else 0
        offs = sorted({0} | {int(o) for o in rng.integers(-lo, hi + 1, size=4)} | ({-lo} if lo else set()) | ({hi} if hi else set()))
        cband = bool(rng.integers(0, 2))
        dt = str(rng.choice(["complex128", "complex64"] if cband else ["float64", "float32", "complex128", "complex64"]))
        nt = int(rng.integers(1, 4))
        P = rng.standard_normal((nt, len(offs), n))
        W = rng.standard_normal((nt, n_sys))
        if cband:
            P = P + 1j * rng.standard_normal(P.shape)
            W = W + 1j * rng.standard_normal(W.shape)
        B = np.einsum("tf,tdp->fdp", W, P)
        P[0, offs.index(0), :] += np.abs(B).sum(axis=1).max() + 1.0
        W[0, :] = np.abs(W[0, :]) + 1.0
        rhs = rng.standard_normal(shape)
        if dt.startswith("complex"):
            rhs = rhs + 1j * rng.standard_normal(shape)
        rhs = rhs.astype(dt)
        want = O.WavenumberSolver(pa, shape, W, P, tuple(offs)).solve(rhs.astype(np.complex128 if np.iscomplexobj(rhs) else np.float64))
        tol = tolerance(dt)
        for rows in (0, 1):
            if rows and (pa != nd - 1 or max(abs(o) for o in offs) > 4):
                continue
            rc, x, _ = run_emul(emul, shape, pa, tuple(offs), W, P, rhs, rows=rows)
            assert rc == 0, (trial, shape, pa, offs, dt)
            assert np.abs(x - want).max() <= tol * max(np.abs(want).max(), 1e-300), (trial, shape, pa, offs, dt, rows)
