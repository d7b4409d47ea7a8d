"""Parity-folded tensor-core contraction, checked WITHOUT a GPU.

`jaxfun_b200/csrc/dmma_fold.cuh` holds every index formula of the CUDA kernel `dgemm_dmma_fold` (fragment
addresses in the swizzled tiles, plus/minus operand selection, epilogue butterfly and store addresses, the TMA
box copies of a pipeline stage, tensor-map geometry, folded-table construction, symmetry analysis).
`tests/emu/fold_emu.cpp` compiles that header for the host and executes it with a model of TMA tiled copies
(zero fill, 64 B / 128 B swizzle) and of `mma.sync.m8n8k4.f64`, comparing every output element with the plain
contraction `out[o, r, i] = sum_c T[r, c] in[o, c, i]` (orthogonal.py:214-277 of the reference).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("emu") / "fold_emu")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "emu", "fold_emu.cpp")])
    return exe


def test_synthetic_tables_all_variants(emu):
    """160 fold cases (OUT/IN fold x NN/NT order x both parities x sizes with ragged tiles, odd mode counts, padding) and
    18 cases of the complex-last-axis variant CPLX_NT (arbitrary real table, odd extents)."""
    r = subprocess.run([emu], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "FOLD EMU: ALL OK" in r.stdout
    assert r.stdout.count("\nok  ") + r.stdout.startswith("ok  ") >= 190
    assert r.stdout.count("variant 4 (complex last axis)") == 18


def test_random_shapes(emu):
    """80 random fold shapes inside the engine's envelope (extents >= 16, ragged tiles, k tails) + 20 of the complex variant."""
    r = subprocess.run([emu, "--fuzz", "80", "3"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "FOLD EMU: ALL OK" in r.stdout, r.stdout[-3000:]
    assert r.stdout.count("\nok  ") + r.stdout.startswith("ok  ") == 100


@pytest.mark.parametrize("space,n", [("Legendre", 64), ("Legendre", 128), ("ChebyshevU", 64), ("Ultraspherical", 32),
                                     ("Chebyshev", 80), ("Legendre", 320), ("Legendre", 512), ("Legendre", 1024)])
def test_host_tables_fold(emu, tmp_path, space, n):
    """The tables the host really builds (nodes mirrored only to an ulp) are recognised, and the folded result stays
    within 1e-12 of the contraction with the table as given — for random inputs relative to the result, and for single high
    modes relative to that mode's own result (n >= 320 needs the asymmetry-correction k-tiles for that)."""
    import jaxfun_b200 as jf
    from jaxfun_b200 import _lib as L
    sp = getattr(jf, space)(n)
    cases = [(L.OP_BACKWARD, 0), (L.OP_FORWARD, 0), (L.OP_SCALAR_PRODUCT, 0), (L.OP_BACKWARD_PRIMITIVE, 1),
             (L.OP_BACKWARD_PRIMITIVE, 2)]
    for op, k in cases:
        T = np.ascontiguousarray(sp._dense_table(op, n, n, k), dtype=np.float64)
        path = tmp_path / f"t_{op}_{k}.bin"
        T.tofile(path)
        r = subprocess.run([emu, "--table", str(path), str(T.shape[0]), str(T.shape[1])], capture_output=True, text=True,
                           timeout=600)
        assert r.returncode == 0 and "ALL OK" in r.stdout, (space, n, op, k, r.stdout[-2000:])


def test_composite_and_padded_tables_fold(emu, tmp_path):
    import jaxfun_b200 as jf
    from jaxfun_b200 import _lib as L
    from jaxfun_b200.galerkin.composite import FunctionSpace
    D = FunctionSpace(64, jf.Legendre, bcs={"left": {"D": 0}, "right": {"D": 0}})
    tabs = [D._dense_table(L.OP_BACKWARD, D.dim, 64, 0), D._dense_table(L.OP_SCALAR_PRODUCT, D.dim, 64, 0),
            jf.Legendre(64)._dense_table(L.OP_BACKWARD, 64, 96, 0)]
    for i, T in enumerate(tabs):
        T = np.ascontiguousarray(T, dtype=np.float64)
        path = tmp_path / f"c{i}.bin"
        T.tofile(path)
        r = subprocess.run([emu, "--table", str(path), str(T.shape[0]), str(T.shape[1])], capture_output=True, text=True,
                           timeout=600)
        assert r.returncode == 0 and "ALL OK" in r.stdout, (i, T.shape, r.stdout[-2000:])
