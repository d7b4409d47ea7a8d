"""Golden data for the wavenumber-batched banded solver, made by the reference's OWN functions.

Runs only where /root/reference exists (the build container):

    python tests/golden/make_golden_banded.py

The reference's `la` package cannot be imported as a whole here (flax.nnx classes), so the two pure functions of the path are
taken, unmodified, from where they lie — `_lu_banded_no_pivot_kernel` (la/diamatrix.py:1937-1973) and
`_make_wavenumber_vmap_solve` (la/tpmatrix.py:590-683) — by compiling exactly their `def` blocks out of the reference files
and executing them on the numpy stand-in for jax (tools/jaxshim).  The batched assembly and factor extraction around them
(la/tpmatrix.py:743-768, 1345-1347: einsum, band scatter, row slices) are three lines restated below.
Output: tests/golden/reference_banded.npz (inputs W, P, offsets, rhs and the reference's band_lu and solution per case).
"""
import ast
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tools", "jaxshim"))
import jax  # noqa: E402  (the stand-in)
import jax.numpy as jnp  # noqa: E402

assert "jaxshim" in jax.__file__
REF = "/root/reference/src/jaxfun/la"


def reference_function(path, name):
    """Compile the `def name` block of a reference file, as it stands, into a namespace that holds only jax / jnp."""
    src = open(path).read()
    tree = ast.parse(src)
    node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    mod = ast.Module(body=[node], type_ignores=[])
    from collections.abc import Callable
    from typing import Any
    ns = {"jax": jax, "jnp": jnp, "np": np, "Array": jax.Array, "Any": Any, "Callable": Callable, "__name__": "reference_slice"}
    code = compile(ast.fix_missing_locations(mod), path, "exec", flags=__import__("__future__").annotations.compiler_flag)
    exec(code, ns)
    return ns[name]


lu_kernel = reference_function(os.path.join(REF, "diamatrix.py"), "_lu_banded_no_pivot_kernel")
make_solve = reference_function(os.path.join(REF, "tpmatrix.py"), "_make_wavenumber_vmap_solve")


def reference_factor_and_solve(W, P, offsets, rhs2d):
    data = jnp.einsum("tf,tdp->fdp", jnp.asarray(W), jnp.asarray(P))                 # tpmatrix.py:1347
    n_sys, _, n = data.shape
    p = max((-o for o in offsets if o < 0), default=0)
    q = max((o for o in offsets if o > 0), default=0)
    center, bw = p, p + q + 1
    band = jnp.zeros((n_sys, bw, n), dtype=data.dtype)
    band = band.at[:, [center + o for o in offsets], :].set(data)                    # tpmatrix.py:759-764
    band_lu = jax.vmap(lambda b: lu_kernel(b, p, q, center))(band)                   # tpmatrix.py:765-767
    L = np.stack([np.asarray(band_lu)[:, center + o, :] for o in range(-p, 0)], axis=1) if p else np.zeros((n_sys, 0, n))
    U = np.stack([np.asarray(band_lu)[:, center + o, :] for o in range(0, q + 1)], axis=1)
    solve = make_solve(tuple(range(-p, 0)), tuple(range(0, q + 1)), n, data.dtype)   # tpmatrix.py:922-924 (no pruning)
    x = solve(jnp.asarray(L), jnp.asarray(U), jnp.asarray(rhs2d))
    return np.asarray(band_lu), np.asarray(x)


def case(seed, n, n_sys, offsets, n_terms=2, cplx_band=False, cplx_rhs=True):
    rng = np.random.default_rng(seed)
    P = rng.standard_normal((n_terms, len(offsets), n))
    W = rng.standard_normal((n_terms, n_sys))
    if cplx_band:
        P = P + 1j * rng.standard_normal(P.shape)
        W = W + 1j * rng.standard_normal(W.shape)
    # diagonally dominant main diagonal (what LU without pivoting is meant for): |B_kk| > sum of the other entries
    main = offsets.index(0)
    B = np.einsum("tf,tdp->fdp", W, P)
    need = np.abs(B).sum(axis=1).max() + 1.0
    P[0, main, :] += need
    W[0, :] = np.abs(W[0, :]) + 1.0
    rhs = rng.standard_normal((n_sys, n))
    if cplx_rhs:
        rhs = rhs + 1j * rng.standard_normal((n_sys, n))
    return W, P, rhs


CASES = {
    "penta_2d": dict(seed=1, n=30, n_sys=16, offsets=(-2, 0, 2)),
    "tri_real": dict(seed=2, n=17, n_sys=5, offsets=(-1, 0, 1), cplx_rhs=False),
    "wide_upper": dict(seed=3, n=24, n_sys=9, offsets=(-2, 0, 2, 4, 6)),
    "nona": dict(seed=4, n=33, n_sys=12, offsets=(-4, -2, 0, 2, 4), n_terms=3),
    "complex_band": dict(seed=5, n=20, n_sys=8, offsets=(-2, -1, 0, 1, 2), cplx_band=True),
    "upper_only": dict(seed=6, n=12, n_sys=4, offsets=(0, 1, 3)),
    "dense_upper": dict(seed=7, n=14, n_sys=6, offsets=(-2, 0, 2, 4, 6, 8, 10, 12)),
    "lower_only": dict(seed=8, n=11, n_sys=3, offsets=(-3, -1, 0)),
}

if __name__ == "__main__":
    out = {}
    for name, kw in CASES.items():
        offsets = kw["offsets"]
        W, P, rhs = case(**kw)
        band_lu, x = reference_factor_and_solve(W, P, offsets, rhs)
        out[name + "/W"], out[name + "/P"], out[name + "/rhs"] = W, P, rhs
        out[name + "/offsets"] = np.array(offsets, dtype=np.int64)
        out[name + "/band_lu"], out[name + "/x"] = band_lu, x
        print(name, band_lu.shape, x.shape, float(np.abs(x).max()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "reference_banded.npz"), **out)
