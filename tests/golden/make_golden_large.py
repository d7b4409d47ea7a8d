#!/usr/bin/env python
"""Golden vectors at the sizes where the fast kernels and the tensor-core contraction run
(n = 128 / 256): same method as make_golden.py (the reference's own modules on the numpy stand-in for jax),
written to tests/golden/reference_vectors_large.npz.  Run from the repo root in the build container."""
from __future__ import annotations

import importlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "tools", "jaxshim"))

import load_reference  # noqa: E402

load_reference.mount()
import jax.numpy as jnp  # noqa: E402

tps = importlib.import_module("jaxfun.galerkin.tensorproductspace")
from jaxfun.galerkin.Chebyshev import Chebyshev  # noqa: E402
from jaxfun.galerkin.Fourier import Fourier  # noqa: E402
from jaxfun.galerkin.Legendre import Legendre  # noqa: E402

CLASSES = {"Chebyshev": Chebyshev, "Fourier": Fourier, "Legendre": Legendre}


def main():
    rng = np.random.default_rng(20261018)
    out, manifest = {}, {"cases_1d": [], "cases_nd": []}
    for i, (basis, N, pad) in enumerate([("Chebyshev", 256, 384), ("Chebyshev", 128, 128), ("Fourier", 256, 384),
                                         ("Fourier", 128, 128), ("Legendre", 128, 160), ("Legendre", 256, 256)]):
        S = CLASSES[basis](N)
        cplx = basis == "Fourier"
        c = rng.standard_normal(N) + (1j * rng.standard_normal(N) if cplx else 0)
        key = f"L1/{i}"
        out[f"{key}/c"] = c
        u = np.asarray(S.backward(jnp.asarray(c)))
        out[f"{key}/backward"] = u
        out[f"{key}/backward_pad"] = np.asarray(S.backward(jnp.asarray(c), N=pad))
        out[f"{key}/forward"] = np.asarray(S.forward(jnp.asarray(u)))
        out[f"{key}/scalar_product"] = np.asarray(S.scalar_product(jnp.asarray(u)))
        out[f"{key}/backward_primitive1"] = np.asarray(S.backward_primitive(jnp.asarray(c), 1))
        x, w = S.quad_points_and_weights()
        out[f"{key}/x"], out[f"{key}/w"] = np.asarray(x, dtype=float), np.asarray(w, dtype=float) * np.ones(N)
        manifest["cases_1d"].append({"key": key, "basis": basis, "N": N, "pad": pad})
        print("1d", basis, N, flush=True)
    for i, factors in enumerate([[("Chebyshev", 32), ("Chebyshev", 128)], [("Fourier", 16), ("Chebyshev", 128)],
                                 [("Legendre", 16), ("Fourier", 128)]]):
        T = tps.TensorProduct(*[CLASSES[b](n) for b, n in factors])
        shape = tuple(n for _, n in factors)
        cplx = any(b == "Fourier" for b, _ in factors)
        c = rng.standard_normal(shape) + (1j * rng.standard_normal(shape) if cplx else 0)
        key = f"Lnd/{i}"
        out[f"{key}/c"] = c
        u = np.asarray(T.backward(jnp.asarray(c)))
        out[f"{key}/backward"] = u
        out[f"{key}/forward"] = np.asarray(T.forward(jnp.asarray(u)))
        manifest["cases_nd"].append({"key": key, "factors": [[b, n] for b, n in factors], "complex": cplx})
        print("nd", factors, flush=True)
    np.savez_compressed(os.path.join(HERE, "reference_vectors_large.npz"), **out)
    with open(os.path.join(HERE, "reference_vectors_large.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print("wrote", len(out), "arrays", os.path.getsize(os.path.join(HERE, "reference_vectors_large.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
