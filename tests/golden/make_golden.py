#!/usr/bin/env python
"""Generate tests/golden/reference_vectors.npz (+ .json manifest) from the REFERENCE's own code.

jax is not installable in this image, so the reference cannot run as shipped.  What runs here is the
reference's unmodified source files (/root/reference/src/jaxfun/galerkin/*.py, utils/fastgl.py,
integrators/nonlinear.py, ...) executed on a small numpy stand-in for the jax API
(tools/jaxshim: jit = identity, vmap / scan / fori_loop = Python loops, jnp.* = numpy,
jax.scipy.fft.dct = scipy.fft.dct, float64 / complex128 throughout).  The vectors therefore pin every
formula, scaling, sign, ordering, padding and truncation rule of the reference; they do NOT pin
XLA's last-bit rounding of cos / FFT / dot ("Parity").

Run from the repo root in the build container (needs /root/reference):

    python tests/golden/make_golden.py

The manifest lists, per case, how to rebuild the spaces so that tests can replay the same inputs
through the oracle (CPU) and through the CUDA path (GPU).
"""
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
import sympy as sp  # noqa: E402
import jax.numpy as jnp  # noqa: E402

tps = importlib.import_module("jaxfun.galerkin.tensorproductspace")
args = importlib.import_module("jaxfun.galerkin.arguments")
_g = sys.modules["jaxfun.galerkin"]
_g.TestFunction, _g.TrialFunction, _g.JAXFunction = args.TestFunction, args.TrialFunction, args.JAXFunction
nl = importlib.import_module("jaxfun.integrators.nonlinear")
from jaxfun.galerkin.Chebyshev import Chebyshev  # noqa: E402
from jaxfun.galerkin.ChebyshevU import ChebyshevU  # noqa: E402
from jaxfun.galerkin.Fourier import Fourier  # noqa: E402
from jaxfun.galerkin.Jacobi import Jacobi  # noqa: E402
from jaxfun.galerkin.Legendre import Legendre  # noqa: E402
from jaxfun.galerkin.Ultraspherical import Ultraspherical  # noqa: E402
from jaxfun.utils.common import Domain  # noqa: E402
from jaxfun.utils.fastgl import leggauss  # noqa: E402

CLASSES = {"Chebyshev": Chebyshev, "ChebyshevU": ChebyshevU, "Fourier": Fourier, "Jacobi": Jacobi,
           "Legendre": Legendre, "Ultraspherical": Ultraspherical}


def make_space(spec):
    """spec = {"basis": name, "N": int, "kw": {...}, "domain": [a, b] | None}."""
    kw = dict(spec.get("kw", {}))
    if spec.get("domain") is not None:
        a, b = spec["domain"]
        kw["domain"] = Domain(sp.nsimplify(a), sp.nsimplify(b)) if spec.get("domain_sympy") else Domain(a, b)
    return CLASSES[spec["basis"]](spec["N"], **kw)


def rand(rng, shape, cplx):
    x = rng.standard_normal(shape)
    if cplx:
        x = x + 1j * rng.standard_normal(shape)
    return x


def main():
    out, manifest = {}, {"cases_1d": [], "cases_nd": [], "cases_nonlinear": [], "leggauss": []}
    rng = np.random.default_rng(20261017)

    # ---- Gauss-Legendre nodes (utils/fastgl.py): tabulated (n <= 100) and asymptotic (n > 100) branches
    for n in (1, 2, 3, 8, 13, 64, 100, 101, 128, 256, 257):
        xw = np.asarray(leggauss(n))
        out[f"leggauss/{n}"] = xw
        manifest["leggauss"].append(n)

    # ---- 1-D spaces ---------------------------------------------------------------------------
    specs = []
    for basis, kw in (("Chebyshev", {}), ("Legendre", {}), ("Jacobi", {"alpha": 1, "beta": 2}),
                      ("Jacobi", {"alpha": 1, "beta": 0}), ("ChebyshevU", {}), ("Ultraspherical", {"lambda_": 2})):
        for N in (8, 13, 16, 32, 64):
            specs.append({"basis": basis, "N": N, "kw": kw, "domain": None})
        specs.append({"basis": basis, "N": 16, "kw": kw, "domain": [-2.0, 3.0]})
    for N in (8, 16, 32, 64):
        specs.append({"basis": "Fourier", "N": N, "kw": {}, "domain": None})
    specs.append({"basis": "Fourier", "N": 16, "kw": {}, "domain": [0.0, 1.0]})
    for i, spec in enumerate(specs):
        S = make_space(spec)
        N = spec["N"]
        cplx = spec["basis"] == "Fourier"
        key = f"s1d/{i}"
        c = rand(rng, (N,), cplx)
        cb = rand(rng, (3, N), True)               # complex batch on a real basis = linear extension
        x, w = S.quad_points_and_weights()
        out[f"{key}/x"], out[f"{key}/w"] = np.asarray(x, dtype=float), np.asarray(w, dtype=float) * np.ones(N)
        out[f"{key}/mesh"] = np.asarray(S.mesh(), dtype=float)
        out[f"{key}/norm_squared"] = np.asarray(S.norm_squared(), dtype=float) * np.ones(N)
        out[f"{key}/vandermonde"] = np.asarray(S.vandermonde(jnp.asarray(x)))
        out[f"{key}/c"] = c
        u = np.asarray(S.backward(jnp.asarray(c)))
        out[f"{key}/backward"] = u
        pad = N + N // 2 + (N // 2) % 2 if cplx else N + 5
        out[f"{key}/backward_pad"] = np.asarray(S.backward(jnp.asarray(c), N=pad))
        out[f"{key}/forward"] = np.asarray(S.forward(jnp.asarray(u)))
        out[f"{key}/scalar_product"] = np.asarray(S.scalar_product(jnp.asarray(u)))
        for k in (1, 2):
            out[f"{key}/backward_primitive{k}"] = np.asarray(S.backward_primitive(jnp.asarray(c), k))
            out[f"{key}/derivative_coeffs{k}"] = np.asarray(S.derivative_coeffs(jnp.asarray(c), k))
        if cplx:
            out[f"{key}/wavenumbers"] = np.asarray(S.wavenumbers())
            out[f"{key}/wavenumbers_elim"] = np.asarray(S.wavenumbers(eliminate_highest_freq=True))
        else:
            # complex coefficients through a real basis, one line at a time (what vmap does)
            out[f"{key}/cb"] = cb
            out[f"{key}/backward_cb"] = np.stack([np.asarray(S.backward(jnp.asarray(r))) for r in cb])
        manifest["cases_1d"].append({"key": key, "space": spec, "pad": pad})

    # ---- tensor products ----------------------------------------------------------------------
    nd = [
        ([("Chebyshev", 16), ("Chebyshev", 16)], False),
        ([("Legendre", 8), ("Legendre", 12), ("Legendre", 10)], False),
        ([("Chebyshev", 16), ("Chebyshev", 8), ("Chebyshev", 32)], False),
        ([("Fourier", 8), ("Chebyshev", 16), ("Legendre", 6)], True),
        ([("Fourier", 16), ("Fourier", 8)], True),
        ([("Fourier", 8), ("Fourier", 8), ("Legendre", 5)], True),
        ([("Legendre", 9), ("Chebyshev", 16)], False),
        ([("Jacobi", 7), ("ChebyshevU", 9)], False),
    ]
    for i, (factors, cplx) in enumerate(nd):
        fspecs = [{"basis": b, "N": n, "kw": ({"alpha": 1, "beta": 2} if b == "Jacobi" else {}), "domain": None}
                  for b, n in factors]
        T = tps.TensorProduct(*[make_space(s) for s in fspecs])
        shape = tuple(n for _, n in factors)
        key = f"nd/{i}"
        c = rand(rng, shape, cplx)
        out[f"{key}/c"] = c
        u = np.asarray(T.backward(jnp.asarray(c)))
        out[f"{key}/backward"] = u
        out[f"{key}/forward"] = np.asarray(T.forward(jnp.asarray(u)))
        out[f"{key}/scalar_product"] = np.asarray(T.scalar_product(jnp.asarray(u)))
        k = tuple((j + 1) % 3 for j in range(len(shape)))
        out[f"{key}/backward_primitive"] = np.asarray(T.backward_primitive(jnp.asarray(c), k))
        pad = tuple(n + (4 if b == "Fourier" else 3) for b, n in factors)
        out[f"{key}/backward_pad"] = np.asarray(T.backward(jnp.asarray(c), N=pad))
        manifest["cases_nd"].append({"key": key, "spaces": fspecs, "k": list(k), "pad": list(pad), "complex": cplx})

    # ---- nonlinear terms (integrators/nonlinear.py + base.py:230-248) --------------------------------
    def nonlinear_case(name, fspecs, build_expr, cplx, scale=1.0, N=None, expr_str=""):
        spaces = [make_space(s) for s in fspecs]
        V = spaces[0] if len(spaces) == 1 else tps.TensorProduct(*spaces)
        shape = tuple(s["N"] for s in fspecs)
        jf = args.JAXFunction(jnp.zeros(shape, dtype=complex if cplx else float), V, name="u_jax")
        u = jf.doit()
        xs = V.system.base_scalars()
        expr = build_expr(u, xs)
        ev = nl.compile_nonlinear_evaluator(expr, V, u)
        c = scale * rand(rng, shape, cplx)
        phys = np.asarray(ev(jnp.asarray(c), N))
        key = f"nl/{name}"
        out[f"{key}/c"] = c
        out[f"{key}/physical"] = phys
        out[f"{key}/forward"] = np.asarray(V.forward(jnp.asarray(phys)))
        out[f"{key}/scalar_product"] = np.asarray(V.scalar_product(jnp.asarray(phys)))
        manifest["cases_nonlinear"].append({"key": key, "spaces": fspecs, "expr": expr_str, "complex": cplx,
                                            "N": None if N is None else list(N) if isinstance(N, tuple) else N})

    F = lambda n, dom=None: {"basis": "Fourier", "N": n, "kw": {}, "domain": dom}
    nonlinear_case("kdv", [F(32)], lambda u, xs: -u * sp.Derivative(u, xs[0]), True, expr_str="-u*u.diff(x)")
    nonlinear_case("kdv_sq", [F(32)], lambda u, xs: -sp.Derivative(u**2, xs[0]) / 2, True,
                   expr_str="-(u**2).diff(x)/2")
    nonlinear_case("burgers_leg", [{"basis": "Legendre", "N": 16, "kw": {}, "domain": None}],
                   lambda u, xs: -u * sp.Derivative(u, xs[0]), False, expr_str="-u*u.diff(x)")
    nonlinear_case("sq_cheb", [{"basis": "Chebyshev", "N": 16, "kw": {}, "domain": None}],
                   lambda u, xs: (u + sp.Derivative(u, xs[0]))**2, False, expr_str="(u+u.diff(x))**2")
    nonlinear_case("nls", [F(32)], lambda u, xs: -sp.I * sp.Abs(u)**2 * u, True, expr_str="-I*Abs(u)**2*u")
    nonlinear_case("gl", [F(16)], lambda u, xs: (1 + 1.5j) * u * sp.Abs(u)**2, True,
                   expr_str="(1+1.5j)*u*Abs(u)**2")
    nonlinear_case(
        "cahn_hilliard", [F(16, [0.0, 1.0]), F(16, [0.0, 1.0])],
        lambda u, xs: -(6 * u * (sp.Derivative(u, xs[0])**2 + sp.Derivative(u, xs[1])**2)
                        + 3 * u**2 * (sp.Derivative(u, xs[0], 2) + sp.Derivative(u, xs[1], 2))),
        True, scale=1e-2,
        expr_str="-(6*u*(u.diff(x)**2+u.diff(y)**2)+3*u**2*(u.diff(x,2)+u.diff(y,2)))")
    nonlinear_case("zk2d", [F(16), F(8)], lambda u, xs: -u * sp.Derivative(u, xs[0]), True,
                   expr_str="-u*u.diff(x)")
    nonlinear_case("burgers2d", [F(8), F(8)],
                   lambda u, xs: -(u * sp.Derivative(u, xs[0]) + u * sp.Derivative(u, xs[1])), True,
                   expr_str="-(u*u.diff(x)+u*u.diff(y))")
    nonlinear_case("kdv_padded", [F(32)], lambda u, xs: -u * sp.Derivative(u, xs[0]), True, N=48,
                   expr_str="-u*u.diff(x)")

    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    with open(os.path.join(HERE, "reference_vectors.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    sz = os.path.getsize(os.path.join(HERE, "reference_vectors.npz"))
    print(f"wrote {len(out)} arrays, {sz / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
