"""Golden ETDRK4 coefficient functions, made by the reference's OWN `_phi1`, `_phi2`, `_phi3` and
`_etdrk4_nonlinear_weights` (integrators/etdrk4.py:21-52).  The module itself cannot be imported here (its base class pulls
in flax), so the four function definitions are extracted from the source file with `ast` and executed, unmodified, on the
numpy stand-in for jax.numpy (tools/jaxshim).  Runs only where /root/reference exists:

    python tests/golden/make_golden_etd.py   ->  tests/golden/reference_etd.npz
"""
import ast
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tools", "jaxshim"))
import jax.numpy as jnp  # noqa: E402  (the stand-in)

SRC = "/root/reference/src/jaxfun/integrators/etdrk4.py"
tree = ast.parse(open(SRC).read())
want = {"_phi1", "_phi2", "_phi3", "_etdrk4_nonlinear_weights"}
mod = ast.Module(body=[n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in want], type_ignores=[])
ns = {"jnp": jnp, "Array": object}
exec(compile(mod, SRC, "exec"), ns)   # annotations are strings / names resolved lazily: `from __future__ import annotations` not needed

# real decay rates (diffusion), tiny arguments (series branches), imaginary symbols (dispersive), mixed
z = np.concatenate([-np.logspace(-12, 3, 61), np.array([0.0, 1e-9, -1e-8, 3e-7, -2e-6, 5e-6, 2e-5]),
                    1j * np.linspace(-40, 40, 41), (-1 + 2j) * np.logspace(-10, 2, 25)]).astype(complex)
phi1, phi2, phi3 = ns["_phi1"](z), ns["_phi2"](z), ns["_phi3"](z)
f1, f2, f3 = ns["_etdrk4_nonlinear_weights"](phi1, phi2, phi3)
out = os.path.join(ROOT, "tests", "golden", "reference_etd.npz")
np.savez(out, z=z, phi1=np.asarray(phi1), phi2=np.asarray(phi2), phi3=np.asarray(phi3), f1=np.asarray(f1), f2=np.asarray(f2),
         f3=np.asarray(f3), q=0.5 * np.asarray(ns["_phi1"](z / 2)))
print("wrote", out, z.shape)
