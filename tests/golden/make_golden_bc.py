"""Golden data for the boundary-lifting basis of DirectSum spaces, made by the reference's OWN functions.

Runs only where /root/reference exists (the build container):

    python tests/golden/make_golden_bc.py

`jaxfun.galerkin.composite.get_bc_basis` (composite.py:835-896) and `BoundaryConditions` (composite.py:40-118) are
imported unmodified from the reference on the numpy stand-in for jax (tools/jaxshim) and evaluated for a set of
inhomogeneous boundary conditions; the resulting lifting matrices S (rows = lifting functions, columns = orthogonal
modes), the ordered names and the ordered values are written to tests/golden/reference_bc_basis.json.
"""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tools", "jaxshim"))
import load_reference  # noqa: E402

load_reference.mount()
comp = importlib.import_module("jaxfun.galerkin.composite")
from jaxfun.galerkin.Chebyshev import Chebyshev  # noqa: E402
from jaxfun.galerkin.Legendre import Legendre  # noqa: E402

BCS = [
    {"left": {"D": 1.0}, "right": {"D": -2.0}},
    {"left": {"D": 0.5}, "right": {"N": 1.5}},
    {"left": {"N": 1.0}, "right": {"D": 3.0}},
    {"left": {"N": 1.0}, "right": {"N": -1.0}},
    {"left": {"D": 1.0, "N": 0.5}, "right": {"D": 0, "N": 2}},
    {"left": {"D": 2.0}},
    {"right": {"D": -1.0, "N": 0.25}},
    {"left": {"D": 1.0, "N": 0.0, "N2": -1.0}, "right": {"D": 2.0, "N": 0.0, "N2": 0.5}},
    {"left": {"R": (0.3, 1.0)}, "right": {"D": -1.0}},              # Robin: u + 0.3 u' = 1 on the left
    {"left": {"R": (2.0, 0.5)}, "right": {"R": (-1.0, 0.25)}},
]

cases = []
for name, cls in (("Legendre", Legendre), ("Chebyshev", Chebyshev)):
    for bcs in BCS:
        B = comp.BoundaryConditions(bcs)
        orth = cls(B.num_bcs() + B.num_derivatives())
        S = comp.get_bc_basis(B, orth)
        cases.append({"space": name, "bcs": bcs, "names": B.orderednames(), "vals": [float(v) for v in B.orderedvals()],
                      "num_bcs": int(B.num_bcs()), "num_derivatives": int(B.num_derivatives()),
                      "S": np.array(S.tolist(), dtype=float).tolist()})
out = os.path.join(ROOT, "tests", "golden", "reference_bc_basis.json")
json.dump({"source": "jaxfun.galerkin.composite.get_bc_basis / BoundaryConditions of the reference, unmodified", "cases": cases},
          open(out, "w"), indent=1)
print("wrote", out, len(cases), "cases")
