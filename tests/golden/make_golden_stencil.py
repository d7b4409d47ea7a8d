"""Golden stencils of Composite bases, made by the reference's OWN `get_stencil_matrix` (composite.py:765-838), imported
unmodified on the numpy stand-in for jax (tools/jaxshim).  Runs only where /root/reference exists:

    python tests/golden/make_golden_stencil.py   ->  tests/golden/reference_stencils.json

For every (family, homogeneous boundary conditions) the symbolic stencil {shift: expr(n)} is evaluated for the rows
n = 0 .. N - nb - 1 of an N = 12 space."""
import importlib
import json
import os
import sys

import numpy as np
import sympy as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tools", "jaxshim"))
import load_reference  # noqa: E402

load_reference.mount()
comp = importlib.import_module("jaxfun.galerkin.composite")
from jaxfun.galerkin.Chebyshev import Chebyshev  # noqa: E402
from jaxfun.galerkin.Legendre import Legendre  # noqa: E402

N = 12
BCS = [
    {"left": {"D": 0}, "right": {"D": 0}},
    {"left": {"N": 0}, "right": {"N": 0}},
    {"left": {"D": 0}, "right": {"N": 0}},
    {"left": {"N": 0}, "right": {"D": 0}},
    {"left": {"D": 0, "N": 0}, "right": {"D": 0, "N": 0}},
    {"left": {"D": 0}},
    {"right": {"D": 0}},
    {"left": {"D": 0, "N": 0}, "right": {"D": 0}},
    {"left": {"R": (0.3, 0)}, "right": {"D": 0}},                   # Robin: u + 0.3 u' = 0 on the left (0.5 makes row 0 singular)
    {"left": {"R": (2.0, 0)}, "right": {"R": (-1.0, 0)}},
]
cases = []
for name, cls in (("Legendre", Legendre), ("Chebyshev", Chebyshev)):
    for bcs in BCS:
        B = comp.BoundaryConditions(bcs)
        st = comp.get_stencil_matrix(B, cls(N))
        rows = N - B.num_bcs()
        out = {}
        for k, v in st.items():
            e = sp.sympify(v)
            vals = []
            for i in range(rows):
                v = e.subs({s: i for s in e.free_symbols})
                if not v.is_real or v in (sp.nan, sp.zoo, sp.oo, -sp.oo):       # removable singularity of the closed form at small n
                    (sym,) = e.free_symbols
                    v = sp.limit(e, sym, i)
                vals.append(float(v))
            out[str(int(k))] = vals
        cases.append({"space": name, "bcs": bcs, "N": N, "rows": rows, "stencil": out})
path = os.path.join(ROOT, "tests", "golden", "reference_stencils.json")
json.dump({"source": "jaxfun.galerkin.composite.get_stencil_matrix of the reference, unmodified", "cases": cases}, open(path, "w"), indent=1)
print("wrote", path, len(cases), "cases")
