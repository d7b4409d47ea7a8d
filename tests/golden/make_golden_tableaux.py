"""Golden IMEX tableaux: the reference's own `integrators/tableau.py` (pure Python, loaded from where it lies) dumped to JSON.
Runs only where /root/reference exists:   python tests/golden/make_golden_tableaux.py  ->  tests/golden/reference_tableaux.json"""
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
spec = importlib.util.spec_from_file_location("ref_tableau", "/root/reference/src/jaxfun/integrators/tableau.py")
m = importlib.util.module_from_spec(spec)
sys.modules["ref_tableau"] = m
spec.loader.exec_module(m)
out = {}
for name in dir(m):
    t = getattr(m, name)
    if isinstance(t, m.IMEXTableau) and name in ("IMEX_EULER", "ARS222", "ARS443"):   # the schemes the product restates
        out[name] = {"stages": t.stages,
                     "explicit": {"A": [list(map(float, r)) for r in t.explicit.A], "b": list(map(float, t.explicit.b)),
                                  "c": list(map(float, t.explicit.c))},
                     "implicit": {"A": [list(map(float, r)) for r in t.implicit.A], "b": list(map(float, t.implicit.b)),
                                  "c": list(map(float, t.implicit.c))},
                     "is_stiffly_accurate": bool(t.is_stiffly_accurate),
                     "implicit_is_stiffly_accurate": bool(t.implicit_is_stiffly_accurate),
                     "distinct_diagonal_coeffs": list(map(float, t.distinct_diagonal_coeffs))}
path = os.path.join(ROOT, "tests", "golden", "reference_tableaux.json")
json.dump({"source": "jaxfun/integrators/tableau.py of the reference, unmodified", "tableaux": out}, open(path, "w"), indent=1)
print("wrote", path, sorted(out))
