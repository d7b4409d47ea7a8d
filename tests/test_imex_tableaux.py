"""Order of accuracy of the IMEX tableaux (restated from the literature, jaxfun_b200/integrators/tableau.py)
through the oracle's IMEX step on the split linear test problem y' = a y + b y (a implicit, b explicit).
CPU only: pins the constants without a GPU."""
import importlib.util
import os
import sys

import numpy as np
import pytest

import jaxfun_oracle as O

_spec = importlib.util.spec_from_file_location(
    "jfx_tableau", os.path.join(os.path.dirname(__file__), "..", "jaxfun_b200", "integrators", "tableau.py"))
T = importlib.util.module_from_spec(_spec)
sys.modules["jfx_tableau"] = T   # dataclasses resolve annotations through sys.modules
_spec.loader.exec_module(T)


@pytest.mark.parametrize("name,order", [("IMEX_EULER", 1), ("ARS222", 2), ("ARS443", 3)])
def test_tableau_order(name, order):
    tab = getattr(T, name)
    a, b = -2.0 + 0.5j, 0.7 - 0.3j
    M = np.array([1.0])
    y0 = np.array([1.0 + 0.0j])
    errs = []
    for steps in (20, 40, 80):
        dt = 1.0 / steps
        y = y0
        for _ in range(steps):
            y = O.imex_rk_step(y, dt, tab, M, M * a, lambda v: b * v)
        errs.append(abs(y[0] - np.exp(a + b)))
    rates = [np.log2(errs[i] / errs[i + 1]) for i in range(2)]
    assert all(r > order - 0.25 for r in rates), (errs, rates)


def test_tableau_structure():
    assert T.IMEX_EULER.is_stiffly_accurate is False or T.IMEX_EULER.stages == 2
    assert T.ARS222.implicit_is_stiffly_accurate and T.ARS443.implicit_is_stiffly_accurate
    assert T.ARS222.is_stiffly_accurate and T.ARS443.is_stiffly_accurate
    assert len(T.ARS222.distinct_diagonal_coeffs) == 1 and len(T.ARS443.distinct_diagonal_coeffs) == 1
    for tab in (T.IMEX_EULER, T.ARS222, T.ARS443):
        for bt in (tab.explicit, tab.implicit):
            assert abs(sum(bt.b) - 1.0) < 1e-14
            for i in range(bt.stages):
                assert abs(sum(bt.A[i]) - bt.c[i]) < 1e-14


@pytest.mark.parametrize("name", ["IMEX_EULER", "ARS222", "ARS443"])
def test_tableau_equals_the_reference_constants(name):
    """The restated tableaux against the reference's own `integrators/tableau.py` (dumped by
    tests/golden/make_golden_tableaux.py), including the derived properties the IMEX step branches on."""
    import json
    G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_tableaux.json")))["tableaux"][name]
    tab = getattr(T, name)
    assert tab.stages == G["stages"]
    for part in ("explicit", "implicit"):
        mine = getattr(tab, part)
        assert np.abs(np.array(mine.A, dtype=float) - np.array(G[part]["A"])).max() < 1e-15
        assert np.abs(np.array(mine.b, dtype=float) - np.array(G[part]["b"])).max() < 1e-15
        assert np.abs(np.array(mine.c, dtype=float) - np.array(G[part]["c"])).max() < 1e-15
    assert tab.is_stiffly_accurate == G["is_stiffly_accurate"]
    assert tab.implicit_is_stiffly_accurate == G["implicit_is_stiffly_accurate"]
    assert np.allclose(np.array(tab.distinct_diagonal_coeffs, dtype=float), np.array(G["distinct_diagonal_coeffs"]), rtol=0, atol=1e-15)
