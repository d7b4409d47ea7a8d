"""ETDRK4 coefficient functions against the reference's own `_phi1.._phi3` / `_etdrk4_nonlinear_weights`
(integrators/etdrk4.py:21-52, executed unmodified by tests/golden/make_golden_etd.py): the oracle's restatement and the
product's host routine `etd_coefficients` (whose arrays the device combinations of ETDRK4.step consume)."""
import os

import numpy as np

import jaxfun_oracle as O

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_etd.npz"))


def _close(a, b):
    return np.abs(a - b).max() <= 1e-15 * max(1.0, np.abs(b).max())


def test_product_etd_coefficients_match_reference_functions():
    from jaxfun_b200.integrators import etd_coefficients
    z = G["z"]
    E, E2, Q, f1, f2, f3 = etd_coefficients(1.0, z)                 # dt * L = z
    assert _close(E, np.exp(z)) and _close(E2, np.exp(z / 2))      # etdrk4.py:111-112
    assert _close(Q, G["q"])                                        # Q = phi1(z/2)/2, etdrk4.py:113-116
    assert _close(f1, G["f1"]) and _close(f2, G["f2"]) and _close(f3, G["f3"])
    # a second step size through the same entry point
    E, E2, Q, f1, f2, f3 = etd_coefficients(0.5, 2.0 * z)
    assert _close(Q, G["q"]) and _close(f1, G["f1"]) and _close(f3, G["f3"])


def test_oracle_etd_coefficients_match_reference_functions():
    z = G["z"]
    c = O.etdrk4_coefficients(1.0, z)
    names = ("E", "E2", "Q", "f1", "f2", "f3")
    got = dict(zip(names, c)) if not isinstance(c, dict) else c
    assert _close(np.asarray(got["Q"]), G["q"])
    assert _close(np.asarray(got["f1"]), G["f1"]) and _close(np.asarray(got["f2"]), G["f2"]) and _close(np.asarray(got["f3"]), G["f3"])
