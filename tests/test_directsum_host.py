"""Boundary-lifting basis of DirectSum spaces (composite.py:411-488, 502-638, 835-896 of the reference) on the host:
the oracle's symbolic restatement and the product's numeric construction against lifting matrices produced by the
reference's OWN `get_bc_basis` / `BoundaryConditions` (tests/golden/make_golden_bc.py)."""
import json
import os

import numpy as np
import pytest

import jaxfun_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "reference_bc_basis.json")))["cases"]
IDS = [f"{c['space']}-{'-'.join(c['names'])}" for c in CASES]


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_oracle_bc_basis_matches_reference(case):
    B = O.BoundaryConditions(case["bcs"])
    assert B.orderednames() == case["names"]
    assert [float(v) for v in B.orderedvals()] == case["vals"]
    assert B.num_bcs() == case["num_bcs"] and B.num_derivatives() == case["num_derivatives"]
    S = O.get_bc_basis(B, getattr(O, case["space"])(B.num_bcs() + B.num_derivatives()))
    ref = np.array(case["S"])
    assert S.shape == ref.shape
    assert np.abs(S - ref).max() < 1e-15


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_product_bc_basis_matches_reference(case):
    import jaxfun_b200 as jf
    from jaxfun_b200.galerkin.composite import bc_basis, ordered_bc_names
    S = bc_basis(case["bcs"], getattr(jf, case["space"]))
    ref = np.array(case["S"])
    assert S.shape == ref.shape
    assert np.abs(S - ref).max() < 1e-14
    assert ["LR"[side == "right"] + kind for side, kind in ordered_bc_names(case["bcs"])] == case["names"]


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_lift_satisfies_the_boundary_values(case):
    """The lift sum_b val_b B_b takes exactly the prescribed boundary values (derivatives in the reference coordinate)."""
    import jaxfun_b200 as jf
    from jaxfun_b200.galerkin.composite import _bc_functional, ordered_bc_names
    N = 16
    D = jf.FunctionSpace(N, getattr(jf, case["space"]), case["bcs"])
    assert type(D).__name__ == "DirectSum" and D.dim == N - case["num_bcs"]
    assert np.allclose(D.bnd_vals(), case["vals"])
    for (side, kind), val in zip(ordered_bc_names(case["bcs"]), case["vals"]):
        f = _bc_functional(D.orthogonal, side, kind, case["bcs"][side][kind])     # incl. the Robin combinations
        assert abs(f @ D.c_b - val) < 1e-12 * max(1.0, abs(val)), (side, kind)
        # and the homogeneous part contributes nothing there: every composite basis function satisfies the zero condition
        assert np.abs(D.a.S @ f).max() < 1e-9 * max(1.0, np.abs(f).max()), (side, kind)


def test_oracle_directsum_roundtrip_and_boundary_values():
    rng = np.random.default_rng(3)
    bcs = {"left": {"D": 1.0}, "right": {"D": -2.0}}
    D = O.DirectSum(O.Composite(24, O.Legendre, {0: 1, 2: -1}, domain=(0, 3)), bcs)
    c = rng.standard_normal((5, 22))
    u = D.backward(c)
    assert np.abs(D.forward(u) - c).max() < 1e-12
    ends = D.evaluate(np.array([0.0, 3.0]), c)
    assert np.abs(ends - np.array([1.0, -2.0])).max() < 1e-12
    assert np.abs(D.from_orthogonal(D.to_orthogonal(c)) - c).max() < 1e-12


@pytest.mark.parametrize("space", ["Legendre", "Chebyshev"])
def test_inhomogeneous_neumann_on_a_mapped_domain_is_in_physical_units(space):
    """BoundaryConditions(dict, domain) divides Neumann-type values by df**nd, df = 2 / (b - a) (composite.py:66-73 through
    functionspace.py:134): {'N': 1} on (0, 4) prescribes du/dx = 1 at x = 4, not dU/dX = 1 in the reference coordinate."""
    import jaxfun_b200 as jf
    dom = (0.0, 4.0)
    bcs = {"left": {"D": 0.5}, "right": {"N": 1.0, "N2": -3.0}}
    D = jf.FunctionSpace(14, getattr(jf, space), bcs, domain=dom)
    Do = O.DirectSum(O.Composite(14, getattr(O, space), D.a.stencil, domain=dom), bcs)
    assert np.abs(D.c_b - Do.c_b).max() < 1e-13
    V = D.orthogonal
    df = float(V.domain_factor)                                   # = 0.5
    for k, want in ((0, None), (1, 1.0), (2, -3.0)):
        # k-th physical derivative of the lift at the right end: df^k sum_j c_j P_j^(k)(1)
        got = df**k * (V.evaluate_basis_derivative(np.array([1.0]), k)[0] @ D.c_b)
        if want is not None:
            assert abs(got - want) < 1e-11, (k, got, want)
    left = V.eval_basis_functions(np.array([-1.0]))[0] @ D.c_b
    assert abs(left - 0.5) < 1e-12
    # the reference domain leaves the values alone
    D1 = jf.FunctionSpace(14, getattr(jf, space), bcs)
    assert np.allclose(D1.bnd_vals(), [0.5, 1.0, -3.0])
    assert np.allclose(D.bnd_vals(), [0.5, 1.0 / df, -3.0 / df**2])


# ---- homogeneous stencils: numeric derivation vs the reference's symbolic get_stencil_matrix ---------------------------
STENCILS = json.load(open(os.path.join(HERE, "golden", "reference_stencils.json")))["cases"]


@pytest.mark.parametrize("case", STENCILS, ids=[f"{c['space']}-{c['bcs']}" for c in STENCILS])
def test_numeric_stencil_matches_reference_get_stencil_matrix(case):
    """`stencil_from_bcs` (per-row solve of the boundary functionals) reproduces the values of the reference's symbolic
    stencil {shift: expr(n)} (composite.py:765-838; golden made by tests/golden/make_golden_stencil.py) to the last digits."""
    import jaxfun_b200 as jf
    from jaxfun_b200.galerkin.composite import stencil_from_bcs
    rows = case["rows"]
    ref = {int(k): np.array(v) for k, v in case["stencil"].items()}
    mine = stencil_from_bcs(case["bcs"], getattr(jf, case["space"])(case["N"]))
    for k in set(ref) | set(mine):
        a = ref.get(k, np.zeros(rows))
        b = np.asarray(mine.get(k, np.zeros(rows))) * np.ones(rows)
        assert np.abs(a - b).max() < 1e-13, (k, a, b)
