"""Golden vectors produced by the REFERENCE's own source files (tests/golden/make_golden.py: the
reference modules executed on a numpy stand-in for jax) replayed through

  * the CPU oracle  (not gpu)  -> pins the oracle restatement against the reference's code;
  * the CUDA path   (gpu)      -> parity of libjfx.so (through the C ABI) with the reference.

Tolerances: quadrature nodes / weights / wavenumbers bit-exact (oracle) — 1e-12 relative in
float64 / complex128 for transforms (BASELINE.json), against the max-norm of the golden result.
"""
import json
import os

import numpy as np
import pytest
import sympy as sp

import jaxfun_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "reference_vectors.npz"))
with open(os.path.join(HERE, "golden", "reference_vectors.json")) as f:
    M = json.load(f)

TOL = 1e-12


def relerr(a, b):
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def build(mod, spec):
    kw = dict(spec.get("kw", {}))
    if spec.get("domain") is not None:
        kw["domain"] = tuple(spec["domain"])
    return getattr(mod, spec["basis"])(spec["N"], **kw)


def ids(cases):
    out = []
    for c in cases:
        sp_ = c["spaces"] if "spaces" in c else [c["space"]]
        out.append(c["key"] + ":" + "x".join(f"{s['basis'][:4]}{s['N']}" for s in sp_) +
                   ("d" if any(s.get("domain") for s in sp_) else ""))
    return out


# ==================================================================================================
# oracle vs reference vectors (CPU)
# ==================================================================================================
@pytest.mark.parametrize("n", M["leggauss"])
def test_oracle_leggauss_bit_exact(n):
    """utils/fastgl.py:562-567 — node tables must be bit-identical."""
    assert np.array_equal(np.asarray(O.leggauss(n)), G[f"leggauss/{n}"])


@pytest.mark.parametrize("case", M["cases_1d"], ids=ids(M["cases_1d"]))
def test_oracle_1d(case):
    key, S = case["key"], build(O, case["space"])
    x, w = S.quad_points_and_weights()
    N = case["space"]["N"]
    assert np.array_equal(np.asarray(x, dtype=float), G[f"{key}/x"]), "quadrature nodes must be bit-exact"
    assert np.array_equal(np.asarray(w, dtype=float) * np.ones(N), G[f"{key}/w"]), "weights must be bit-exact"
    assert np.array_equal(np.asarray(S.mesh(), dtype=float), G[f"{key}/mesh"])
    assert relerr(np.asarray(S.norm_squared(), dtype=float) * np.ones(N), G[f"{key}/norm_squared"]) < 1e-15
    assert relerr(S.vandermonde(np.asarray(x)), G[f"{key}/vandermonde"]) < 1e-13
    c, u = G[f"{key}/c"], G[f"{key}/backward"]
    assert relerr(S.backward(c), u) < TOL
    assert relerr(S.backward(c, N=case["pad"]), G[f"{key}/backward_pad"]) < TOL
    assert relerr(S.forward(u), G[f"{key}/forward"]) < TOL
    assert relerr(S.scalar_product(u), G[f"{key}/scalar_product"]) < TOL
    for k in (1, 2):
        assert relerr(S.backward_primitive(c, k), G[f"{key}/backward_primitive{k}"]) < TOL
        assert relerr(S.derivative_coeffs(c, k), G[f"{key}/derivative_coeffs{k}"]) < TOL
    if case["space"]["basis"] == "Fourier":
        assert np.array_equal(S.wavenumbers(), G[f"{key}/wavenumbers"])
        assert np.array_equal(S.wavenumbers(eliminate_highest_freq=True), G[f"{key}/wavenumbers_elim"])
    elif case["space"]["basis"] != "ChebyshevU":
        # complex lines through a real basis = the real transform of re and im (linear extension).
        # ChebyshevU is excluded: the reference's DST takes .imag of an FFT (utils/common.py:197-230), so
        # it silently drops the imaginary part of complex input — not a usage the reference supports.
        assert relerr(S.backward(G[f"{key}/cb"], axis=-1), G[f"{key}/backward_cb"]) < TOL


@pytest.mark.parametrize("case", M["cases_nd"], ids=ids(M["cases_nd"]))
def test_oracle_nd(case):
    key = case["key"]
    T = O.TensorProductSpace(*[build(O, s) for s in case["spaces"]])
    c, u = G[f"{key}/c"], G[f"{key}/backward"]
    assert relerr(T.backward(c), u) < TOL
    assert relerr(T.forward(u), G[f"{key}/forward"]) < TOL
    assert relerr(T.scalar_product(u), G[f"{key}/scalar_product"]) < TOL
    assert relerr(T.backward_primitive(c, tuple(case["k"])), G[f"{key}/backward_primitive"]) < TOL
    assert relerr(T.backward(c, N=tuple(case["pad"])), G[f"{key}/backward_pad"]) < TOL


def _sym_expr(case, u, xs):
    ns = {"u": u, "I": sp.I, "Abs": sp.Abs}
    for name, s in zip("xyz", xs):
        ns[name] = s
    return eval(case["expr"], {"__builtins__": {}}, ns)  # noqa: S307 - fixture strings we wrote ourselves


def _oracle_nonlinear(case, final):
    spaces = [build(O, s) for s in case["spaces"]]
    V = spaces[0] if len(spaces) == 1 else O.TensorProductSpace(*spaces)
    d = len(spaces)
    xs = sp.symbols("x y z", real=True)[:d]
    u = sp.Function("u")(*xs)
    expr = sp.expand(_sym_expr(case, u, xs).doit())
    # leaves: every derivative order tuple that appears
    leaves = {}

    def sub(node):
        if node == u:
            return leaves.setdefault((0,) * d, sp.Symbol("L" + "0" * d))
        if isinstance(node, sp.Derivative) and node.expr == u:
            k = tuple(node.variables.count(s) for s in xs)
            return leaves.setdefault(k, sp.Symbol("L" + "".join(map(str, k))))
        if not node.args:
            return node
        return node.func(*[sub(a) for a in node.args])

    e2 = sub(expr)
    ks = list(leaves)
    fn = sp.lambdify([leaves[k] for k in ks], e2, modules="numpy")
    N = case["N"]
    if d == 1:
        ks = [k[0] for k in ks]
    elif N is not None:
        N = tuple(N)
    return O.nonlinear_rhs(V, ks, fn, G[f"{case['key']}/c"], N=N, final=final)


@pytest.mark.parametrize("case", M["cases_nonlinear"], ids=ids(M["cases_nonlinear"]))
def test_oracle_nonlinear(case):
    key = case["key"]
    assert relerr(_oracle_nonlinear(case, "forward"), G[f"{key}/forward"]) < TOL
    assert relerr(_oracle_nonlinear(case, "scalar_product"), G[f"{key}/scalar_product"]) < TOL


# ==================================================================================================
# CUDA path vs reference vectors (GPU, through the C ABI)
# ==================================================================================================
def _dev(x, cuda):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x)).to(cuda)


@pytest.mark.gpu
@pytest.mark.parametrize("case", M["cases_1d"], ids=ids(M["cases_1d"]))
def test_cuda_1d(cuda, case):
    import jaxfun_b200 as jf
    key, S = case["key"], build(jf, case["space"])
    N = case["space"]["N"]
    x, w = S.quad_points_and_weights()
    # host tables of the product: same bits as the reference's
    assert np.array_equal(np.asarray(x, dtype=float), G[f"{key}/x"])
    assert np.array_equal(np.asarray(w, dtype=float) * np.ones(N), G[f"{key}/w"])
    c, u = G[f"{key}/c"], G[f"{key}/backward"]
    assert relerr(S.backward(_dev(c, cuda)), u) < TOL
    assert relerr(S.backward(_dev(c, cuda), N=case["pad"]), G[f"{key}/backward_pad"]) < TOL
    assert relerr(S.forward(_dev(u, cuda)), G[f"{key}/forward"]) < TOL
    assert relerr(S.scalar_product(_dev(u, cuda)), G[f"{key}/scalar_product"]) < TOL
    for k in (1, 2):
        assert relerr(S.backward_primitive(_dev(c, cuda), k), G[f"{key}/backward_primitive{k}"]) < TOL
        assert relerr(S.derivative_coeffs(_dev(c, cuda), k), G[f"{key}/derivative_coeffs{k}"]) < TOL
    if case["space"]["basis"] == "Fourier":
        assert np.array_equal(np.asarray(S.wavenumbers()), G[f"{key}/wavenumbers"])
    elif case["space"]["basis"] != "ChebyshevU":
        assert relerr(S.backward(_dev(G[f"{key}/cb"], cuda)), G[f"{key}/backward_cb"]) < TOL
    # host path (numpy in -> numpy out through jfx_execute_host)
    assert relerr(S.backward(c), u) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("case", M["cases_nd"], ids=ids(M["cases_nd"]))
def test_cuda_nd(cuda, case):
    import jaxfun_b200 as jf
    key = case["key"]
    T = jf.TensorProduct(*[build(jf, s) for s in case["spaces"]])
    c, u = G[f"{key}/c"], G[f"{key}/backward"]
    assert relerr(T.backward(_dev(c, cuda)), u) < TOL
    assert relerr(T.forward(_dev(u, cuda)), G[f"{key}/forward"]) < TOL
    assert relerr(T.scalar_product(_dev(u, cuda)), G[f"{key}/scalar_product"]) < TOL
    assert relerr(T.backward_primitive(_dev(c, cuda), tuple(case["k"])), G[f"{key}/backward_primitive"]) < TOL
    assert relerr(T.backward(_dev(c, cuda), N=tuple(case["pad"])), G[f"{key}/backward_pad"]) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("case", M["cases_nonlinear"], ids=ids(M["cases_nonlinear"]))
def test_cuda_nonlinear(cuda, case):
    import jaxfun_b200 as jf
    from jaxfun_b200.integrators.nonlinear import NonlinearTerm, field
    key = case["key"]
    spaces = [build(jf, s) for s in case["spaces"]]
    V = spaces[0] if len(spaces) == 1 else jf.TensorProduct(*spaces)
    u, xs = field(V)
    expr = _sym_expr(case, u, xs)
    N = case["N"]
    if N is not None and len(spaces) == 1:
        N = (N,)
    for final in ("forward", "scalar_product"):
        nl = NonlinearTerm(V, expr, final=final, N=N)
        assert relerr(nl(_dev(G[f"{key}/c"], cuda)), G[f"{key}/{final}"]) < TOL
