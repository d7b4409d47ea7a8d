"""Golden vectors at n = 128 / 256 (tests/golden/make_golden_large.py: the reference's own modules on the numpy
stand-in for jax) — the sizes at which the FFT / DCT kernels and the tensor-core contraction run.
CPU: the oracle; GPU: the CUDA path through the C ABI.  Tolerance 1e-12 relative (fp64 / c128)."""
import json
import os

import numpy as np
import pytest

import jaxfun_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
_npz = os.path.join(HERE, "golden", "reference_vectors_large.npz")
if not os.path.exists(_npz):   # pragma: no cover
    pytest.skip("large golden vectors not generated", allow_module_level=True)
G = np.load(_npz)
M = json.load(open(os.path.join(HERE, "golden", "reference_vectors_large.json")))
TOL = 1e-12


def rel(a, b):
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return float(np.abs(np.asarray(a) - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("case", M["cases_1d"], ids=[f"{c['basis']}{c['N']}" for c in M["cases_1d"]])
def test_oracle_1d_large(case):
    S = getattr(O, case["basis"])(case["N"])
    k = case["key"]
    x, w = S.quad_points_and_weights()
    assert np.array_equal(np.asarray(x, dtype=float), G[f"{k}/x"])
    assert np.array_equal(np.asarray(w, dtype=float) * np.ones(case["N"]), G[f"{k}/w"])
    c, u = G[f"{k}/c"], G[f"{k}/backward"]
    assert rel(S.backward(c), u) < TOL
    assert rel(S.backward(c, N=case["pad"]), G[f"{k}/backward_pad"]) < TOL
    assert rel(S.forward(u), G[f"{k}/forward"]) < TOL
    assert rel(S.scalar_product(u), G[f"{k}/scalar_product"]) < TOL
    assert rel(S.backward_primitive(c, 1), G[f"{k}/backward_primitive1"]) < 1e-11


@pytest.mark.parametrize("case", M["cases_nd"], ids=[c["key"] for c in M["cases_nd"]])
def test_oracle_nd_large(case):
    T = O.TensorProductSpace(*[getattr(O, b)(n) for b, n in case["factors"]])
    k = case["key"]
    assert rel(T.backward(G[f"{k}/c"]), G[f"{k}/backward"]) < TOL
    assert rel(T.forward(G[f"{k}/backward"]), G[f"{k}/forward"]) < TOL


def _dev(x, cuda):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x)).to(cuda)


@pytest.mark.gpu
@pytest.mark.parametrize("case", M["cases_1d"], ids=[f"{c['basis']}{c['N']}" for c in M["cases_1d"]])
def test_cuda_1d_large(cuda, case):
    import jaxfun_b200 as jf
    S = getattr(jf, case["basis"])(case["N"])
    k = case["key"]
    x, w = S.quad_points_and_weights()
    assert np.array_equal(np.asarray(x, dtype=float), G[f"{k}/x"])
    c, u = G[f"{k}/c"], G[f"{k}/backward"]
    # a single line and a batch of copies (the batch takes the tiled / tensor-core kernels)
    assert rel(S.backward(_dev(c, cuda)), u) < TOL
    cb = np.broadcast_to(c, (64,) + c.shape).copy()
    assert rel(S.backward(_dev(cb, cuda))[17], u) < TOL
    assert rel(S.backward(_dev(cb, cuda), N=case["pad"])[3], G[f"{k}/backward_pad"]) < TOL
    ub = np.broadcast_to(u, (64,) + u.shape).copy()
    assert rel(S.forward(_dev(ub, cuda))[40], G[f"{k}/forward"]) < TOL
    assert rel(S.scalar_product(_dev(ub, cuda))[5], G[f"{k}/scalar_product"]) < TOL
    assert rel(S.backward_primitive(_dev(cb, cuda), 1)[9], G[f"{k}/backward_primitive1"]) < 1e-11


@pytest.mark.gpu
@pytest.mark.parametrize("case", M["cases_nd"], ids=[c["key"] for c in M["cases_nd"]])
def test_cuda_nd_large(cuda, case):
    import jaxfun_b200 as jf
    T = jf.TensorProduct(*[getattr(jf, b)(n) for b, n in case["factors"]])
    k = case["key"]
    assert rel(T.backward(_dev(G[f"{k}/c"], cuda)), G[f"{k}/backward"]) < TOL
    assert rel(T.forward(_dev(G[f"{k}/backward"], cuda)), G[f"{k}/forward"]) < TOL
