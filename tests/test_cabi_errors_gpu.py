"""Error behaviour and edge cases of the C ABI on a device: every bad request returns a negative jfx_status
with a message (never a crash, never a silent fallback); empty and size-1 arrays are valid requests."""
import ctypes as C

import numpy as np
import pytest
import torch

import jaxfun_b200 as jf
from jaxfun_b200 import _lib as L
from jaxfun_b200.engine import AxisSpec, Plan, _fill_plan_desc

pytestmark = pytest.mark.gpu


def create(desc):
    h = C.c_void_p()
    rc = L.load().jfx_plan_create(C.byref(desc), C.byref(h))
    msg = L.load().jfx_last_error().decode()
    if rc == 0:
        L.load().jfx_plan_destroy(h)
    return rc, msg


def test_plan_create_rejects_bad_descriptors(cuda):
    tab = np.eye(8)
    good, keep = _fill_plan_desc(L.OP_FORWARD, L.F64, (4, 8), [None, AxisSpec(L.BASIS_TABLE, table=tab)])
    assert create(good)[0] == 0
    d, _ = _fill_plan_desc(L.OP_FORWARD, L.F64, (4, 8), [None, AxisSpec(L.BASIS_TABLE, table=tab)])
    d.abi_version = 99
    rc, msg = create(d)
    assert rc == -1 and "ABI" in msg
    d, _ = _fill_plan_desc(L.OP_FORWARD, L.F64, (4, 8), [None, AxisSpec(L.BASIS_TABLE, table=tab)])
    d.axis[1].table = None
    assert create(d)[0] == -1
    d, _ = _fill_plan_desc(L.OP_FORWARD, L.F64, (4, 9), [None, AxisSpec(L.BASIS_TABLE, table=tab)])
    rc, msg = create(d)
    assert rc == -1 and "columns" in msg
    # Fourier axis on real data, fast basis at a length without a kernel, slab plans, nonlinear op
    d, _ = _fill_plan_desc(L.OP_FORWARD, L.F64, (16,), [AxisSpec(L.BASIS_FOURIER, n_modes=16, n_quad=16)])
    assert create(d)[0] == -1
    d, _ = _fill_plan_desc(L.OP_FORWARD, L.C128, (24,), [AxisSpec(L.BASIS_FOURIER, n_modes=24, n_quad=24)])
    assert create(d)[0] == -2
    d, _ = _fill_plan_desc(L.OP_FORWARD, L.F64, (4, 8), [None, AxisSpec(L.BASIS_TABLE, table=tab)])
    d.slab_size = 2
    assert create(d)[0] == -2
    d, _ = _fill_plan_desc(L.OP_NONLINEAR, L.F64, (4, 8), [None, AxisSpec(L.BASIS_TABLE, table=tab)])
    assert create(d)[0] == -1
    d, _ = _fill_plan_desc(L.OP_FORWARD, 7, (4, 8), [None, AxisSpec(L.BASIS_TABLE, table=tab)])
    assert create(d)[0] == -1
    del keep


def test_execute_rejects_null_and_missing_workspace(cuda):
    lib = L.load()
    T = jf.TensorProduct(jf.Legendre(8), jf.Legendre(8), jf.Legendre(8))
    x = torch.randn(8, 8, 8, dtype=torch.float64, device=cuda)
    plan = T._plan(L.OP_BACKWARD, x)
    assert plan.workspace_bytes > 0
    out = torch.empty_like(x)
    assert lib.jfx_execute(plan._h, None, None, C.c_void_p(out.data_ptr()), None) == -1
    rc = lib.jfx_execute(plan._h, None, C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), None)
    assert rc == -1 and b"workspace" in lib.jfx_last_error()


def test_python_layer_raises_like_the_reference(cuda):
    V = jf.Legendre(8)
    with pytest.raises(AssertionError):   # backward only pads (orthogonal.py:216-224)
        V.backward(torch.zeros(8, dtype=torch.float64, device=cuda), N=4)
    with pytest.raises(AssertionError):   # more coefficients than modes
        V.backward(torch.zeros(9, dtype=torch.float64, device=cuda))
    plan = Plan(L.OP_FORWARD, L.F64, (8,), [V.axis_spec(L.OP_FORWARD, 8, L.F64)])
    with pytest.raises(ValueError):
        plan.execute(torch.zeros(9, dtype=torch.float64, device=cuda))
    with pytest.raises(TypeError):
        plan.execute(torch.zeros(8, dtype=torch.float32, device=cuda))
    with pytest.raises(L.JfxError):       # device path needs device memory: no silent host fallback
        plan.execute(torch.zeros(8, dtype=torch.float64))


@pytest.mark.parametrize("cls", ["Legendre", "Chebyshev", "Fourier"])
def test_empty_batch_and_single_line(cuda, cls):
    V = getattr(jf, cls)(16)
    dt = torch.complex128 if cls == "Fourier" else torch.float64
    e = torch.zeros(0, 16, dtype=dt, device=cuda)
    assert tuple(V.backward(e).shape) == (0, 16) and tuple(V.forward(e).shape) == (0, 16)
    one = torch.randn(1, 16, dtype=torch.float64, device=cuda).to(dt)
    back = V.forward(V.backward(one))
    assert float((back - one).abs().max()) < 1e-13
    # strided axis with a single line on each side
    x = torch.randn(1, 16, 1, dtype=torch.float64, device=cuda).to(dt)
    if cls != "Chebyshev":   # real Chebyshev data with an odd inner extent uses the dense table (still valid)
        y = V.forward(V.backward(x, axis=1), axis=1)
        assert float((y - x).abs().max()) < 1e-13


def test_plan_registry_keys_are_content_hashes_and_plans_are_per_device(cuda):
    """jfx_registry_register / jfx_registry_acquire (include/jfx.h): the key depends on the descriptor's CONTENTS (table values,
    not pointers), the same key returns the same plan, an unknown key fails loudly, and a registered plan computes what a
    directly created plan computes."""
    import ctypes as C
    import numpy as np
    import torch
    import jaxfun_b200 as jf
    from jaxfun_b200 import _lib as L
    from jaxfun_b200.engine import _fill_plan_desc, AxisSpec, Plan
    lib = L.load()
    V = jf.Legendre(48)
    T = np.ascontiguousarray(V._dense_table(L.OP_BACKWARD, 48, 48, 0))
    shape = (7, 48)
    keys = []
    for table in (T, T.copy(), T * (1 + 1e-15)):                      # same contents at another address; different contents
        desc, keep = _fill_plan_desc(L.OP_APPLY, L.F64, shape, [None, AxisSpec(L.BASIS_TABLE, table=table)])
        k = C.c_uint64()
        L.check(lib.jfx_registry_register(C.byref(desc), C.byref(k)))
        keys.append(int(k.value))
    assert keys[0] == keys[1] and keys[0] != keys[2]
    h1, h2 = C.c_void_p(), C.c_void_p()
    L.check(lib.jfx_registry_acquire(C.c_uint64(keys[0]), C.byref(h1)))
    L.check(lib.jfx_registry_acquire(C.c_uint64(keys[0]), C.byref(h2)))
    assert h1.value == h2.value and h1.value
    x = torch.randn(shape, dtype=torch.float64, device=cuda)
    out = torch.empty_like(x)
    L.check(lib.jfx_execute(h1, None, C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), None))
    torch.cuda.synchronize()
    ref = Plan(L.OP_APPLY, L.F64, shape, [None, AxisSpec(L.BASIS_TABLE, table=T)])(x)
    assert torch.equal(out, ref)
    bad = C.c_void_p()
    assert lib.jfx_registry_acquire(C.c_uint64(12345), C.byref(bad)) == -1 and b"not registered" in lib.jfx_last_error()
