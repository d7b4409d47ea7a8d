"""The C-ABI library loads on a CPU-only host and exports every symbol include/jfx.h declares;
compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "jfx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(jfx_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_contract_entry_points():
    syms = header_symbols()
    for must in ("jfx_plan_create", "jfx_plan_destroy", "jfx_plan_workspace_bytes", "jfx_execute",
                 "jfx_last_error", "jfx_execute_host", "jfx_nonlinear_create", "jfx_nonlinear_execute",
                 "jfx_slab_pack", "jfx_slab_unpack", "jfx_slab_create", "jfx_slab_bind", "jfx_slab_execute",
                 "jfx_slab_sizes", "jfx_slab_destroy", "jfx_execute_scatter", "jfx_registry_register"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from jaxfun_b200 import _lib
    lib = _lib.load()
    for name in header_symbols():
        assert hasattr(lib, name), f"libjfx.so does not export {name}"
    assert sorted(_lib.exported_symbols()) == header_symbols()
    assert lib.jfx_abi_version() == _lib.JFX_ABI_VERSION


def test_struct_layout_matches_header():
    """sizeof checks against the C compiler's view of the header."""
    import subprocess
    import tempfile
    from jaxfun_b200 import _lib
    code = ('#include "jfx.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu\\n", sizeof(jfx_axis_desc),'
            'sizeof(jfx_plan_desc), sizeof(jfx_nonlinear_desc), sizeof(jfx_pw_instr));'
            'printf("%zu\\n", sizeof(jfx_banded_desc));return 0;}\n')
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "s.c")
        open(src, "w").write(code)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
        sizes = [int(v) for v in subprocess.check_output([exe]).split()]
    assert sizes == [C.sizeof(_lib.AxisDesc), C.sizeof(_lib.PlanDesc), C.sizeof(_lib.NonlinearDesc),
                     C.sizeof(_lib.PwInstr), C.sizeof(_lib.BandedDesc)]


def test_no_cpu_fallback():
    """Without a device, plan creation must fail with JFX_ERR_CUDA (-3), not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    import jaxfun_b200 as jf
    from jaxfun_b200 import _lib
    assert jf.device_count() == 0
    with pytest.raises(_lib.JfxError) as e:
        jf.Legendre(8).forward(np.zeros(8))
    assert e.value.code == -3


def test_slab_object_argument_checks_and_no_cpu_fallback():
    """jfx_slab_create validates its arguments before touching a device and, without one, fails with JFX_ERR_CUDA like a
    plan (the slab transform has no host route either)."""
    import torch
    from jaxfun_b200 import _lib
    lib = _lib.load()
    d = _lib.PlanDesc()
    d.abi_version, d.op, d.dtype, d.ndim = _lib.JFX_ABI_VERSION, _lib.OP_BACKWARD, _lib.F64, 3
    for i in range(3):
        d.shape_in[i] = 8
    h = C.c_void_p()
    d.slab_rank, d.slab_size = 2, 2                                   # rank outside 0 .. size - 1
    assert lib.jfx_slab_create(C.byref(d), _lib.SLAB_SPECTRAL, C.byref(h)) == -1 and not h.value
    assert b"slab rank" in lib.jfx_last_error()
    d.slab_rank = 0
    assert lib.jfx_slab_create(C.byref(d), 7, C.byref(h)) == -1          # unknown sharding
    d.ndim = 1
    assert lib.jfx_slab_create(C.byref(d), _lib.SLAB_SPECTRAL, C.byref(h)) == -1   # a slab needs two axes to exchange
    d.ndim = 3
    assert lib.jfx_slab_execute(None, None, None, None, None) == -1      # null arguments are refused, not dereferenced
    if not torch.cuda.is_available():
        assert lib.jfx_slab_create(C.byref(d), _lib.SLAB_SPECTRAL, C.byref(h)) == -3 and not h.value


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "jaxfun_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "jaxfun_oracle" not in txt and "import oracle" not in txt, f
