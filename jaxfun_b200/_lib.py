"""ctypes binding of libjfx.so (the C ABI declared in include/jfx.h).

Loading never falls back to anything else: if the shared object is missing the import raises, and
every compute call fails loudly (JfxError) when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# JFX_LIB_PATH: A/B experiments load a differently-built copy of the same library (tools/build_variant.py)
LIB_PATH = os.environ.get("JFX_LIB_PATH") or os.path.join(_HERE, "libjfx.so")

JFX_ABI_VERSION = 1
JFX_MAX_DIMS = 4
JFX_MAX_LEAVES = 8
JFX_MAX_PROGRAM = 128

# dtypes
F32, F64, C64, C128 = 0, 1, 2, 3
# ops
OP_FORWARD, OP_SCALAR_PRODUCT, OP_BACKWARD, OP_BACKWARD_PRIMITIVE, OP_NONLINEAR, OP_APPLY = range(6)
# slab shardings (of the INPUT of a slab transform)
SLAB_SPECTRAL, SLAB_PHYSICAL = 0, 1
# bases
BASIS_NONE, BASIS_TABLE, BASIS_CTABLE, BASIS_CHEBYSHEV, BASIS_FOURIER = range(5)
# pointwise opcodes
(PW_LEAF, PW_CONST, PW_ADD, PW_MUL, PW_POWI, PW_ABS, PW_NEG, PW_FUNC, PW_POWR, PW_CONJ,
 PW_STATIC) = range(11)
FN = {name: i for i, name in enumerate(
    ["exp", "log", "sin", "cos", "tan", "sinh", "cosh", "tanh", "sqrt", "sign", "Heaviside",
     "asin", "acos", "atan", "asinh", "acosh", "atanh", "re", "im"])}


class AxisDesc(C.Structure):
    _fields_ = [
        ("basis", C.c_int32), ("n_modes", C.c_int32), ("n_quad", C.c_int32), ("deriv", C.c_int32),
        ("domain_factor", C.c_double), ("table", C.c_void_p),
        ("table_rows", C.c_int32), ("table_cols", C.c_int32),
    ]


class PlanDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("op", C.c_int32), ("dtype", C.c_int32), ("ndim", C.c_int32),
        ("shape_in", C.c_int64 * JFX_MAX_DIMS), ("axis", AxisDesc * JFX_MAX_DIMS),
        ("slab_rank", C.c_int32), ("slab_size", C.c_int32), ("reserved", C.c_int32 * 14),
    ]


class PwInstr(C.Structure):
    _fields_ = [("op", C.c_int32), ("arg", C.c_int32)]


class NonlinearDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("n_leaves", C.c_int32),
        ("leaves", C.POINTER(PlanDesc) * JFX_MAX_LEAVES), ("final_transform", C.POINTER(PlanDesc)),
        ("n_program", C.c_int32), ("program", PwInstr * JFX_MAX_PROGRAM),
        ("n_consts", C.c_int32), ("consts", (C.c_double * 2) * 32),
        ("n_statics", C.c_int32), ("statics", C.c_void_p * JFX_MAX_LEAVES),
        ("reserved", C.c_int32 * 8),
    ]


JFX_BANDED_MAX_TERMS = 8


class BandedDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("dtype", C.c_int32), ("band_complex", C.c_int32), ("n_terms", C.c_int32),
        ("n", C.c_int64), ("n_sys", C.c_int64), ("n_diags", C.c_int32), ("reserved0", C.c_int32),
        ("offsets", C.POINTER(C.c_int32)), ("weights", C.c_void_p), ("diags", C.c_void_p),
        ("reserved", C.c_int32 * 8),
    ]


class JfxError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"jfx error {code}: {message}")
        self.code = code


# every symbol include/jfx.h declares, with its ctypes signature
_SIGNATURES = {
    "jfx_abi_version": (C.c_int, []),
    "jfx_last_error": (C.c_char_p, []),
    "jfx_device_count": (C.c_int, []),
    "jfx_fast_path_available": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "jfx_plan_create": (C.c_int, [C.POINTER(PlanDesc), C.POINTER(C.c_void_p)]),
    "jfx_plan_destroy": (None, [C.c_void_p]),
    "jfx_registry_register": (C.c_int, [C.POINTER(PlanDesc), C.POINTER(C.c_uint64)]),
    "jfx_registry_acquire": (C.c_int, [C.c_uint64, C.POINTER(C.c_void_p)]),
    "jfx_registry_clear": (None, []),
    "jfx_plan_ndim": (C.c_int, [C.c_void_p]),
    "jfx_plan_shape_out": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "jfx_plan_workspace_bytes": (C.c_int, [C.c_void_p, C.POINTER(C.c_size_t)]),
    "jfx_plan_work": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "jfx_plan_executed_flops": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "jfx_plan_launches": (C.c_int, [C.c_void_p]),
    "jfx_plan_scatter_supported": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "jfx_execute_scatter": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int,
                                      C.c_void_p]),
    "jfx_execute": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "jfx_execute_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "jfx_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "jfx_host_free": (C.c_int, [C.c_void_p]),
    "jfx_nonlinear_create": (C.c_int, [C.POINTER(NonlinearDesc), C.POINTER(C.c_void_p)]),
    "jfx_nonlinear_destroy": (None, [C.c_void_p]),
    "jfx_nonlinear_workspace_bytes": (C.c_int, [C.c_void_p, C.POINTER(C.c_size_t)]),
    "jfx_nonlinear_shape_out": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int)]),
    "jfx_nonlinear_launches": (C.c_int, [C.c_void_p]),
    "jfx_nonlinear_execute": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "jfx_pointwise": (C.c_int, [C.c_void_p, C.POINTER(PwInstr), C.c_int, C.c_void_p, C.c_int,
                                C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p), C.c_void_p,
                                C.c_int64, C.c_int]),
    "jfx_slab_pack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int,
                                C.c_int, C.c_int, C.c_int]),
    "jfx_slab_unpack": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int,
                                  C.c_int, C.c_int, C.c_int]),
    "jfx_slab_create": (C.c_int, [C.POINTER(PlanDesc), C.c_int, C.POINTER(C.c_void_p)]),
    "jfx_slab_destroy": (None, [C.c_void_p]),
    "jfx_slab_sizes": (C.c_int, [C.c_void_p, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                 C.POINTER(C.c_int64)]),
    "jfx_slab_fused": (C.c_int, [C.c_void_p]),
    "jfx_slab_bind": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "jfx_slab_execute": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "jfx_axpby_diag": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_double),
                                 C.POINTER(C.c_void_p), C.c_void_p, C.c_int64, C.c_int, C.c_int]),
    "jfx_point_contract": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_int,
                                   C.c_int]),
    "jfx_banded_create": (C.c_int, [C.POINTER(BandedDesc), C.POINTER(C.c_void_p)]),
    "jfx_banded_destroy": (None, [C.c_void_p]),
    "jfx_banded_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_size_t)]),
    "jfx_banded_factors": (C.c_int, [C.c_void_p, C.c_void_p]),
    "jfx_banded_solve": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]),
    "jfx_calibrate_dmma": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double)]),
    "jfx_calibrate_dfma": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double)]),
}

_lib = None


def load() -> C.CDLL:
    """Load libjfx.so (once).  Raises if it has not been built — there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python jaxfun_b200/_build.py` "
            "(or __graft_entry__.build()). jaxfun_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library drift
        fn.restype = res
        fn.argtypes = args
    if lib.jfx_abi_version() != JFX_ABI_VERSION:
        raise ImportError("libjfx.so ABI version mismatch; rebuild")
    _lib = lib
    return lib


def check(code: int) -> None:
    if code != 0:
        raise JfxError(code, load().jfx_last_error().decode())


def exported_symbols() -> list[str]:
    return sorted(_SIGNATURES)
