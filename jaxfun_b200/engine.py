"""Python face of the jfx engine: plans over device (torch.cuda) or host (numpy) arrays.

PyTorch is used only as plumbing here — device allocations, the current CUDA stream and
`torch.distributed`; every transform runs in libjfx.so through the C ABI (include/jfx.h).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Sequence

import numpy as np

from . import _lib as L

try:  # plumbing only
    import torch
except Exception:  # pragma: no cover - torch is part of the image
    torch = None


_NP2JFX = {np.dtype("float32"): L.F32, np.dtype("float64"): L.F64,
           np.dtype("complex64"): L.C64, np.dtype("complex128"): L.C128}
_JFX2NP = {v: k for k, v in _NP2JFX.items()}


def jfx_dtype(dt) -> int:
    if torch is not None and isinstance(dt, torch.dtype):
        dt = {torch.float32: "float32", torch.float64: "float64",
              torch.complex64: "complex64", torch.complex128: "complex128"}[dt]
    return _NP2JFX[np.dtype(dt)]


def _torch_dtype(code: int):
    return {L.F32: torch.float32, L.F64: torch.float64, L.C64: torch.complex64, L.C128: torch.complex128}[code]


def device_count() -> int:
    return int(L.load().jfx_device_count())


def require_device() -> None:
    if device_count() < 1:
        raise L.JfxError(-3, "no CUDA device visible: jaxfun_b200 has no CPU fallback")


def fast_path_available(basis: int, n: int, dtype: int) -> bool:
    return bool(L.load().jfx_fast_path_available(int(basis), int(n), int(dtype)))


@dataclass
class AxisSpec:
    """One transformed axis of a plan (mirrors jfx_axis_desc)."""
    basis: int
    n_modes: int = 0
    n_quad: int = 0
    deriv: int = 0
    domain_factor: float = 1.0
    table: np.ndarray | None = None  # [n_out, n_in] float64 / complex128, C-contiguous


def current_stream_ptr() -> int:
    return int(torch.cuda.current_stream().cuda_stream)


class device_scope:
    """Make the device of a CUDA array current for the duration of a block (plans, tables and workspaces live on ONE device:
    the one that is current when they are created; launches go to the current device's current stream).  A no-op for host
    arrays and when the device is current already."""

    def __init__(self, x):
        self._ctx = None
        dev = getattr(x, "device", None)
        if torch is not None and getattr(x, "is_cuda", False) and dev.index is not None \
                and dev.index != torch.cuda.current_device():
            self._ctx = torch.cuda.device(dev)

    def __enter__(self):
        if self._ctx is not None:
            self._ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self._ctx is not None:
            return self._ctx.__exit__(*exc)
        return False


def device_key(x):
    """Cache-key component: the CUDA device index of a device array, "host" for numpy arrays."""
    return x.device.index if getattr(x, "is_cuda", False) else "host"


class PinnedArray:
    """numpy array living in CUDA pinned host memory (jfx_host_alloc)."""

    def __init__(self, shape, dtype):
        self._lib = L.load()
        dtype = np.dtype(dtype)
        nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
        ptr = C.c_void_p()
        L.check(self._lib.jfx_host_alloc(C.byref(ptr), nbytes))
        self._ptr = ptr
        buf = (C.c_char * max(nbytes, 1)).from_address(ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape, dtype=np.int64))).reshape(shape)

    def __del__(self):
        try:
            if self._ptr:
                self._lib.jfx_host_free(self._ptr)
                self._ptr = None
        except Exception:
            pass


def _fill_plan_desc(op: int, dtype: int, shape_in: Sequence[int], axes: Sequence[AxisSpec | None]):
    """Build a PlanDesc; returns (desc, keepalive list of table arrays)."""
    ndim = len(shape_in)
    if not 1 <= ndim <= L.JFX_MAX_DIMS:
        raise ValueError(f"arrays of rank {ndim} are not supported (1..{L.JFX_MAX_DIMS})")
    if len(axes) != ndim:
        raise ValueError("one AxisSpec (or None) per array axis is required")
    d = L.PlanDesc()
    d.abi_version = L.JFX_ABI_VERSION
    d.op, d.dtype, d.ndim = int(op), int(dtype), ndim
    keep = []
    for i, (n, a) in enumerate(zip(shape_in, axes)):
        d.shape_in[i] = int(n)
        ad = d.axis[i]
        if a is None or a.basis == L.BASIS_NONE:
            ad.basis = L.BASIS_NONE
            continue
        ad.basis = int(a.basis)
        ad.n_modes, ad.n_quad, ad.deriv = int(a.n_modes), int(a.n_quad), int(a.deriv)
        ad.domain_factor = float(a.domain_factor)
        if a.table is not None:
            want = np.complex128 if a.basis == L.BASIS_CTABLE else np.float64
            t = np.ascontiguousarray(a.table, dtype=want)
            keep.append(t)
            ad.table = t.ctypes.data
            ad.table_rows, ad.table_cols = t.shape
    d.slab_rank, d.slab_size = 0, 1
    return d, keep


class Plan:
    """An immutable transform plan (jfx_plan).  Callable on torch.cuda tensors or numpy arrays."""

    def __init__(self, op: int, dtype: int, shape_in: Sequence[int], axes: Sequence[AxisSpec | None]):
        self._lib = L.load()
        require_device()
        # the plan's tables are uploaded to the device that is current now; execute() refuses arrays of another device
        self.device_index = int(torch.cuda.current_device()) if torch is not None and torch.cuda.is_available() else 0
        self.op, self.dtype = int(op), int(dtype)
        self.shape_in = tuple(int(s) for s in shape_in)
        desc, keep = _fill_plan_desc(op, dtype, shape_in, axes)
        handle = C.c_void_p()
        L.check(self._lib.jfx_plan_create(C.byref(desc), C.byref(handle)))
        del keep
        self._h = handle
        so = (C.c_int64 * L.JFX_MAX_DIMS)()
        L.check(self._lib.jfx_plan_shape_out(self._h, so))
        self.shape_out = tuple(int(so[i]) for i in range(len(shape_in)))
        ws = C.c_size_t()
        L.check(self._lib.jfx_plan_workspace_bytes(self._h, C.byref(ws)))
        self.workspace_bytes = int(ws.value)
        fl, by = C.c_double(), C.c_double()
        L.check(self._lib.jfx_plan_work(self._h, C.byref(fl), C.byref(by)))
        self.flops, self.bytes = float(fl.value), float(by.value)
        L.check(self._lib.jfx_plan_executed_flops(self._h, C.byref(fl)))
        self.flops_executed = float(fl.value)   # parity-folded table passes issue half the algorithmic flops
        self.launches = int(self._lib.jfx_plan_launches(self._h))
        self._ws = {}

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.jfx_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # -- device path -----------------------------------------------------------------------
    def workspace(self, device):
        key = (device.index, int(torch.cuda.current_stream(device).cuda_stream))
        ws = self._ws.get(key)
        if ws is None and self.workspace_bytes:
            ws = torch.empty(self.workspace_bytes, dtype=torch.uint8, device=device)
            self._ws[key] = ws
        return ws

    def execute(self, x, out=None):
        if tuple(x.shape) != self.shape_in:
            raise ValueError(f"plan expects shape {self.shape_in}, got {tuple(x.shape)}")
        if jfx_dtype(x.dtype) != self.dtype:
            raise TypeError(f"plan dtype {_JFX2NP[self.dtype]} != array dtype {x.dtype}")
        if not x.is_cuda:
            raise L.JfxError(-3, "device path needs a CUDA tensor; jaxfun_b200 has no CPU fallback")
        if x.device.index != self.device_index:
            raise ValueError(f"this plan lives on cuda:{self.device_index}, the array on {x.device}: plans are per device "
                             "(create them under engine.device_scope(array))")
        with device_scope(x):
            return self._execute_on_device(x, out)

    def _execute_on_device(self, x, out):
        x = x.contiguous()
        if out is None:
            out = torch.empty(self.shape_out, dtype=x.dtype, device=x.device)
        elif (tuple(out.shape) != self.shape_out or out.dtype != x.dtype or out.device != x.device
              or not out.is_contiguous()):
            # the kernels write shape_out elements of the plan's dtype through a raw pointer: anything else would be silent
            # memory corruption
            raise ValueError(f"out must be a contiguous {x.dtype} array of shape {self.shape_out} on {x.device}")
        elif out.data_ptr() == x.data_ptr():
            raise ValueError("a transform cannot run in place: out must not alias the input")
        ws = self.workspace(x.device)
        L.check(self._lib.jfx_execute(self._h, C.c_void_p(current_stream_ptr()), C.c_void_p(x.data_ptr()),
                                      C.c_void_p(out.data_ptr()), C.c_void_p(ws.data_ptr() if ws is not None else 0)))
        return out

    # -- slab exchange fused into the last pass (peer stores over NVLink) --------------------
    def scatter_supported(self, parts: int, split_axis: int) -> bool:
        return bool(self._lib.jfx_plan_scatter_supported(self._h, int(parts), int(split_axis)))

    def execute_scatter(self, x, peer_ptrs: Sequence[int], rank: int, split_axis: int):
        """Run the plan; its final pass stores the result into the receive buffers `peer_ptrs[p]` (device pointers valid
        on this device, one per rank) in the layout the tiled all-to-all of sharding.py:83-89 would leave there."""
        if tuple(x.shape) != self.shape_in:
            raise ValueError(f"plan expects shape {self.shape_in}, got {tuple(x.shape)}")
        if not x.is_cuda:
            raise L.JfxError(-3, "device path needs a CUDA tensor; jaxfun_b200 has no CPU fallback")
        if x.device.index != self.device_index:
            raise ValueError(f"this plan lives on cuda:{self.device_index}, the array on {x.device}")
        x = x.contiguous()
        ws = self.workspace(x.device)
        ptrs = (C.c_void_p * len(peer_ptrs))(*[C.c_void_p(int(p)) for p in peer_ptrs])
        L.check(self._lib.jfx_execute_scatter(self._h, C.c_void_p(current_stream_ptr()), C.c_void_p(x.data_ptr()), ptrs,
                                              len(peer_ptrs), int(rank), int(split_axis),
                                              C.c_void_p(ws.data_ptr() if ws is not None else 0)))

    # -- host path (e2e: H2D + transform + D2H inside the call) ---------------------------
    def execute_host(self, x: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        if tuple(x.shape) != self.shape_in:
            raise ValueError(f"plan expects shape {self.shape_in}, got {tuple(x.shape)}")
        x = np.ascontiguousarray(x, dtype=_JFX2NP[self.dtype])
        if out is None:
            out = np.empty(self.shape_out, dtype=x.dtype)
        assert out.flags.c_contiguous and out.dtype == x.dtype and tuple(out.shape) == self.shape_out
        L.check(self._lib.jfx_execute_host(self._h, C.c_void_p(0), C.c_void_p(x.ctypes.data),
                                           C.c_void_p(out.ctypes.data)))
        return out

    def __call__(self, x, out=None):
        if isinstance(x, np.ndarray):
            return self.execute_host(x, out)
        return self.execute(x, out)


def as_jfx_array(x, complex_required: bool = False):
    """Return (array, is_host).  numpy -> host path; torch tensor -> device path.

    With `complex_required` real input is promoted to the matching complex dtype (what
    `jnp.fft.fft` does implicitly in the reference, galerkin/Fourier.py:160,177)."""
    if isinstance(x, np.ndarray):
        if x.dtype not in _NP2JFX:
            x = x.astype(np.complex128 if np.iscomplexobj(x) else np.float64)
        if complex_required and not np.iscomplexobj(x):
            x = x.astype(np.complex128 if x.dtype == np.float64 else np.complex64)
        return x, True
    if torch is not None and isinstance(x, torch.Tensor):
        if complex_required and not x.is_complex():
            x = x.to(torch.complex128 if x.dtype == torch.float64 else torch.complex64)
        return x, False
    return as_jfx_array(np.asarray(x), complex_required)
