"""Slab-decomposed tensor-product transforms over the GPUs of one box.

Mirror of `jaxfun.sharding` (`src/jaxfun/sharding.py:9-105`): spectral arrays are sharded along
axis 0 (`spectral_sharding = P("k")`), physical arrays along axis 1 (`physical_sharding =
P(None, "k")`); one transform is

    phase 1: the unsharded axes, locally
    exchange: lax.all_to_all(split_axis=unsharded[0], concat_axis=sharded[0], tiled=True)
    phase 2: the originally sharded axis, locally

One process per GPU; the exchange is `torch.distributed.all_to_all_single` (NCCL over
NVLink/NVSwitch on the GPU box, gloo in the CPU tests).  The block layout is chosen so that only ONE
side of each exchange needs a repack kernel:

* spectral -> physical (`backward`): split axis 1 -> `jfx_slab_pack` into [P, n0/P, n1/P, ...];
  the received blocks concatenate along axis 0, i.e. they already ARE the contiguous result.
* physical -> spectral (`forward`, `scalar_product`): split axis 0 -> the send blocks are already
  contiguous; the received blocks interleave along axis 1 -> `jfx_slab_unpack`.

The local phases are ordinary engine plans on the local block; `SlabBackend` abstracts them so the
orchestration can be exercised on CPU ranks (gloo) with the oracle's 1-D functions injected.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Sequence

import numpy as np

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None

SPECTRAL = "spectral"   # axis 0 sharded  (sharding.py:10)
PHYSICAL = "physical"   # axis 1 sharded  (sharding.py:11)


def get_transposed_sharding(sharding: str) -> str:
    """sharding.py:14-21."""
    if sharding == SPECTRAL:
        return PHYSICAL
    if sharding == PHYSICAL:
        return SPECTRAL
    raise ValueError(f"Provided {sharding} does not match spectral or physical.")


def sharded_axis(sharding: str) -> int:
    return 0 if sharding == SPECTRAL else 1


class SlabBackend:
    """What the slab algorithm needs from its local engine."""

    def apply_axes(self, x, axes: Sequence[int]):
        """Apply the per-axis transforms for `axes` (in that order) to the local block."""
        raise NotImplementedError

    def pack(self, x, split_axis: int, parts: int):
        """[.., L, ..] -> [parts, .., L/parts, ..] contiguous."""
        raise NotImplementedError

    def unpack(self, blocks, concat_axis: int, parts: int):
        """[parts, .., L/parts, ..] -> [.., L, ..] contiguous."""
        raise NotImplementedError

    #: process group of the exchange (None = the default group); set by SlabTensorProduct
    group = None

    def all_to_all(self, send):
        """Tiled all-to-all of the leading `parts` dimension; same shape out."""
        out = torch.empty_like(send)
        dist.all_to_all_single(out, send, group=self.group)
        return out

    def all_to_all_async(self, send):
        """Same exchange, not waited for: returns (out, work).  `work.wait()` orders the current stream after the
        exchange (NCCL runs it on its own stream, so kernels issued in between overlap it); `send` must stay
        referenced until then."""
        out = torch.empty_like(send)
        work = dist.all_to_all_single(out, send, async_op=True, group=self.group)
        return out, work


def slab_p2p() -> bool:
    """The last local pass of phase 1 stores straight into the peers' receive buffers (symmetric memory over NVLink)
    instead of pack + NCCL all-to-all (+ unpack).  Default since it ran on real GPUs (2 x B200, 512^3: 7.3 ms per
    backward + forward against 9.6 ms with the NCCL exchange, results identical); JFX_SLAB_P2P=0 keeps the NCCL path, which
    is also what runs when the plan has no scatter epilogue or symmetric memory cannot be set up."""
    import os
    return os.environ.get("JFX_SLAB_P2P", "1") != "0"


def slab_native() -> bool:
    """The whole slab transform is ONE C-ABI call (`jfx_slab_execute`: phase 1 with peer stores or strided peer copies,
    device-side barrier on peer-mapped flags, phase 2) — no NCCL collective, no torch op on the data path; torch only
    provides the symmetric-memory allocation the peers map.  Default; JFX_SLAB_NATIVE=0 keeps the Python orchestration."""
    import os
    return os.environ.get("JFX_SLAB_NATIVE", "1") != "0"


def slab_fused_pack() -> bool:
    """JFX_SLAB_FUSED_PACK=1: spectral -> physical keeps the NCCL all-to-all, but the last local pass writes the packed send
    buffer itself (the scatter epilogue of jfx_execute_scatter aimed at this rank's own buffer) — no jfx_slab_pack launch."""
    import os
    return os.environ.get("JFX_SLAB_FUSED_PACK", "0") == "1"


def slab_chunks() -> int:
    """Number of chunks of the overlapped exchange (JFX_SLAB_CHUNKS, default 1 = one blocking all-to-all)."""
    import os
    try:
        return max(1, int(os.environ.get("JFX_SLAB_CHUNKS", "1")))
    except ValueError:
        return 1


def _chunk_bounds(n: int, chunks: int):
    per = -(-n // max(1, min(chunks, n)))
    return [(a, min(n, a + per)) for a in range(0, n, per)]


def _apply_separable_slab_chunked(x, sharding: str, backend: SlabBackend, world_size: int, chunks: int):
    """Same result as apply_separable_slab, with the all-to-all cut into chunks that overlap the local passes.

    spectral -> physical (split axis 1, concat axis 0): the phase-1 passes act on axes >= 1, so the local planes of axis 0
    are independent: chunk c of planes is transformed, packed and sent while chunk c + 1 is being transformed; phase 2
    (axis 0) starts when every chunk has arrived.
    physical -> spectral (split axis 0, concat axis 1): phase 1 runs on the whole block; the rows each peer receives are
    sent in chunks, and chunk c is unpacked and taken through phase 2 (axis 1) while chunk c + 1 is in flight."""
    P = world_size
    sh = sharded_axis(sharding)
    unsharded = [ax for ax in range(x.ndim) if ax != sh]
    if sh == 0:
        L = x.shape[0]
        pending = []
        y_full = None
        for a, b in _chunk_bounds(L, chunks):
            y = backend.apply_axes(x[a:b].contiguous(), unsharded)
            if y.shape[1] % P != 0:
                raise ValueError(f"split axis 1 has extent {y.shape[1]}, not divisible by {P} devices")
            send = backend.pack(y, 1, P)                              # [P, b-a, m1/P, ...]
            recv, work = backend.all_to_all_async(send)
            if y_full is None:
                y_full = recv.new_empty((P, L) + tuple(recv.shape[2:]))
            pending.append((a, b, send, recv, work))
        for a, b, send, recv, work in pending:
            work.wait()
            y_full[:, a:b] = recv                                     # block p of chunk c = planes a..b of rank p
        y = y_full.reshape((P * L,) + tuple(y_full.shape[2:]))
        return backend.apply_axes(y, [0])
    y = backend.apply_axes(x, unsharded)
    if y.shape[0] % P != 0:
        raise ValueError(f"split axis 0 has extent {y.shape[0]}, not divisible by {P} devices")
    rows = y.shape[0] // P
    blocks = y.reshape((P, rows) + tuple(y.shape[1:]))
    bounds = _chunk_bounds(rows, chunks)
    pending = []
    for a, b in bounds:
        send = blocks[:, a:b].contiguous()                            # [P, b-a, n1/P, ...]
        recv, work = backend.all_to_all_async(send)
        pending.append((send, recv, work))
    out = None
    for (a, b), (send, recv, work) in zip(bounds, pending):
        work.wait()
        z = backend.apply_axes(backend.unpack(recv, 1, P), [1])       # [b-a, n1, ...] -> phase 2 along axis 1
        if out is None:
            out = z.new_empty((rows,) + tuple(z.shape[1:]))
        out[a:b] = z
    return out


def apply_separable_slab(x, sharding: str, backend: SlabBackend, world_size: int, chunks: int | None = None):
    """`_apply_separable_spmd_shard_map` (sharding.py:43-105) for one rank's local block `x`.

    Returns the local block of the result, which carries the transposed sharding.  chunks > 1 (default: JFX_SLAB_CHUNKS)
    selects the overlapped exchange."""
    chunks = slab_chunks() if chunks is None else chunks
    if world_size > 1 and chunks == 1 and slab_native() and not slab_fused_pack() and hasattr(backend, "native_transform"):
        y = backend.native_transform(x, sharding, world_size)
        if y is not None:                      # None: symmetric memory unavailable -> host-composed routes below
            return y
    if world_size > 1 and slab_p2p() and hasattr(backend, "scatter_exchange"):
        y = backend.scatter_exchange(x, sharding, world_size)
        if y is not None:                      # None: this plan / shape has no fused path -> ordinary exchange below
            return backend.apply_axes(y, [sharded_axis(sharding)])
    if chunks > 1 and world_size > 1:
        return _apply_separable_slab_chunked(x, sharding, backend, world_size, chunks)
    if world_size > 1 and sharding == SPECTRAL and slab_fused_pack() and hasattr(backend, "packed_phase1"):
        send = backend.packed_phase1(x, world_size)
        if send is not None:
            recv = backend.all_to_all(send)
            y = recv.reshape((recv.shape[0] * recv.shape[1],) + tuple(recv.shape[2:]))
            return backend.apply_axes(y, [0])
    dim = x.ndim
    sh = sharded_axis(sharding)
    unsharded = [ax for ax in range(dim) if ax != sh]
    split_axis, concat_axis = unsharded[0], sh
    # Phase 1 — unsharded axes: fully local
    y = backend.apply_axes(x, unsharded)
    if y.shape[split_axis] % world_size != 0:
        raise ValueError(  # sharding.py:59-63
            f"split axis {split_axis} has extent {y.shape[split_axis]}, not divisible by {world_size} devices")
    # Exchange
    if world_size > 1:
        if split_axis == 0:
            send = y.reshape((world_size, y.shape[0] // world_size) + tuple(y.shape[1:]))
        else:
            send = backend.pack(y, split_axis, world_size)
        recv = backend.all_to_all(send)
        if concat_axis == 0:
            y = recv.reshape((recv.shape[0] * recv.shape[1],) + tuple(recv.shape[2:]))
        else:
            y = backend.unpack(recv, concat_axis, world_size)
    # Phase 2 — the originally sharded axis
    return backend.apply_axes(y, [sh])


# ---------------------------------------------------------------------------------------------------
# engine-backed implementation
# ---------------------------------------------------------------------------------------------------
class EngineSlabBackend(SlabBackend):
    """Local phases = engine plans restricted to a subset of axes; repacks = jfx_slab_pack/unpack."""

    def __init__(self, space, op: int, N=None, k=None):
        self.space, self.op, self.N, self.k = space, op, N, k
        self._plans = {}

    def _plan_for(self, x, axes):
        from .engine import Plan, jfx_dtype
        dtype = jfx_dtype(x.dtype)
        key = (tuple(x.shape), dtype, tuple(axes))
        plan = self._plans.get(key)
        if plan is None:
            specs = [None] * x.ndim
            for ax in axes:
                sp = self.space.basespaces[ax]
                specs[ax] = sp.axis_spec(self.op, x.shape[ax], dtype, None if self.N is None else self.N[ax],
                                         0 if self.k is None else self.k[ax],
                                         inner=int(np.prod(x.shape[ax + 1:], dtype=np.int64)))
            plan = self._plans[key] = Plan(self.op, dtype, tuple(x.shape), specs)
        return plan

    def apply_axes(self, x, axes):
        return self._plan_for(x, axes)(x)

    # ---- the whole transform as one jfx_slab call ---------------------------------------------------
    def native_transform(self, x, sharding: str, world_size: int):
        """jfx_slab_create / bind / execute (include/jfx.h): returns the local block of the result, or None when the peers'
        buffers cannot be mapped (no symmetric memory) — the caller then uses the host-composed exchange."""
        key = ("native", tuple(x.shape), x.dtype, sharding)
        ent = self._plans.get(key)
        if ent is False:
            return None
        if ent is None:
            from ._lib import JfxError
            try:
                slab = NativeSlab(self.space, self.op, tuple(x.shape), x.dtype, sharding, dist.get_rank(self.group), world_size,
                                  self.N, self.k)
            except JfxError as e:
                if "not divisible" in str(e):
                    raise ValueError(str(e)) from None      # the reference's error for this case (sharding.py:59-63)
                raise
            try:
                import torch.distributed._symmetric_memory as symm_mem
                grp = (self.group or dist.group.WORLD).group_name
                if hasattr(symm_mem, "enable_symm_mem_for_group"):
                    try:
                        symm_mem.enable_symm_mem_for_group(grp)
                    except Exception:
                        pass
                bufs, hdls = [], []
                for nbytes in (slab.recv_bytes, slab.recv_bytes, max(slab.signal_bytes, 256)):
                    b = symm_mem.empty((max(nbytes, 8) + 7) // 8, dtype=torch.int64, device=x.device)
                    b.zero_()
                    hdls.append(symm_mem.rendezvous(b, grp))
                    bufs.append(b)
                torch.cuda.synchronize(x.device)
                hdls[2].barrier(channel=0)          # every pad is zero before anybody raises a flag
                torch.cuda.synchronize(x.device)
            except Exception as e:
                self.p2p_error = f"{type(e).__name__}: {e}"
                slab.close()
                self._plans[key] = False
                return None
            slab.bind(*[[int(p) for p in hd.buffer_ptrs] for hd in hdls])
            ent = self._plans[key] = {"slab": slab, "bufs": bufs, "hdls": hdls, "fused": slab.fused}
        return ent["slab"](x)

    # ---- exchange fused into the last pass of phase 1 (peer stores) ---------------------------------
    def scatter_exchange(self, x, sharding: str, world_size: int):
        """Phase 1 + exchange in one go: returns the array phase 2 transforms, or None when the plan has no fused path.

        Receive buffers live in symmetric memory (torch.distributed._symmetric_memory): every rank maps every peer's
        buffer, the final contraction pass of phase 1 stores each output row into the rank that owns it after the
        exchange, and one device-side barrier orders the stores before phase 2 reads them.  Two buffers alternate per
        (shape, direction): a rank can run ahead by at most one transform before it meets the next barrier, so the
        buffer it overwrites has been read by everybody."""
        import torch.distributed._symmetric_memory as symm_mem
        if x.ndim != 3 or x.dtype != torch.float64:
            return None
        sh = sharded_axis(sharding)
        unsharded = [ax for ax in range(3) if ax != sh]
        split_axis = unsharded[0]
        plan = self._plan_for(x, unsharded)
        if not plan.scatter_supported(world_size, split_axis):
            return None
        s0, s1, s2 = plan.shape_out
        recv_shape = (world_size * s0, s1 // world_size, s2) if split_axis == 1 else (s0 // world_size, world_size * s1, s2)
        key = ("p2p", tuple(x.shape), split_axis)
        ent = self._plans.get(key)
        if ent is False:
            return None                     # symmetric memory could not be set up for this shape: NCCL path
        if ent is None:
            try:
                ent = self._p2p_setup(symm_mem, recv_shape, x)
            except Exception as e:          # e.g. a backend / driver without symmetric memory: keep the NCCL exchange
                self.p2p_error = f"{type(e).__name__}: {e}"
                self._plans[key] = False
                return None
            self._plans[key] = ent
        t = ent["turn"]
        ent["turn"] = t ^ 1
        hdl, buf = ent["hdls"][t], ent["bufs"][t]
        plan.execute_scatter(x, [int(p) for p in hdl.buffer_ptrs], dist.get_rank(self.group), split_axis)
        hdl.barrier(channel=t)
        return buf

    p2p_error = None

    def _p2p_setup(self, symm_mem, recv_shape, x):
        if True:
            bufs, hdls = [], []
            if hasattr(symm_mem, "enable_symm_mem_for_group"):      # needed by older torch releases, a no-op in newer ones
                try:
                    symm_mem.enable_symm_mem_for_group((self.group or dist.group.WORLD).group_name)
                except Exception:
                    pass
            for _ in range(2):
                b = symm_mem.empty(recv_shape, dtype=torch.float64, device=x.device)
                hdls.append(symm_mem.rendezvous(b, (self.group or dist.group.WORLD).group_name))
                bufs.append(b)
            return {"bufs": bufs, "hdls": hdls, "turn": 0}

    def packed_phase1(self, x, world_size: int, rank: int | None = None):
        """Phase 1 of spectral -> physical with the pack fused into its last pass: returns the send buffer
        [P, s0, s1/P, s2] (block p goes to rank p), or None when the plan has no scatter epilogue.  The epilogue computes
        peer[p] + ((rank*s0 + a)*(s1/P) + b')*s2; pointing peer[p] at (block p of the send buffer) - rank*s0*(s1/P)*s2
        elements turns that into block p, row (a, b') of THIS rank's buffer."""
        if x.ndim != 3 or x.dtype != torch.float64:
            return None
        plan = self._plan_for(x, [1, 2])
        if not plan.scatter_supported(world_size, 1):
            return None
        rank = dist.get_rank(self.group) if rank is None else rank
        s0, s1, s2 = plan.shape_out
        send = torch.empty((world_size, s0, s1 // world_size, s2), dtype=x.dtype, device=x.device)
        block = s0 * (s1 // world_size) * s2 * 8
        plan.execute_scatter(x, [send.data_ptr() + (p - rank) * block for p in range(world_size)], rank, 1)
        return send

    def _repack(self, fn_name: str, x, out_shape, full_shape, axis: int, parts: int):
        from . import _lib as L
        from .engine import current_stream_ptr, jfx_dtype
        lib = L.load()
        out = torch.empty(out_shape, dtype=x.dtype, device=x.device)
        shp = (C.c_int64 * len(full_shape))(*full_shape)
        L.check(getattr(lib, fn_name)(C.c_void_p(current_stream_ptr()), C.c_void_p(x.data_ptr()),
                                      C.c_void_p(out.data_ptr()), shp, len(full_shape), axis, parts,
                                      jfx_dtype(x.dtype)))
        return out

    def pack(self, x, split_axis, parts):
        x = x.contiguous()
        shp = list(x.shape)
        shp[split_axis] //= parts
        return self._repack("jfx_slab_pack", x, [parts] + shp, list(x.shape), split_axis, parts)

    def unpack(self, blocks, concat_axis, parts):
        blocks = blocks.contiguous()
        shp = list(blocks.shape[1:])
        shp[concat_axis] *= parts
        return self._repack("jfx_slab_unpack", blocks, shp, shp, concat_axis, parts)


class NativeSlab:
    """One rank's `jfx_slab` (include/jfx.h): the whole slab transform of a local block as one C-ABI call.

    The caller allocates two receive buffers of `recv_bytes` and a zeroed flag pad of `signal_bytes` per rank in memory
    every rank can map, and binds the pointers of all ranks (index = rank) once."""

    def __init__(self, space, op: int, local_shape, torch_dtype, sharding: str, rank: int, world_size: int, N=None, k=None):
        from . import _lib as L
        from .engine import _fill_plan_desc, jfx_dtype
        self._lib = lib = L.load()
        dtype = jfx_dtype(torch_dtype)
        sh = sharded_axis(sharding)
        shape = [int(v) for v in local_shape]
        specs = []
        for ax in range(len(shape)):
            sp = space.basespaces[ax]
            n_in = shape[ax] * world_size if ax == sh else shape[ax]   # the sharded axis is transformed at its global extent
            inner = int(np.prod(shape[ax + 1:], dtype=np.int64))
            if ax == sh == 0:
                inner //= world_size                                   # phase 2 sees the split axis at 1/P of its extent
            specs.append(sp.axis_spec(op, n_in, dtype, None if N is None else N[ax], 0 if k is None else k[ax], inner=inner))
        desc, keep = _fill_plan_desc(op, dtype, shape, specs)
        desc.slab_rank, desc.slab_size = int(rank), int(world_size)
        h = C.c_void_p()
        L.check(lib.jfx_slab_create(C.byref(desc), L.SLAB_SPECTRAL if sharding == SPECTRAL else L.SLAB_PHYSICAL, C.byref(h)))
        del keep
        self._h = h
        rb, sb, wb = C.c_size_t(), C.c_size_t(), C.c_size_t()
        so = (C.c_int64 * L.JFX_MAX_DIMS)()
        L.check(lib.jfx_slab_sizes(h, C.byref(rb), C.byref(sb), C.byref(wb), so))
        self.recv_bytes, self.signal_bytes, self.workspace_bytes = rb.value, sb.value, wb.value
        self.shape_out = tuple(int(v) for v in so[:len(shape)])
        self.fused = bool(lib.jfx_slab_fused(h))
        self.world_size, self.dtype, self._ws = int(world_size), torch_dtype, None

    def bind(self, recv0, recv1, signal):
        from . import _lib as L
        P = self.world_size
        arrs = [(C.c_void_p * P)(*[int(p) for p in ptrs]) for ptrs in (recv0, recv1, signal)]
        L.check(self._lib.jfx_slab_bind(self._h, *arrs))

    def __call__(self, x, out=None):
        from . import _lib as L
        from .engine import current_stream_ptr
        x = x.contiguous()
        if self._ws is None:
            self._ws = torch.empty(max(self.workspace_bytes, 8), dtype=torch.uint8, device=x.device)
        if out is None:
            out = torch.empty(self.shape_out, dtype=x.dtype, device=x.device)
        L.check(self._lib.jfx_slab_execute(self._h, C.c_void_p(current_stream_ptr()), C.c_void_p(x.data_ptr()),
                                           C.c_void_p(out.data_ptr()), C.c_void_p(self._ws.data_ptr())))
        return out

    def close(self):
        if self._h:
            self._lib.jfx_slab_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SlabTensorProduct:
    """Distributed face of a TensorProductSpace: same four transforms on local blocks.

    `backward` / `backward_primitive` take the spectral block (axis 0 sharded) and return the
    physical block (axis 1 sharded); `forward` / `scalar_product` the reverse — the sharding contract
    pinned by the reference's `tests/galerkin/test_forward_backward_spmd.py:71-75`."""

    def __init__(self, space, group=None):
        self.space = space
        self.group = group
        self._backends = {}

    @property
    def world_size(self) -> int:
        return dist.get_world_size(self.group) if dist is not None and dist.is_initialized() else 1

    def _backend(self, op, N=None, k=None):
        N = None if N is None else tuple(N)
        k = None if k is None else tuple(k)
        key = (op, N, k)
        b = self._backends.get(key)
        if b is None:
            b = self._backends[key] = EngineSlabBackend(self.space, op, N, k)
            b.group = self.group
        return b

    def backward(self, c_local, N=None):
        from . import _lib as L
        return apply_separable_slab(c_local, SPECTRAL, self._backend(L.OP_BACKWARD, N), self.world_size)

    def backward_primitive(self, c_local, k, N=None):
        from . import _lib as L
        return apply_separable_slab(c_local, SPECTRAL, self._backend(L.OP_BACKWARD_PRIMITIVE, N, tuple(k)),
                                    self.world_size)

    def forward(self, u_local):
        from . import _lib as L
        return apply_separable_slab(u_local, PHYSICAL, self._backend(L.OP_FORWARD), self.world_size)

    def scalar_product(self, u_local):
        from . import _lib as L
        return apply_separable_slab(u_local, PHYSICAL, self._backend(L.OP_SCALAR_PRODUCT), self.world_size)


def local_block(x, sharding: str, rank: int, world_size: int):
    """The block of a global array owned by `rank` under `sharding` (numpy or torch)."""
    ax = sharded_axis(sharding)
    n = x.shape[ax]
    if n % world_size:
        raise ValueError(f"axis {ax} of extent {n} is not divisible by {world_size} devices")
    b = n // world_size
    idx = [slice(None)] * x.ndim
    idx[ax] = slice(rank * b, (rank + 1) * b)
    return x[tuple(idx)]
