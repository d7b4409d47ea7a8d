"""Slab orchestration (jaxfun_b200.sharding) on world_size-2/4 CPU ranks over gloo.

The local phases are injected from the oracle (numpy); what is under test is the product's
exchange logic: axis order per phase, block layout around all_to_all_single, sharding contract
(spectral in -> physical out and back), divisibility error — sharding.py:43-105 of the reference
and tests/galerkin/test_forward_backward_spmd.py:51-114."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import jaxfun_oracle as O
from jaxfun_b200 import sharding as S


class OracleBackend(S.SlabBackend):
    """CPU stand-in for the engine: oracle 1-D transforms, torch reshapes for the repacks."""

    def __init__(self, spaces, op, N=None):
        self.spaces, self.op, self.N = spaces, op, N

    def apply_axes(self, x, axes):
        a = x.numpy()
        for ax in axes:
            sp = self.spaces[ax]
            if self.op == "backward":
                a = sp.backward(a, N=None if self.N is None else self.N[ax], axis=ax)
            elif self.op == "forward":
                a = sp.forward(a, axis=ax)
            else:
                a = sp.scalar_product(a, axis=ax)
        return torch.from_numpy(np.ascontiguousarray(a))

    def pack(self, x, split_axis, parts):
        return torch.stack(torch.chunk(x, parts, dim=split_axis), dim=0).contiguous()

    def unpack(self, blocks, concat_axis, parts):
        return torch.cat([blocks[p] for p in range(parts)], dim=concat_axis).contiguous()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, names, N, seed, q, chunks=1):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cls = {"C": O.Chebyshev, "L": O.Legendre, "F": O.Fourier}
        spaces = [cls[n](Ni) for n, Ni in zip(names, N)]
        T = O.TensorProductSpace(*spaces)
        rng = np.random.default_rng(seed)
        c = rng.standard_normal(N) + (1j * rng.standard_normal(N) if "F" in names else 0)
        u_ref = T.backward(c)
        c_loc = torch.from_numpy(np.ascontiguousarray(S.local_block(c, S.SPECTRAL, rank, world)))
        u_loc = S.apply_separable_slab(c_loc, S.SPECTRAL, OracleBackend(spaces, "backward"), world, chunks)
        e1 = np.abs(u_loc.numpy() - S.local_block(u_ref, S.PHYSICAL, rank, world)).max()
        c_back = S.apply_separable_slab(u_loc, S.PHYSICAL, OracleBackend(spaces, "forward"), world, chunks)
        e2 = np.abs(c_back.numpy() - S.local_block(c, S.SPECTRAL, rank, world)).max()
        sp_loc = S.apply_separable_slab(u_loc, S.PHYSICAL, OracleBackend(spaces, "scalar_product"), world, chunks)
        e3 = np.abs(sp_loc.numpy() - S.local_block(T.scalar_product(u_ref), S.SPECTRAL, rank, world)).max()
        q.put((rank, float(e1), float(e2), float(e3), tuple(u_loc.shape), tuple(c_back.shape)))
    finally:
        dist.destroy_process_group()


_CASES = [("CC", (8, 12)), ("FL", (8, 8)), ("CCC", (8, 8, 6)), ("FCL", (8, 12, 5)), ("FFL", (8, 4, 7))]


@pytest.mark.parametrize("world,chunks,names,N",
                         [(w, 1, nm, N) for w in (2, 4) for nm, N in _CASES] +
                         [(2, 2, "CC", (8, 12)), (2, 3, "FCL", (8, 12, 5)), (4, 2, "CCC", (8, 8, 6)), (2, 3, "FFL", (8, 4, 7))])
def test_slab_roundtrip_gloo(world, chunks, names, N):
    """chunks > 1: the overlapped exchange (chunked all-to-all interleaved with the local passes; ragged last chunk when
    the chunk count does not divide the local extent) must give the same blocks as the single all-to-all."""
    if any(n % world for n in N[:2]):
        pytest.skip("extent not divisible")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, names, N, 11, q, chunks)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, e1, e2, e3, ushape, cshape in res:
        assert e1 < 1e-12 and e2 < 1e-12 and e3 < 1e-12, (rank, e1, e2, e3)
        exp_u = list(N); exp_u[1] //= world
        exp_c = list(N); exp_c[0] //= world
        assert ushape == tuple(exp_u) and cshape == tuple(exp_c)   # physical: axis 1 sharded; spectral: axis 0


def test_slab_matches_oracle_simulation():
    """Single-process cross-check of the block algebra against oracle.slab_transform (P simulated ranks)."""
    rng = np.random.default_rng(0)
    N, P = (8, 12, 6), 4
    spaces = [O.Chebyshev(N[0]), O.Legendre(N[1]), O.Chebyshev(N[2])]
    T = O.TensorProductSpace(*spaces)
    c = rng.standard_normal(N)
    fns = [lambda a, ax=ax: spaces[ax].backward(a, axis=ax) for ax in range(3)]
    blocks = [S.local_block(c, S.SPECTRAL, r, P) for r in range(P)]
    out = O.slab_transform(fns, blocks, sharded_axis=0, split_axis=1)
    u = T.backward(c)
    for r in range(P):
        assert np.abs(out[r] - S.local_block(u, S.PHYSICAL, r, P)).max() < 1e-12


def test_indivisible_split_axis_raises():
    class Ident(S.SlabBackend):
        def apply_axes(self, x, axes):
            return x
    with pytest.raises(ValueError):
        S.apply_separable_slab(torch.zeros(4, 7), S.SPECTRAL, Ident(), 2)
    with pytest.raises(ValueError):
        S.local_block(np.zeros((5, 4)), S.SPECTRAL, 0, 2)
    assert S.get_transposed_sharding(S.SPECTRAL) == S.PHYSICAL
    assert S.get_transposed_sharding(S.PHYSICAL) == S.SPECTRAL
