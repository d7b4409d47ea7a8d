"""Multi-GPU correctness of the slab transform under the current exchange mode (NCCL all-to-all, JFX_SLAB_CHUNKS=c,
JFX_SLAB_P2P=1): every rank also runs the SAME global transform on its own GPU and compares its block.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 --master-port 29533 \
        tools/check_slab_ranks.py [n]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import jaxfun_b200 as jf
from jaxfun_b200 import sharding as S

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
T = jf.TensorProduct(*[jf.Legendre(n)] * 3)
D = S.SlabTensorProduct(T)
g = torch.Generator(device=dev).manual_seed(1234)          # same global array on every rank
c = torch.randn(n, n, n, dtype=torch.float64, device=dev, generator=g)
u_ref = T.backward(c)
ok = True
for it in range(3):                                         # several rounds: exercises the alternating receive buffers
    u_loc = D.backward(S.local_block(c, S.SPECTRAL, rank, world).contiguous())
    e1 = float((u_loc - S.local_block(u_ref, S.PHYSICAL, rank, world)).abs().max() / u_ref.abs().max())
    c_loc = D.forward(u_loc)
    e2 = float((c_loc - S.local_block(c, S.SPECTRAL, rank, world)).abs().max() / c.abs().max())
    ok = ok and e1 < 1e-12 and e2 < 1e-11
    print(f"rank {rank} round {it}: backward {e1:.2e} forward(round trip) {e2:.2e}", flush=True)
# other spaces (round 2): the C-ABI slab transform handles any basis / dtype — where the contraction epilogue cannot
# store into the peers (FFT axes, complex data, 2-D) the exchange is one strided peer copy per rank
if os.environ.get("JFX_SLAB_CHECK_MORE", "1") != "0":
    m = 64
    cases = [("Chebyshev^3 f64", [jf.Chebyshev(m)] * 3, torch.float64),
             ("Fourier x Fourier x Legendre c128", [jf.Fourier(m), jf.Fourier(m), jf.Legendre(m)], torch.complex128),
             ("Fourier x Chebyshev c128 (2-D)", [jf.Fourier(2 * m), jf.Chebyshev(2 * m)], torch.complex128),
             ("Legendre^3 f64 (folded, peer stores)", [jf.Legendre(m)] * 3, torch.float64)]
    for name, spaces, dt in cases:
        T2 = jf.TensorProduct(*spaces)
        D2 = S.SlabTensorProduct(T2)
        shape = tuple(sp.N for sp in spaces)
        c2 = torch.randn(shape, dtype=torch.float64, device=dev, generator=g).to(dt)
        if dt.is_complex:
            c2 = c2 + 1j * torch.randn(shape, dtype=torch.float64, device=dev, generator=g)
        u2 = T2.backward(c2)
        for it in range(2):
            ul = D2.backward(S.local_block(c2, S.SPECTRAL, rank, world).contiguous())
            e1 = float((ul - S.local_block(u2, S.PHYSICAL, rank, world)).abs().max() / u2.abs().max())
            cl = D2.forward(ul)
            e2 = float((cl - S.local_block(c2, S.SPECTRAL, rank, world)).abs().max() / c2.abs().max())
            ok = ok and e1 < 1e-12 and e2 < 1e-11
        be = next(iter(D2._backends.values()))
        route = [("native fused" if v["fused"] else "native copies") for k, v in be._plans.items() if k[0] == "native" and v]
        print(f"rank {rank} {name}: backward {e1:.2e} round trip {e2:.2e} route {route or 'host-composed'}", flush=True)
t = torch.tensor([0 if ok else 1], device=dev)
dist.all_reduce(t)
if rank == 0:
    print("SLAB CHECK", "OK" if int(t.item()) == 0 else "FAILED", f"chunks={S.slab_chunks()} p2p={S.slab_p2p()} fused_pack={S.slab_fused_pack()} native={S.slab_native()}", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 0 else 1)
