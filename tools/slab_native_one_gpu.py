"""The C-ABI slab transform (jfx_slab_*) with P ranks EMULATED on one GPU: P slab objects, one stream per rank, "peer"
buffers that are simply the other ranks' buffers on the same device.  The flag barrier needs the ranks' kernels to be
co-resident, so this runs in its own process (a profiler that serialises kernels would make it time out and trap).
Reference: the single-device transform of the same global array.  Exit code 0 = all cases agree to 1e-12.

    python tools/slab_native_one_gpu.py [P]"""
import os
import sys
# the ranks' kernels must be able to run concurrently: with lazy module loading the FIRST launch of a kernel synchronises
# the context, i.e. waits for a barrier kernel that waits for a rank this host thread has not enqueued yet (CUDA
# programming guide, "Lazy Loading": kernels that must run concurrently).  Real ranks are separate processes.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jaxfun_b200 as jf
from jaxfun_b200 import _lib as L
from jaxfun_b200 import sharding as S


def run_all(slabs, xs, streams):
    outs = [None] * len(slabs)
    for r, (sl, x) in enumerate(zip(slabs, xs)):
        with torch.cuda.stream(streams[r]):
            outs[r] = sl(x)
    torch.cuda.synchronize()
    return outs


def make(T, op, shape, dt, sharding, P, dev):
    slabs = [S.NativeSlab(T, op, shape, dt, sharding, r, P) for r in range(P)]
    recv = [[torch.zeros(max(sl.recv_bytes, 8), dtype=torch.uint8, device=dev) for sl in slabs] for _ in range(2)]
    pads = [torch.zeros(max(sl.signal_bytes, 8), dtype=torch.uint8, device=dev) for sl in slabs]
    for sl in slabs:
        sl.bind([b.data_ptr() for b in recv[0]], [b.data_ptr() for b in recv[1]], [b.data_ptr() for b in pads])
    return slabs, (recv, pads)


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(7)
    streams = [torch.cuda.Stream() for _ in range(P)]
    cases = [("Legendre^3 f64", [jf.Legendre(64)] * 3, torch.float64),
             ("Chebyshev^3 f64", [jf.Chebyshev(64)] * 3, torch.float64),
             ("Fourier x Chebyshev x Legendre c128", [jf.Fourier(16), jf.Chebyshev(24), jf.Legendre(16)], torch.complex128),
             ("Fourier^2 c128", [jf.Fourier(32), jf.Fourier(64)], torch.complex128)]
    ok = True
    for name, spaces, dt in cases:
        T = jf.TensorProduct(*spaces)
        shape = tuple(sp.N for sp in spaces)
        c = torch.randn(shape, dtype=torch.float64, device=dev, generator=g).to(dt)
        if dt.is_complex:
            c = c + 1j * torch.randn(shape, dtype=torch.float64, device=dev, generator=g)
        u = T.backward(c)
        cb = [S.local_block(c, S.SPECTRAL, r, P).contiguous() for r in range(P)]
        bwd, keep1 = make(T, L.OP_BACKWARD, cb[0].shape, dt, S.SPECTRAL, P, dev)
        ub_shape = S.local_block(u, S.PHYSICAL, 0, P).shape
        fwd, keep2 = make(T, L.OP_FORWARD, ub_shape, dt, S.PHYSICAL, P, dev)
        for it in range(3):                        # three rounds: both receive buffers, flags raised and lowered repeatedly
            ub = run_all(bwd, cb, streams)
            e1 = max(float((ub[r] - S.local_block(u, S.PHYSICAL, r, P)).abs().max() / u.abs().max()) for r in range(P))
            cf = run_all(fwd, ub, streams)
            e2 = max(float((cf[r] - cb[r]).abs().max() / c.abs().max()) for r in range(P))
            ok = ok and e1 < 1e-12 and e2 < 1e-11
        print(f"P={P} {name}: backward {e1:.2e} round trip {e2:.2e} fused={bwd[0].fused}/{fwd[0].fused}", flush=True)
    print("SLAB NATIVE ONE GPU", "OK" if ok else "FAILED", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
