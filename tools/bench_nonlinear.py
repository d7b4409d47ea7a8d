"""Timing of the nonlinear-term evaluation (jfx_nonlinear_execute), fused vs JFX_NL_FUSE=0.

    python tools/bench_nonlinear.py kdv [--n 1024 --batch 65536]
    python tools/bench_nonlinear.py ch  [--n 1024]
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jaxfun_b200 as jf
from jaxfun_b200.integrators.nonlinear import NonlinearTerm, field


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what")
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--step", action="store_true", help="also time one full ETDRK4 step (ch)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    n = a.n
    if a.what == "kdv":
        V = jf.Fourier(n)
        u, (x,) = field(V)
        nl = NonlinearTerm(V, -u * u.diff(x))
        uh = torch.randn(a.batch, n, dtype=torch.complex128, device=dev)
    else:
        V = jf.TensorProduct(jf.Fourier(n, domain=(0, 1)), jf.Fourier(n, domain=(0, 1)))
        u, (x, y) = field(V)
        nl = NonlinearTerm(V, -(6 * u * (u.diff(x)**2 + u.diff(y)**2) + 3 * u**2 * (u.diff(x, 2) + u.diff(y, 2))))
        uh = 1e-2 * torch.randn(n, n, dtype=torch.complex128, device=dev)
    outs = [nl(uh) for _ in range(3)]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for i in range(reps):
        nl(uh, outs[i % 3])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if a.what == "ch" and a.step:
        # one full ETDRK4 step of Cahn-Hilliard (examples/cahn_hilliard2D_etdrk4.py:52-53, 86): 4 nonlinear terms +
        # the diagonal stage arithmetic
        import numpy as np
        from jaxfun_b200.integrators import ETDRK4
        k = np.asarray(V.basespaces[0].wavenumbers(), dtype=float) * float(V.basespaces[0].domain_factor)
        k2 = k[:, None] ** 2 + k[None, :] ** 2
        Ldiag = torch.from_numpy((-k2 - 1.5e-2 * k2**2) * (1.0 + 0j)).to(dev)   # nu = -1 Laplacian + mu = -1.5e-2 bi-Laplacian
        integ = ETDRK4(V, linear_diag=Ldiag, nonlinear=nl)
        dt = 5e-2 / 320
        u1 = integ.step(uh, dt)
        torch.cuda.synchronize()
        e0.record()
        for i in range(5):
            u1 = integ.step(uh, dt)
        e1.record()
        torch.cuda.synchronize()
        print(f"ch n={n} ETDRK4 step: {e0.elapsed_time(e1) / 5:.3f} ms  (finite: {bool(torch.isfinite(torch.view_as_real(u1)).all())})")
    comp = 2 * uh.numel() * 16
    print(f"{a.what} n={n} fused={os.environ.get('JFX_NL_FUSE', '1')}: {ms * 1e3:9.1f} us per nonlinear term, "
          f"launches={nl.launches(uh)}, compulsory {comp / 1e6:.0f} MB -> {comp / ms / 1e6:.1f} GB/s")


if __name__ == "__main__":
    main()
