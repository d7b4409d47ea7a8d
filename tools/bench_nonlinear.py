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
    comp = 2 * uh.numel() * 16
    print(f"{a.what} n={n} fused={os.environ.get('JFX_NL_FUSE', '1')}: {ms * 1e3:9.1f} us per nonlinear term, "
          f"launches={nl.launches(uh)}, compulsory {comp / 1e6:.0f} MB -> {comp / ms / 1e6:.1f} GB/s")


if __name__ == "__main__":
    main()
