"""Small driver for ncu: a few transforms of the bench workload (Legendre^3 / Chebyshev^3 256^3 fp64,
plus batched 1-D Fourier 65536 x 1024 and Fourier^2 4096^2) so `-k regex:` can pick a kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jaxfun_b200 as jf

which = sys.argv[1] if len(sys.argv) > 1 else "all"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda:0")
reps = 3
if which in ("all", "legendre"):
    T = jf.TensorProduct(*[jf.Legendre(n)] * 3)
    c = torch.randn(n, n, n, dtype=torch.float64, device=dev)
    for _ in range(reps):
        u = T.backward(c); T.forward(u)
if which in ("all", "chebyshev"):
    T = jf.TensorProduct(*[jf.Chebyshev(n)] * 3)
    c = torch.randn(n, n, n, dtype=torch.float64, device=dev)
    for _ in range(reps):
        u = T.backward(c); T.forward(u)
if which in ("all", "fourier1d"):
    F = jf.Fourier(1024)
    c = torch.randn(65536, 1024, dtype=torch.complex128, device=dev)
    for _ in range(reps):
        u = F.backward(c); F.forward(u)
if which in ("all", "fourier2d"):
    T = jf.TensorProduct(jf.Fourier(4096), jf.Fourier(4096))
    c = torch.randn(4096, 4096, dtype=torch.complex128, device=dev)
    for _ in range(reps):
        u = T.backward(c); T.forward(u)
torch.cuda.synchronize()
print("done", which)
