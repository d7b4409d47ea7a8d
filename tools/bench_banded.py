"""Time jfx_banded_solve on Fourier x polynomial shapes (CUDA events, inputs larger than L2 or L2 flushed by rotation).
    [JFX_LIB_PATH=jaxfun_b200/variants/libjfx_TAG.so] python tools/bench_banded.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jaxfun_b200.galerkin.tpsolve import WavenumberBandedSolver  # noqa: E402

dev = torch.device("cuda:0")
CASES = [((1024, 1022), 1), ((4096, 4094), 1), ((1022, 1024), 0), ((256, 256, 254), 2), ((256, 254, 256), 1), ((65536, 1022), 1)]
print("lib:", os.environ.get("JFX_LIB_PATH", "default"))
for shape, pa in CASES:
    n = shape[pa]
    n_sys = int(np.prod(shape)) // n
    rng = np.random.default_rng(1)
    P = np.zeros((2, 3, n))
    P[0, 1] = 4.0 + 0.01 * np.arange(n)
    P[1, 0, :n - 2] = -0.2
    P[1, 1] = 1.0
    P[1, 2, 2:] = -0.2
    W = np.stack([np.ones(n_sys), 1.0 + rng.random(n_sys)])
    S = WavenumberBandedSolver(pa, shape, W, P, (-2, 0, 2))
    nbuf = max(2, int(300e6 // (int(np.prod(shape)) * 16)) + 1)           # rotate buffers: > 2 x L2 between reuses
    nbuf = min(nbuf, 8)
    rs = [torch.randn(*shape, dtype=torch.complex128, device=dev) for _ in range(nbuf)]
    out = torch.empty_like(rs[0])
    for i in range(3):
        S.solve(rs[i % nbuf], out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for i in range(reps):
        S.solve(rs[i % nbuf], out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nbytes = 2 * rs[0].numel() * 16 + 5 * 8 * n * n_sys
    print(f"{str(shape):>18} axis {pa}: {ms * 1e3:8.1f} us   {nbytes / ms / 1e6:7.0f} GB/s compulsory (rhs in + x out + factors)")
    del rs, out, S
