"""Per-axis timing of the fast transform kernels (CUDA events, inputs > L2 rotated between launches).

    python tools/bench_axes.py [cheb|four|four2d|batched] [--n 256]

Prints one line per (op, axis): microseconds per launch and the compulsory GB/s (read once + write once)."""
import argparse
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jaxfun_b200 as jf
from jaxfun_b200 import _lib as L
from jaxfun_b200.engine import Plan, jfx_dtype


def time_plan(plan, x, reps=20, nbuf=3):
    xs = [x.clone() for _ in range(nbuf)]
    outs = [torch.empty(plan.shape_out, dtype=x.dtype, device=x.device) for _ in range(nbuf)]
    for i in range(3):
        plan.execute(xs[i % nbuf], outs[i % nbuf])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        plan.execute(xs[i % nbuf], outs[i % nbuf])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", nargs="?", default="cheb")
    ap.add_argument("--n", type=int, default=256)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    n = a.n
    if a.what == "cheb":
        shape, dt, sp = (n, n, n), torch.float64, jf.Chebyshev(n)
    elif a.what == "four":
        shape, dt, sp = (n, n, n), torch.complex128, jf.Fourier(n)
    elif a.what == "four2d":
        shape, dt, sp = (n, n), torch.complex128, jf.Fourier(n)
    elif a.what == "batched":
        shape, dt, sp = (65536, n), torch.complex128, jf.Fourier(n)
    else:
        raise SystemExit("unknown workload")
    x = torch.randn(shape, dtype=torch.float64, device=dev).to(dt)
    dtype = jfx_dtype(dt)
    nbytes = x.numel() * x.element_size()
    axes = range(len(shape)) if a.what != "batched" else [1]
    for op, name in ((L.OP_BACKWARD, "backward"), (L.OP_FORWARD, "forward")):
        for ax in axes:
            specs = [None] * len(shape)
            inner = 1
            for s in shape[ax + 1:]:
                inner *= s
            specs[ax] = sp.axis_spec(op, shape[ax], dtype, inner=inner)
            plan = Plan(op, dtype, shape, specs)
            us = time_plan(plan, x)
            print(f"{a.what} n={n} {name:9s} axis {ax}: {us:8.1f} us  {2 * nbytes / us / 1e3:8.1f} GB/s", flush=True)


if __name__ == "__main__":
    main()
