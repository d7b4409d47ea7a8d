"""Chebyshev^3 (or Fourier^3) n^3 backward / forward timing under the current environment switches
(JFX_LIB_PATH, JFX_PAIR, JFX_FFT_STREAM ...), with a correctness check against the plain per-axis transform
of the default library path being unnecessary: the round trip error is printed instead.

    python tools/bench_cheb3.py [--n 256] [--basis cheb|four] [--reps 20]"""
import argparse
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jaxfun_b200 as jf


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--basis", default="cheb")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    n = a.n
    if a.basis == "cheb":
        T = jf.TensorProduct(*[jf.Chebyshev(n)] * 3)
        cs = [torch.randn(n, n, n, dtype=torch.float64, device=dev) for _ in range(3)]
    else:
        T = jf.TensorProduct(*[jf.Fourier(n)] * 3)
        cs = [torch.randn(n, n, n, dtype=torch.complex128, device=dev) for _ in range(3)]
    us = [T.backward(c) for c in cs]
    err = float((T.forward(us[0]) - cs[0]).abs().max() / cs[0].abs().max())
    res = {}
    for name, fn, xs in (("backward", T.backward, cs), ("forward", T.forward, us)):
        for i in range(3):
            fn(xs[i % 3])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.reps):
            fn(xs[i % 3])
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / a.reps * 1e3
    nbytes = cs[0].numel() * cs[0].element_size()
    pair = res["backward"] + res["forward"]
    print(f"[{a.tag}] {a.basis}^3 n={n}: backward {res['backward']:.1f} us, forward {res['forward']:.1f} us, pair {pair:.1f} us, "
          f"8(d) {4 * nbytes / pair / 1e3:.0f} GB/s = {4 * nbytes / pair / 1e3 / 6544.7:.3f} of HBM, round trip {err:.1e}", flush=True)


if __name__ == "__main__":
    main()
