"""Generate the tabulated Gauss-Legendre data (n <= 100) used by FastGL.

The reference (`src/jaxfun/utils/fastgl.py:13-225`) ships Bogaert's published
tables: for every rule n <= 100 the angles theta_k in (0, pi/2) of the
positive nodes x_k = cos(theta_k) (largest theta first), the matching weights,
and Cl[n] = P_n(0) (n even) / P_n'(0) (n odd) for the centre weight of odd rules.

Instead of copying those literals we recompute them here with mpmath at 60
digits (Newton on P_n(cos theta)), round to float64, and -- when the reference
checkout is available -- verify bit-for-bit against the literals it holds.
Output: jaxfun_b200/data/fastgl_tables.npz (committed).

Run:  python tools/gen_fastgl_tables.py [--verify-only]
"""
from __future__ import annotations

import os
import re
import sys

import mpmath as mp
import numpy as np

mp.mp.dps = 60
OUT = os.path.join(os.path.dirname(__file__), "..", "jaxfun_b200", "data", "fastgl_tables.npz")
REF = "/root/reference/src/jaxfun/utils/fastgl.py"


def legendre_and_derivative(n: int, x):
    """(P_n(x), P_n'(x)) by the three-term recurrence in mp arithmetic."""
    p0, p1 = mp.mpf(1), x
    if n == 0:
        return p0, mp.mpf(0)
    for k in range(2, n + 1):
        p0, p1 = p1, ((2 * k - 1) * x * p1 - (k - 1) * p0) / k
    dp = n * (x * p1 - p0) / (x * x - 1)
    return p1, dp


def positive_rule(n: int):
    """theta (descending) and weights of the strictly positive nodes of rule n."""
    x0, _ = np.polynomial.legendre.leggauss(n)
    pos = sorted(float(v) for v in x0 if v > 1e-12)  # ascending x == descending theta
    thetas, weights = [], []
    for xg in pos:
        x = mp.mpf(xg)
        for _ in range(8):
            p, dp = legendre_and_derivative(n, x)
            x = x - p / dp
        _, dp = legendre_and_derivative(n, x)
        thetas.append(mp.acos(x))
        weights.append(2 / ((1 - x * x) * dp * dp))
    return thetas, weights


def centre_constant(n: int):
    p, dp = legendre_and_derivative(n, mp.mpf(0)) if n > 0 else (mp.mpf(1), mp.mpf(0))
    if n % 2 == 0:
        return p
    # P_n'(0) = n P_{n-1}(0)
    pm1, _ = legendre_and_derivative(n - 1, mp.mpf(0))
    return n * pm1


def pub(v) -> float:
    """Round to the 25 significant digits of the published tables, then to float64."""
    return float(mp.nstr(v, 25, strip_zeros=False))


def generate():
    theta = np.zeros((101, 50))
    weight = np.zeros((101, 50))
    cl = np.zeros(101)
    for n in range(0, 101):
        cl[n] = pub(centre_constant(n))
    for n in range(2, 101):
        th, w = positive_rule(n)
        theta[n, : len(th)] = [pub(t) for t in th]
        weight[n, : len(w)] = [pub(v) for v in w]
    return theta, weight, cl


def parse_reference():
    """Pull the literal tables out of the reference source text (no jax import)."""
    src = open(REF).read()
    arr = {}
    for m in re.finditer(r"^(\w+) = jnp\.array\(\[([^\]]*)\]\)", src, re.M):
        arr[m.group(1)] = np.array([float(s) for s in m.group(2).split(",") if s.strip()])
    return arr


def verify(theta, weight, cl) -> int:
    ref = parse_reference()
    bad = 0
    for n in range(2, 101):
        m = n // 2
        if n % 2 == 0:
            t, w = ref[f"EvenThetaZero{m}"], ref[f"EvenW{m}"]
        else:
            t, w = ref[f"OddThetaZero{m}"], ref[f"OddW{m}"]
        dt = np.flatnonzero(t != theta[n, :m])
        dw = np.flatnonzero(w != weight[n, :m])
        for i in dt:
            print(f"theta mismatch n={n} i={i}: ref={t[i]!r} gen={theta[n, i]!r}")
        for i in dw:
            print(f"weight mismatch n={n} i={i}: ref={w[i]!r} gen={weight[n, i]!r}")
        bad += len(dt) + len(dw)
    c = ref["Cl"]
    dc = np.flatnonzero(c[:101] != cl[: len(c[:101])])
    for i in dc:
        print(f"Cl mismatch n={i}: ref={c[i]!r} gen={cl[i]!r}")
    bad += len(dc)
    print(f"verify: {bad} mismatching entries (len(Cl ref)={len(c)})")
    return bad


if __name__ == "__main__":
    if "--verify-only" in sys.argv:
        d = np.load(OUT)
        sys.exit(1 if verify(d["theta"], d["weight"], d["cl"]) else 0)
    theta, weight, cl = generate()
    np.savez_compressed(OUT, theta=theta, weight=weight, cl=cl)
    print("wrote", os.path.abspath(OUT))
    if os.path.exists(REF):
        verify(theta, weight, cl)
