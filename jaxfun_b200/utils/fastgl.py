"""Gauss-Legendre nodes and weights, FastGL style (host side, float64).

Same contract as the reference's `jaxfun.utils.fastgl.leggauss`
(`src/jaxfun/utils/fastgl.py:562-567`): an array of shape (2, N) holding the
nodes in ascending order and the matching weights.

* N <= 100: tabulated angles/weights (`jaxfun_b200/data/fastgl_tables.npz`,
  regenerated with mpmath by `tools/gen_fastgl_tables.py` and verified
  bit-for-bit against the reference's literals), mirrored about pi/2
  (reference `GLPairTabulated`, fastgl.py:512-544).
* N > 100: Bogaert's iteration-free asymptotic expansion (reference
  `GLPairS`/`besseljzero`/`besselj1squared`, fastgl.py:232-508) evaluated for
  all k at once with numpy, in the reference's operation order.

These tables are produced once per plan on the host and shipped to the GPU as
part of the transform matrices; they are not a compute fallback.
"""
from __future__ import annotations

import functools
import json
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(__file__), "..", "data")


@functools.lru_cache(maxsize=1)
def _tables():
    d = np.load(os.path.join(_DATA, "fastgl_tables.npz"))
    return d["theta"], d["weight"], d["cl"]


@functools.lru_cache(maxsize=1)
def _coeffs():
    with open(os.path.join(_DATA, "fastgl_coeffs.json")) as f:
        raw = json.load(f)
    return {k: np.array([float(s) for s in v]) for k, v in raw.items()}


def _horner_desc(coef: np.ndarray, x: np.ndarray) -> np.ndarray:
    """((c0*x + c1)*x + c2)... with c0 the highest power."""
    acc = coef[0] * x + coef[1]
    for c in coef[2:]:
        acc = acc * x + c
    return acc


def _horner_asc(coef: np.ndarray, x: np.ndarray) -> np.ndarray:
    """c0 + x*(c1 + x*(...)) with c0 the constant term."""
    acc = coef[-2] + coef[-1] * x
    for c in coef[-3::-1]:
        acc = c + x * acc
    return acc


def _bessel_j0_zero(k: np.ndarray) -> np.ndarray:
    """k-th positive zero of J0 (k >= 1): table for k <= 20, McMahon beyond."""
    C = _coeffs()
    kk = k.astype(np.float64)
    z = np.pi * (kk - 0.25)
    r = 1.0 / z
    r2 = r * r
    big = z + r * _horner_asc(C["kb20"], r2)
    small = C["JZ"][np.clip(k - 1, 0, 19)]
    return np.where(k > 20, big, small)


def _bessel_j1_squared(k: np.ndarray) -> np.ndarray:
    """J1(j_{0,k})^2: table for k <= 21, asymptotic series beyond."""
    C = _coeffs()
    x = 1.0 / (k.astype(np.float64) - 0.25)
    x2 = x * x
    c = C["km21"]
    big = x * (c[0] + x2 * x2 * _horner_asc(c[1:], x2))
    small = C["J1"][np.clip(k - 1, 0, 20)]
    return np.where(k > 21, big, small)


def _asymptotic_pairs(n: int, k: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """(theta_k, w_k) of the n-point rule for 1 <= k <= ceil(n/2) (theta < pi/2)."""
    C = _coeffs()
    w = 1.0 / (n + 0.5)
    nu = _bessel_j0_zero(k)
    theta = w * nu
    x = theta * theta
    B = _bessel_j1_squared(k)
    sf1, sf2, sf3 = (_horner_desc(C[s], x) for s in ("SF1T", "SF2T", "SF3T"))
    wsf1, wsf2, wsf3 = (_horner_desc(C[s], x) for s in ("WSF1T", "WSF2T", "WSF3T"))
    nu_o_sin = nu / np.sin(theta)
    b_nu_o_sin = B * nu_o_sin
    w_inv_sinc = w * w * nu_o_sin
    wis2 = w_inv_sinc * w_inv_sinc
    theta = w * (nu + theta * w_inv_sinc * (sf1 + wis2 * (sf2 + wis2 * sf3)))
    deno = b_nu_o_sin + b_nu_o_sin * wis2 * (wsf1 + wis2 * (wsf2 + wis2 * wsf3))
    return theta, (2.0 * w) / deno


def gl_theta_weights(N: int) -> tuple[np.ndarray, np.ndarray]:
    """Angles (descending, so cos is ascending) and weights of the N-point rule."""
    if N < 1:
        raise ValueError("N must be >= 1")
    if N == 1:
        # n=1 is odd with n2=0: only the centre node exists (Cl[1] = 1 -> w = 2).
        return np.array([np.pi / 2]), np.array([2.0])
    if N <= 100:
        TH, W, CL = _tables()
        m = N // 2
        th_pos, w_pos = TH[N, :m], W[N, :m]  # descending theta in (0, pi/2)
        # ascending x: first the mirrored (negative) half, smallest theta first
        neg_t = np.pi - th_pos[::-1]
        neg_w = w_pos[::-1]
        if N % 2:
            theta = np.concatenate([neg_t, [np.pi / 2], th_pos])
            weight = np.concatenate([neg_w, [2.0 / (CL[N] * CL[N])], w_pos])
        else:
            theta = np.concatenate([neg_t, th_pos])
            weight = np.concatenate([neg_w, w_pos])
        return theta, weight
    # reference: leggauss -> GLPair(N, kk) for kk = N..1 (fastgl.py:548-567)
    kk = np.arange(N, 0, -1)
    mirrored = 2 * kk - 1 > N
    ks = np.where(mirrored, N - kk + 1, kk)
    theta, weight = _asymptotic_pairs(N, ks)
    theta = np.where(mirrored, np.pi - theta, theta)
    return theta, weight


@functools.lru_cache(maxsize=64)
def _leggauss_cached(N: int) -> np.ndarray:
    theta, weight = gl_theta_weights(N)
    out = np.stack([np.cos(theta), weight])
    out.setflags(write=False)
    return out


def leggauss(N: int) -> np.ndarray:
    """Array (2, N): Gauss-Legendre nodes (ascending) and weights."""
    return _leggauss_cached(int(N))
