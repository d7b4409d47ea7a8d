"""Shared cases of the banded-solver tests: the same systems go through the host emulation of the device code
(test_banded_emul.py, CPU) and through the CUDA kernels (test_banded_gpu.py)."""
import numpy as np

# name: (shape, poly_axis, offsets, n_terms, complex band, rhs dtype)
CASES = {
    "2d_last_penta":      ((37, 50), 1, (-2, 0, 2), 2, False, "complex128"),       # Fourier x Legendre-Dirichlet (W = 2, U = 8)
    "2d_first_penta":     ((50, 37), 0, (-2, 0, 2), 2, False, "complex128"),       # polynomial axis first: inner = n_F
    "3d_last":            ((6, 7, 33), 2, (-2, 0, 2), 3, False, "complex128"),     # Fourier x Fourier x polynomial
    "3d_middle":          ((5, 29, 9), 1, (-2, -1, 0, 1, 2), 2, False, "complex128"),
    "real_tridiag":       ((40, 19), 1, (-1, 0, 1), 1, False, "float64"),
    "biharmonic_w4":      ((12, 41), 1, (-4, -2, 0, 2, 4), 3, False, "complex128"),  # W = 4, U = 4
    "w8":                 ((9, 35), 1, (-7, -3, 0, 2, 8), 2, False, "complex128"),   # W = 8, U = 2
    "generic_dense_upper": ((7, 26), 1, (-2, 0, 2, 4, 6, 8, 10, 12, 14, 16, 18, 20, 22, 24), 2, False, "complex128"),
    "complex_band":       ((11, 23), 1, (-2, -1, 0, 1, 2), 2, True, "complex128"),
    "complex_band_generic": ((4, 30), 1, (-9, 0, 3, 11), 2, True, "complex128"),
    "single":             ((5, 16), 1, (-2, 0, 2), 2, False, "complex64"),
    "single_real":        ((16, 5), 0, (-1, 0, 2), 2, False, "float32"),
    "upper_only":         ((6, 12), 1, (0, 1, 3), 1, False, "float64"),
    "lower_only":         ((6, 12), 1, (-3, -1, 0), 1, False, "complex128"),
    "diagonal":           ((6, 12), 1, (0,), 2, False, "complex128"),
    "n_one":              ((9, 1), 1, (0,), 1, False, "float64"),
    "rows_kernel_w2":     ((19000, 21), 1, (-2, 0, 2), 2, False, "complex128"),     # enough systems for the row-tile kernel on the GPU
    "rows_kernel_w4":     ((593, 33, 13), 2, (-4, -1, 0, 3), 2, True, "complex128"),
    "rows_kernel_f64":    ((19010, 19), 1, (-1, 0, 1), 1, False, "float64"),
    "rows_kernel_c64":    ((19010, 18), 1, (-2, 0, 2), 2, False, "complex64"),
    "chunk_kernel_many":  ((33, 21, 576), 1, (-2, 0, 2), 2, False, "complex128"),    # >= 18 944 systems, axis not last: register chunks
    "edge_n3":            ((5, 3), 1, (-2, 0, 2), 2, False, "complex128"),           # shorter than the window reach and the chunk
    "edge_n12":           ((3, 12), 1, (-2, 0, 2), 2, False, "complex128"),          # exactly one chunk
    "edge_n13":           ((13, 3), 0, (-2, 0, 2), 2, False, "float64"),
    "edge_n24":           ((2, 24, 2), 1, (-2, 0, 2), 2, False, "complex128"),       # whole chunks only
    "edge_n9_w4":         ((4, 9), 1, (-4, 0, 4), 2, False, "complex128"),
    "many_systems":       ((300, 20), 1, (-2, 0, 2), 2, False, "complex128"),      # several CTAs, ragged last one
}


def make_case(name):
    shape, pa, offsets, n_terms, cband, dt = CASES[name]
    rng = np.random.default_rng(abs(hash(name)) % (2**32) if False else sum(map(ord, name)))
    n = shape[pa]
    n_sys = int(np.prod(shape)) // n
    P = rng.standard_normal((n_terms, len(offsets), n))
    W = rng.standard_normal((n_terms, n_sys))
    if cband:
        P = P + 1j * rng.standard_normal(P.shape)
        W = W + 1j * rng.standard_normal(W.shape)
    B = np.einsum("tf,tdp->fdp", W, P)
    main = offsets.index(0)
    P[0, main, :] += np.abs(B).sum(axis=1).max() + 1.0       # diagonally dominant: LU without pivoting is stable
    W[0, :] = np.abs(W[0, :]) + 1.0
    rhs = rng.standard_normal(shape)
    if dt.startswith("complex"):
        rhs = rhs + 1j * rng.standard_normal(shape)
    return shape, pa, offsets, W, P, rhs.astype(dt)


def tolerance(dt):
    return 1e-12 if dt in ("float64", "complex128") else 2e-5
