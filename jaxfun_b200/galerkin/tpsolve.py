"""Tensor-product Galerkin solves — the factored paths of `jaxfun.la.tpmatrix.TPMatrices.solve(method="lu")`
(`src/jaxfun/la/tpmatrix.py:386-587`), SURVEY.md §8(f) rank 3:

* `KroneckerSumSolver` / `tpmats_lu_factor`: per-axis diagonalisation (tpmatrix.py:429-587, 1089-1233), described below;
* `WavenumberBandedSolver` / `tpmats_wavenumber_factor`: one banded LU per Fourier wavenumber combination for Fourier x
  polynomial operators (tpmatrix.py:590-1014, 1236-1354), assembled, factored and solved in libjfx.so (`jfx_banded_*`);
* `TPMatrices(terms)`: the reference's dispatch between the two with cached factors (tpmatrix.py:386-427).

A separable operator  sum_i (B_0 x .. x A_i x .. x B_{d-1})  (e.g. the weak Laplacian with stiffness A_i
and mass B_i per axis) is diagonalised by the generalised eigenvectors A_i V_i = B_i V_i Lambda_i:

        u = (V_0 x .. x V_{d-1})  [ ((B_0 V_0)^-1 x .. x (B_{d-1} V_{d-1})^-1) f  /  (lambda_0 (+) .. (+) lambda_{d-1}) ]

The factorisation (O(n^3) per axis) stays on the host; the solve is two multi-axis `JFX_OP_APPLY` plans —
the same FP64 tensor-core contraction as the Vandermonde transforms, with eigenvector tables — and one
diagonal scaling (`jfx_axpby_diag`).
"""
from __future__ import annotations

import numpy as np

from .. import _lib as L
from ..engine import AxisSpec, Plan, device_key, device_scope, jfx_dtype


def stiffness_matrix(space, k: int = 2, nq: int | None = None) -> np.ndarray:
    """(phi_i, d^k phi_j / dx^k)_w by Gauss quadrature of the space (exact for polynomial bases with nq >= N):
    what `inner(v * Div(Grad(u)))` assembles per axis (`galerkin/inner.py:809-912`)."""
    o = space.get_orthogonal()
    nq = o.N if nq is None else nq
    xq, wq = o.quad_points_and_weights(nq)
    df = float(o.domain_factor)
    P = space.eval_basis_functions(xq)                    # [nq, dim]
    D = space.evaluate_basis_derivative(xq, k)            # [nq, dim]
    return (P * (wq / df)[:, None]).T @ D * df ** k


def mass_matrix(space, nq: int | None = None) -> np.ndarray:
    o = space.get_orthogonal()
    nq = o.N if nq is None else nq
    xq, wq = o.quad_points_and_weights(nq)
    P = space.eval_basis_functions(xq)
    return (P * (wq / float(o.domain_factor))[:, None]).T @ P


class KroneckerSumSolver:
    """Solve  sum_i (B_0 x .. x A_i x .. x B_{d-1}) u = f  for coefficient arrays on the device."""

    def __init__(self, mats):
        self.V, self.W, lams = [], [], []
        for A, B in mats:
            lam, V = np.linalg.eig(np.linalg.solve(B, A))
            if np.abs(lam.imag).max() > 1e-9 * np.abs(lam).max():
                raise ValueError("operator pencil has complex eigenvalues: not diagonalisable over the reals")
            order = np.argsort(lam.real)
            lam, V = lam.real[order], np.ascontiguousarray(V.real[:, order])
            self.V.append(V)
            self.W.append(np.ascontiguousarray(np.linalg.inv(B @ V)))
            lams.append(lam)
        d = len(mats)
        D = 0.0
        for i, lam in enumerate(lams):
            shp = [1] * d
            shp[i] = -1
            D = D + lam.reshape(shp)
        self.Dinv = np.ascontiguousarray(1.0 / D)
        self._plans = {}
        self._dinv_dev = {}

    def solve(self, f):
        import torch
        from ..integrators.base import axpby_diag
        dt = jfx_dtype(f.dtype)
        key = (tuple(f.shape), dt, device_key(f))
        if key not in self._plans:
            lead = f.ndim - len(self.V)
            with device_scope(f):
                pw = Plan(L.OP_APPLY, dt, tuple(f.shape), [None] * lead + [AxisSpec(L.BASIS_TABLE, table=W) for W in self.W])
                pv = Plan(L.OP_APPLY, dt, pw.shape_out, [None] * lead + [AxisSpec(L.BASIS_TABLE, table=V) for V in self.V])
            self._plans[key] = (pw, pv)
            dinv = torch.from_numpy(np.broadcast_to(self.Dinv, pw.shape_out).copy()).to(f.device)
            self._dinv_dev[key] = dinv
        pw, pv = self._plans[key]
        g = pw(f.contiguous())
        g = axpby_diag([(1.0, self._dinv_dev[key], g)])
        return pv(g)


# =====================================================================================================================
# Fourier x polynomial systems: one banded solve per Fourier wavenumber combination
# =====================================================================================================================
def dia_from_dense(A, tol: float = 1e-10):
    """(offsets, data) of a square matrix in column-aligned DIA form, data[d][j] = A[j - offsets[d], j] — the layout of
    `DiaMatrix.data` the reference's banded LU works on (la/diamatrix.py:1944).  Diagonals whose largest entry is below
    tol * max|A| are dropped (quadrature round-off of structurally zero diagonals)."""
    A = np.asarray(A)
    n = A.shape[0]
    scale = float(np.abs(A).max()) if A.size else 0.0
    offs, rows = [], []
    for off in range(-(n - 1), n):
        dg = np.diagonal(A, off)
        if np.abs(dg).max(initial=0.0) > tol * scale:
            row = np.zeros(n, dtype=A.dtype)
            if off >= 0:
                row[off:] = dg
            else:
                row[:n + off] = dg
            offs.append(off)
            rows.append(row)
    return tuple(offs), np.array(rows)


class SolverNotApplicable(ValueError):
    """The operator does not have the Fourier x polynomial structure (la/tpmatrix.py: `SolverNotApplicable`)."""


class WavenumberBandedSolver:
    """`TPMatricesWavenumberSolver` (la/tpmatrix.py:686-1014 of the reference) on the device.

    The systems are given in the separable form `tpmats_wavenumber_factor` derives (tpmatrix.py:1306-1347):
    B_s = sum_t weights[t, s] * P_t with `diags[t]` the DIA data of P_t on the union `offsets`; s runs over the Fourier axes
    of `shape` in C order.  Assembly, LU without pivoting and both substitution sweeps run in libjfx.so (`jfx_banded_*`,
    csrc/kernels_banded.cu); `solve` takes the right-hand side in its natural layout (polynomial axis = `poly_axis`)."""

    def __init__(self, poly_axis: int, shape, weights, diags, offsets):
        self.poly_axis, self.shape = int(poly_axis) % len(shape), tuple(int(v) for v in shape)
        self.offsets = tuple(int(o) for o in offsets)
        n = self.shape[self.poly_axis]
        n_sys = int(np.prod([v for a, v in enumerate(self.shape) if a != self.poly_axis], dtype=np.int64))
        cplx = np.iscomplexobj(weights) or np.iscomplexobj(diags)
        want = np.complex128 if cplx else np.float64
        self.weights = np.ascontiguousarray(weights, dtype=want)
        self.diags = np.ascontiguousarray(diags, dtype=want)
        if self.weights.ndim != 2 or self.weights.shape[1] != n_sys:
            raise ValueError(f"weights must be [n_terms, {n_sys}] (one column per Fourier wavenumber combination)")
        if self.diags.shape != (self.weights.shape[0], len(self.offsets), n):
            raise ValueError(f"diags must be [n_terms, n_diags, n] = {(self.weights.shape[0], len(self.offsets), n)}")
        self.band_complex = bool(cplx)
        self.n, self.n_sys = n, n_sys
        self._lib = L.load()
        self._handles: dict = {}

    def shard(self, rank: int, size: int) -> "WavenumberBandedSolver":
        """The solver of rank `rank`'s block of a coefficient array slab-sharded along axis 0 into `size` equal blocks — the
        multi-device mode of the reference (la/tpmatrix.py:786-812, 985-1013): axis 0 must be a Fourier axis, every device
        keeps the factors of its own contiguous block of wavenumbers and the solve needs no communication."""
        if self.poly_axis == 0:
            raise ValueError(f"Multi-process solve requires axis 0 to be a Fourier axis (poly_axis=0 not supported). "
                             f"Got shape={self.shape}, poly_axis={self.poly_axis}.")
        n0 = self.shape[0]
        if not 0 <= rank < size or n0 % size:
            raise ValueError(f"axis 0 of extent {n0} cannot be split into {size} equal blocks (rank {rank})")
        per = self.n_sys // size                     # C order over the Fourier axes: a block of axis 0 is a contiguous range
        local_shape = (n0 // size,) + self.shape[1:]
        return WavenumberBandedSolver(self.poly_axis, local_shape, self.weights[:, rank * per:(rank + 1) * per], self.diags,
                                      self.offsets)

    def _handle(self, dtype: int, device=None):
        """Factors of the systems in the precision of `dtype` on CUDA device `device` (created on first use: the object
        is assembled and factored on the device that is current, so the device of the right-hand side is made current)."""
        if device is None:
            import torch
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        h = self._handles.get((dtype, device))
        if h is None:
            import ctypes as C
            d = L.BandedDesc()
            d.abi_version, d.dtype, d.band_complex = L.JFX_ABI_VERSION, int(dtype), int(self.band_complex)
            d.n_terms, d.n, d.n_sys, d.n_diags = self.weights.shape[0], self.n, self.n_sys, len(self.offsets)
            offs = (C.c_int32 * len(self.offsets))(*self.offsets)
            d.offsets = offs
            d.weights, d.diags = self.weights.ctypes.data, self.diags.ctypes.data
            h = C.c_void_p()
            rc = self._lib.jfx_banded_create(C.byref(d), C.byref(h))
            if rc == -2:      # zero / non-finite pivot: the reference raises ValueError (la/diamatrix.py:461-471)
                raise ValueError(self._lib.jfx_last_error().decode())
            L.check(rc)
            self._handles[(dtype, device)] = h
        return h

    def __del__(self):
        try:
            for h in self._handles.values():
                self._lib.jfx_banded_destroy(h)
            self._handles = {}
        except Exception:
            pass

    def bandwidths(self, dtype: int = L.C128):
        import ctypes as C
        p, q, nb = C.c_int32(), C.c_int32(), C.c_size_t()
        L.check(self._lib.jfx_banded_info(self._handle(dtype), C.byref(p), C.byref(q), C.byref(nb)))
        return int(p.value), int(q.value), int(nb.value)

    def factors(self, dtype: int = L.C128) -> np.ndarray:
        """The factored band [n_sys, p + q + 1, n] (the reference's `band_lu` layout, la/tpmatrix.py:765-767) on the host."""
        p, q, _ = self.bandwidths(dtype)
        real = np.float64 if dtype in (L.F64, L.C128) else np.float32
        et = (np.complex128 if real is np.float64 else np.complex64) if self.band_complex else real
        buf = np.empty((p + q + 1, self.n, self.n_sys), dtype=et)
        L.check(self._lib.jfx_banded_factors(self._handle(dtype), buf.ctypes.data))
        return np.ascontiguousarray(np.transpose(buf, (2, 0, 1)))

    def solve(self, rhs, out=None):
        import ctypes as C

        from ..engine import current_stream_ptr
        if tuple(rhs.shape) != self.shape:
            raise ValueError(f"right-hand side has shape {tuple(rhs.shape)}, the solver was built for {self.shape}")
        if not getattr(rhs, "is_cuda", False):
            raise L.JfxError(-3, "device path needs a CUDA tensor; jaxfun_b200 has no CPU fallback")
        dt = jfx_dtype(rhs.dtype)
        if self.band_complex and dt in (L.F32, L.F64):
            raise TypeError("complex systems need a complex right-hand side")
        rhs = rhs.contiguous()
        if out is None:
            import torch
            out = torch.empty_like(rhs)
        elif tuple(out.shape) != self.shape or out.dtype != rhs.dtype or out.device != rhs.device or not out.is_contiguous():
            raise ValueError("out must be a contiguous array of the right-hand side's shape, dtype and device")
        outer = int(np.prod(self.shape[:self.poly_axis], dtype=np.int64))
        inner = int(np.prod(self.shape[self.poly_axis + 1:], dtype=np.int64))
        with device_scope(rhs):
            h = self._handle(dt, device_key(rhs))
            L.check(self._lib.jfx_banded_solve(h, C.c_void_p(current_stream_ptr()), C.c_void_p(rhs.data_ptr()),
                                               C.c_void_p(out.data_ptr()), outer, inner))
        return out


def tpmats_wavenumber_factor(terms) -> WavenumberBandedSolver:
    """`tpmats_wavenumber_factor` (la/tpmatrix.py:1236-1354): `terms` = [(scale, [M_0, .., M_{d-1}]), ...] describes
    sum_t scale_t * (M_0 x .. x M_{d-1}); a 1-D array M_a stands for a diagonal matrix.  Axes on which every term is
    diagonal are the Fourier axes; exactly one other (banded) axis is required."""
    if not isinstance(terms, (list, tuple)) or not terms or not all(isinstance(t, (list, tuple)) and len(t) == 2 for t in terms):
        raise TypeError(f"tpmats_wavenumber_factor expects a list of (scale, matrices) terms, got {type(terms).__name__!r}.")
    terms = [(complex(sc) if np.iscomplexobj(sc) else float(sc), [np.asarray(m) for m in mats]) for sc, mats in terms]
    ndim = len(terms[0][1])

    def diagonal(m):
        return m.ndim == 1 or not np.any(m - np.diag(np.diagonal(m)))

    diag_axes = [a for a in range(ndim) if all(diagonal(mats[a]) for _, mats in terms)]
    poly_axes = [a for a in range(ndim) if a not in diag_axes]
    if len(poly_axes) != 1:
        raise SolverNotApplicable(f"tpmats_wavenumber_factor requires exactly 1 polynomial (non-diagonal) axis; found "
                                  f"{len(poly_axes)}: {poly_axes}. Use KroneckerSumSolver for fully polynomial problems.")
    pa = poly_axes[0]
    shape = tuple(int(m.shape[0]) for m in terms[0][1])
    W = []
    for sc, mats in terms:                                   # tpmatrix.py:1306-1316: C order over the Fourier axes
        w = np.array([sc])
        for a in diag_axes:
            dg = mats[a] if mats[a].ndim == 1 else np.diagonal(mats[a])
            w = np.outer(w, dg).ravel()
        W.append(w)
    dias = [dia_from_dense(mats[pa]) for _, mats in terms]
    offsets = tuple(sorted({o for offs, _ in dias for o in offs}))
    P = np.zeros((len(terms), len(offsets), shape[pa]), dtype=np.result_type(*[d.dtype for _, d in dias]))
    for t, (offs, data) in enumerate(dias):                  # tpmatrix.py:1330-1343: aligned to the union of offsets
        for o, row in zip(offs, data):
            P[t, offsets.index(o)] = row
    return WavenumberBandedSolver(pa, shape, np.array(W), P, offsets)


def tpmats_lu_factor(terms) -> KroneckerSumSolver:
    """`tpmats_lu_factor` (la/tpmatrix.py:1089-1233): the diagonalisation solver.  Here it takes a Kronecker SUM
    sum_i (B_0 x .. x A_i x .. x B_{d-1}): one term per axis, and on every axis all terms but one carry the same matrix B
    (what `inner(v * Div(Grad(u)))` and `laplace_terms` produce).  Anything else raises `SolverNotApplicable`."""
    import itertools
    terms = [(float(sc), [np.diag(np.asarray(m)) if np.ndim(m) == 1 else np.asarray(m) for m in mats]) for sc, mats in terms]
    d = len(terms[0][1])
    if len(terms) != d:
        raise SolverNotApplicable(f"a Kronecker sum has one term per axis; got {len(terms)} terms for {d} axes")

    def same(a, b):
        return a.shape == b.shape and np.allclose(a, b, rtol=1e-12, atol=1e-14 * max(1.0, float(np.abs(a).max())))

    for perm in itertools.permutations(range(d)):          # perm[ax] = the term whose matrix on axis ax is the odd one out
        pairs = []
        for ax in range(d):
            t = perm[ax]
            others = [terms[u][1][ax] for u in range(d) if u != t]
            if any(not same(others[0], o) for o in others[1:]):
                break
            B = others[0] if others else np.eye(terms[t][1][ax].shape[0])
            pairs.append((terms[t][0] * terms[t][1][ax], B))                   # the term's scale goes with its A_i
        else:
            return KroneckerSumSolver(pairs)
    raise SolverNotApplicable("the terms are not a Kronecker sum with one common mass matrix per axis")


class TPMatrices:
    """Sum of tensor-product operators  sum_t scale_t (M_0 x .. x M_{d-1})  with the solve interface of the reference's
    `TPMatrices` (la/tpmatrix.py:386-545): `lu_factor()` picks and caches a factored solver — the per-wavenumber banded LU
    for Fourier x polynomial structure, the per-axis diagonalisation for a Kronecker sum — and `solve(rhs)` uses it.
    `terms` = [(scale, [M_0, .., M_{d-1}]), ...]; a 1-D array stands for a diagonal matrix."""

    def __init__(self, terms):
        self.tpmats = [(sc, list(mats)) for sc, mats in terms]
        self._lu_cache = None

    def lu_factor(self):
        if self._lu_cache is None:
            try:
                self._lu_cache = tpmats_wavenumber_factor(self.tpmats)       # dispatch order of la/tpmatrix.py:421-426
            except SolverNotApplicable:
                self._lu_cache = tpmats_lu_factor(self.tpmats)
        return self._lu_cache

    def solve(self, rhs):
        return self.lu_factor().solve(rhs)


def _axis_matrices(space):
    """(stiffness, mass) of one axis; Fourier axes give their diagonals (orthogonal exponentials: (e_l, e_k) = 2 pi / df,
    second derivative -(k df)^2)."""
    if getattr(space, "complex_data", False):
        df = float(space.domain_factor)
        k = np.asarray(space.wavenumbers(), dtype=float)
        m = np.full(space.N, 2 * np.pi / df)
        return -(k * df) ** 2 * m, m
    return stiffness_matrix(space, 2), mass_matrix(space)


def laplace_terms(T, alpha: float = 0.0):
    """Terms of  (v, div grad u)_w + alpha (v, u)_w  on T: one Kronecker product per axis (+ the mass term)."""
    mats = [_axis_matrices(s) for s in T.basespaces]
    terms = [(1.0, [mats[j][0] if j == i else mats[j][1] for j in range(len(mats))]) for i in range(len(mats))]
    if alpha:
        terms.append((float(alpha), [m[1] for m in mats]))
    return terms


def poisson_solver(T):
    """Solver of the weak Laplace operator  (v, div grad u)_w  on the tensor-product space T: per-wavenumber banded LU when
    all axes but one are Fourier (`TPMatrices.lu_factor` dispatch, la/tpmatrix.py:396-427), per-axis diagonalisation otherwise."""
    fourier = [bool(getattr(s, "complex_data", False)) for s in T.basespaces]
    if any(fourier) and fourier.count(False) == 1:
        return tpmats_wavenumber_factor(laplace_terms(T))
    return KroneckerSumSolver([(stiffness_matrix(s, 2), mass_matrix(s)) for s in T.basespaces])
