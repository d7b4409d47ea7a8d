"""1-D orthogonal function spaces — host-side mirror of `jaxfun.galerkin.orthogonal.OrthogonalSpace`.

Same method surface as the reference class (`src/jaxfun/galerkin/orthogonal.py:37-277`):
`forward`, `scalar_product`, `backward`, `backward_primitive`, `evaluate_mesh`, `evaluate`,
`quad_points_and_weights`, `mesh`, `vandermonde`, `evaluate_basis_derivative`, `norm_squared`,
`derivative_coeffs`, domain mapping.  What differs is where the work happens:

* the constant tables the reference re-traces inside every jitted call (quadrature nodes, weighted
  Vandermonde, norms, derivative recurrences) are built ONCE on the host in float64 numpy and
  handed to the engine when a plan is created;
* the transforms themselves run on the GPU through the C ABI (`include/jfx.h`): an FP64 tensor-core
  contraction for table bases, FFT/DCT kernels for Fourier/Chebyshev.

Arrays may be `torch.cuda` tensors (device path, zero copies) or numpy arrays (host path: the
call copies in, transforms on the GPU and copies out).  Transforms act along `axis` (default: last)
and treat all other axes as batch — the `jax.vmap` of the reference.
"""
from __future__ import annotations

from typing import NamedTuple

import numpy as np

from .. import _lib as L
from ..engine import AxisSpec, Plan, as_jfx_array, device_key, device_scope, fast_path_available, jfx_dtype


class Domain(NamedTuple):
    lower: float
    upper: float


def metric_weight(system, mesh):
    """sqrt(det g) of `system` sampled on `mesh` (a tuple of coordinate arrays), or None when it is 1.

    `system` follows the protocol of `jaxfun.coordinates.CoordSys` as far as the transforms use it
    (orthogonal.py:270-276, tensorproductspace.py:376-379): `.sg` is a number or a SymPy expression in `.base_scalars()`."""
    if system is None:
        return None
    import sympy as sp
    sg = sp.sympify(system.sg)
    if sg == 1:
        return None
    if sg.is_number:
        return complex(sg) if sg.has(sp.I) else float(sg)
    vals = sp.lambdify(tuple(system.base_scalars()), sg, modules="numpy")(*mesh)
    return np.asarray(vals)


class OrthogonalSpace:
    is_orthogonal = True
    #: engine basis used when a fast kernel exists for the transform length (else dense table)
    fast_basis = L.BASIS_NONE
    #: Fourier spaces need complex arrays
    complex_data = False

    def __init__(self, N: int, *, domain=None, system=None, name: str = "OrthogonalSpace",
                 fun_str: str = "psi", **kw) -> None:
        self.N = int(N)
        self._num_quad_points = int(N)
        if domain is None:
            domain = self.reference_domain
        self._domain = Domain(*domain)
        self.name = name
        self.fun_str = fun_str
        self.system = system
        self.bcs = None
        self.orthogonal = self
        self.stencil = {0: 1}
        self._plans: dict = {}
        self._tables: dict = {}

    # ---- abstract host tables -------------------------------------------------------------
    @property
    def reference_domain(self) -> Domain:
        raise NotImplementedError

    def quad_points_and_weights(self, N: int | None = None):
        raise NotImplementedError

    def eval_basis_functions(self, X):
        """[len(X), N] array of psi_k(X_j) (reference coordinates)."""
        raise NotImplementedError

    def norm_squared(self) -> np.ndarray:
        raise NotImplementedError

    def _derivative_host(self, c: np.ndarray) -> np.ndarray:
        """First-derivative coefficient map along axis 0 of a [n, m] array (host)."""
        raise NotImplementedError

    # ---- basic properties ----------------------------------------------------------------
    @property
    def num_quad_points(self) -> int:
        return self._num_quad_points

    @property
    def shape(self):
        return (self.num_quad_points,)

    @property
    def dim(self) -> int:
        return self.N

    @property
    def dims(self) -> int:
        return 1

    @property
    def num_dofs(self) -> int:
        return self.dim

    @property
    def domain(self) -> Domain:
        return self._domain

    def __len__(self) -> int:
        return 1

    @property
    def domain_factor(self) -> float:
        a, b = (float(v) for v in self.domain)
        c, d = (float(v) for v in self.reference_domain)
        Lt, R = b - a, d - c
        return R / Lt if abs(Lt - R) > 1e-12 else 1

    def map_reference_domain(self, x):
        if tuple(float(v) for v in self.domain) != tuple(float(v) for v in self.reference_domain):
            a = float(self.domain.lower)
            c = float(self.reference_domain.lower)
            return c + (np.asarray(x) - a) * float(self.domain_factor)
        return x

    def map_true_domain(self, X):
        if tuple(float(v) for v in self.domain) != tuple(float(v) for v in self.reference_domain):
            a = float(self.domain.lower)
            c = float(self.reference_domain.lower)
            return a + (np.asarray(X) - c) / float(self.domain_factor)
        return X

    def mesh(self, kind: str = "quadrature", N: int | None = None) -> np.ndarray:
        N = self.num_quad_points if N is None else N
        kind = getattr(kind, "value", kind)
        if kind == "quadrature":
            return self.map_true_domain(self.quad_points_and_weights(N)[0])
        assert kind == "uniform", f"Unsupported mesh kind: {kind}"
        a, b = self.domain
        return np.linspace(float(a), float(b), N)

    def get_orthogonal(self):
        return self

    def to_orthogonal(self, c):
        return c

    def from_orthogonal(self, c):
        return c

    # ---- host tables derived from the abstract ones -----------------------------------------
    def vandermonde(self, X) -> np.ndarray:
        return self.evaluate_basis_derivative(X, 0)

    def derivative_matrix(self, k: int = 1, n: int | None = None) -> np.ndarray:
        """Dense [n, n] matrix D with derivative_coeffs(c, k) == D @ c for len(c) == n."""
        n = self.N if n is None else n
        key = ("D", k, n)
        D = self._tables.get(key)
        if D is None:
            D = np.eye(n, dtype=self._table_dtype())
            for _ in range(k):
                D = self._derivative_host(D)
            self._tables[key] = D
        return D

    def _table_dtype(self):
        return np.float64

    def evaluate_basis_derivative(self, X, k: int = 0) -> np.ndarray:
        """[len(X), N] Vandermonde of the k-th derivative of the basis (reference coordinate)."""
        V = self.eval_basis_functions(np.atleast_1d(np.asarray(X, dtype=float)))
        if k == 0:
            return V
        return V @ self.derivative_matrix(k)

    def mass_diagonal(self) -> np.ndarray:
        return self.norm_squared() / float(self.domain_factor)

    # ---- per-axis engine specs ---------------------------------------------------------------
    def _weights_scaled(self, n: int):
        """Quadrature weights times sg / df (orthogonal.py:268-276): `system.sg` is the metric weight of a curvilinear 1-D
        coordinate (a number, or a SymPy expression in the true coordinate sampled at the true-domain nodes); None = 1."""
        xj, wj = self.quad_points_and_weights(n)
        wj = wj * (1.0 / float(self.domain_factor))
        sg = metric_weight(self.system, (np.asarray(self.map_true_domain(xj), dtype=float),))
        return xj, (wj if sg is None else wj * sg)

    def _dense_table(self, op: int, n_coeff: int, n_quad: int, deriv: int) -> np.ndarray:
        """Dense [n_out, n_in] table of `op` along one axis (generic Vandermonde definition)."""
        key = ("T", op, n_coeff, n_quad, deriv)
        T = self._tables.get(key)
        if T is not None:
            return T
        if op in (L.OP_FORWARD, L.OP_SCALAR_PRODUCT):
            xj, wj = self._weights_scaled(n_quad)
            Pi = self.vandermonde(xj)                      # [n, N]
            T = np.conj(Pi).T * wj[None, :]                # scalar_product: (u*w) @ conj(Pi)
            if op == L.OP_FORWARD:
                T = T / self.mass_diagonal()[:, None]
        else:
            # the reference evaluates at mesh() (true domain) and maps back (orthogonal.py:226-227,
            # 113-115); keep that round trip so the nodes carry the same rounding
            xj = self.map_reference_domain(self.mesh("quadrature", n_quad))
            # backward: sum_k c_k psi_k(x_j).  Spaces whose reference backward is a series recurrence of its own
            # (Jacobi._evaluate) supply that recurrence's table; the others the Vandermonde (orthogonal.py:117-129)
            T = self.series_table(xj, n_coeff) if hasattr(self, "series_table") else self.vandermonde(xj)[:, :n_coeff]
            if deriv:
                T = (float(self.domain_factor) ** deriv) * (T @ self.derivative_matrix(deriv, n_coeff))
        T = np.ascontiguousarray(T)
        self._tables[key] = T
        return T

    def axis_spec(self, op: int, n_in: int, dtype: int, N: int | None = None, k: int = 0,
                  inner: int = 1) -> AxisSpec:
        """Engine description of this space's transform `op` along one axis of extent n_in."""
        if op in (L.OP_FORWARD, L.OP_SCALAR_PRODUCT):
            n_quad, n_coeff = n_in, self.N
            assert n_quad >= self.N, "Only truncation supported for forward transform"
        else:
            n_quad = self.num_quad_points if N is None else int(N)
            n_coeff = n_in
            assert n_coeff <= self.N, f"Coefficient length {n_coeff} exceeds N={self.N}"
            assert n_quad >= n_coeff, "backward only supports padding, not truncation"
        fast_ok = self.fast_basis != L.BASIS_NONE and fast_path_available(self.fast_basis, n_quad, dtype)
        if self.fast_basis == L.BASIS_CHEBYSHEV and (k != 0 or (dtype in (L.F32, L.F64) and inner > 1 and inner % 2)):
            fast_ok = False   # chebder^k is folded into a dense table; odd real inner extents have no DCT tile
        if fast_ok:
            return AxisSpec(self.fast_basis, n_modes=(self.N if op in (L.OP_FORWARD, L.OP_SCALAR_PRODUCT) else n_coeff),
                            n_quad=n_quad, deriv=k, domain_factor=float(self.domain_factor))
        T = self._dense_table(op, n_coeff, n_quad, k)
        basis = L.BASIS_CTABLE if np.iscomplexobj(T) else L.BASIS_TABLE
        return AxisSpec(basis, n_modes=n_coeff, n_quad=n_quad, deriv=k,
                        domain_factor=float(self.domain_factor), table=T)

    # ---- device transforms ------------------------------------------------------------------------
    def __getstate__(self):
        # engine plans hold ctypes handles: copies (TensorProduct deep-copies its factors, incl. the nested orthogonal space of
        # a Composite) start without them and build their own
        d = dict(self.__dict__)
        d["_plans"] = {}
        return d

    def _run(self, op: int, x, axis: int, N=None, k: int = 0, table: np.ndarray | None = None,
             cache: bool = True, name=None):
        """`table`: explicit [n_out, n_in] table of an APPLY plan; cached plans of such tables are keyed by `name` (a stable
        label chosen by the caller), never by the identity of a temporary array."""
        x, _ = as_jfx_array(x, self.complex_data)
        axis = axis % x.ndim
        dtype = jfx_dtype(x.dtype)
        if table is not None and name is None:
            cache = False
        key = (op, dtype, tuple(x.shape), axis, N, k, name, device_key(x))      # one plan per device
        plan = self._plans.get(key) if cache else None
        with device_scope(x):
            if plan is None:
                if table is not None:
                    spec = AxisSpec(L.BASIS_CTABLE if np.iscomplexobj(table) else L.BASIS_TABLE, table=table)
                else:
                    spec = self.axis_spec(op, x.shape[axis], dtype, N, k,
                                          inner=int(np.prod(x.shape[axis + 1:], dtype=np.int64)))
                axes = [None] * x.ndim
                axes[axis] = spec
                plan = Plan(op, dtype, tuple(x.shape), axes)
                if cache:
                    self._plans[key] = plan
            return plan(x)

    def forward(self, u, axis: int = -1):
        """Samples at quadrature points -> expansion coefficients (orthogonal.py:256-262)."""
        return self._run(L.OP_FORWARD, u, axis)

    def scalar_product(self, u, axis: int = -1):
        """Weighted inner products <u, psi_k> (orthogonal.py:264-277)."""
        return self._run(L.OP_SCALAR_PRODUCT, u, axis)

    def backward(self, c, N: int | None = None, axis: int = -1):
        """Series evaluated at the N quadrature points (orthogonal.py:214-227)."""
        return self._run(L.OP_BACKWARD, c, axis, N=N)

    def backward_primitive(self, c, k: int = 0, N: int | None = None, axis: int = -1):
        """d^k u / dx^k at the quadrature points (orthogonal.py:229-246)."""
        if k == 0:
            return self.backward(c, N, axis)
        return self._run(L.OP_BACKWARD_PRIMITIVE, c, axis, N=N, k=k)

    def derivative_coeffs(self, c, k: int = 0, axis: int = -1):
        """Coefficients of the k-th derivative series (same length as c)."""
        if k == 0:
            return c
        n = c.shape[axis]
        return self._run(L.OP_APPLY, c, axis, k=k, table=self.derivative_matrix(k, n), name=("D", k, n))

    def evaluate(self, x, c, axis: int = -1):
        """sum_k c_k psi_k(x) at arbitrary true-domain points x (orthogonal.py:102-115)."""
        X = np.atleast_1d(np.asarray(self.map_reference_domain(np.asarray(x, dtype=float))))
        n = c.shape[axis]
        assert n <= self.N, f"Coefficient length {n} exceeds N={self.N}"
        # the reference evaluates through _evaluate (orthogonal.py:102-115): Jacobi-type spaces have their own series recurrence
        T = self.series_table(X, n) if hasattr(self, "series_table") and type(self)._dense_table is OrthogonalSpace._dense_table \
            else self.eval_basis_functions(X)[:, :n]
        return self._run(L.OP_APPLY, c, axis, table=np.ascontiguousarray(T), cache=False)

    def evaluate_mesh(self, c, kind: str = "quadrature", N: int | None = None, axis: int = -1):
        kind = getattr(kind, "value", kind)
        if kind == "quadrature":
            return self.backward(c, N, axis)
        assert kind == "uniform", f"Unsupported mesh kind: {kind}"
        return self.evaluate(self.mesh(kind=kind, N=N), c, axis)
