"""d-dimensional tensor-product spaces — mirror of `jaxfun.galerkin.tensorproductspace`
(`src/jaxfun/galerkin/tensorproductspace.py:38-557`).

The reference loops over the axes and dispatches one `jit(vmap(1-D transform))` per axis
(`tensorproductspace.py:348-365`, `sharding.py:24-40`).  Here the whole separable transform is ONE
engine plan: every axis is described once (`OrthogonalSpace.axis_spec`) and the engine walks the
axes on the device without transposes.  With more than one rank (`torch.distributed` initialised
and a sharded array passed in) the slab algorithm of `sharding.py:43-105` is used: local axes,
all-to-all over NCCL, then the originally sharded axis (see `jaxfun_b200.sharding`).
"""
from __future__ import annotations

import copy
from collections.abc import Sequence

import numpy as np

from .. import _lib as L
from ..engine import Plan, as_jfx_array, jfx_dtype
from .orthogonal import OrthogonalSpace

tensor_product_symbol = "⊗"


class TensorProductSpace:
    is_transient = False

    def __init__(self, basespaces: Sequence[OrthogonalSpace], system=None, name: str = "TPS") -> None:
        self.basespaces = list(basespaces)
        self.name = name
        self.system = system
        self.tensorname = tensor_product_symbol.join([b.name for b in basespaces])
        self._plans: dict = {}

    def __len__(self) -> int:
        return len(self.basespaces)

    def __iter__(self):
        return iter(self.basespaces)

    def __getitem__(self, i: int) -> OrthogonalSpace:
        return self.basespaces[i]

    @property
    def dims(self) -> int:
        return len(self)

    @property
    def is_orthogonal(self) -> bool:
        return all(space.is_orthogonal for space in self.basespaces)

    @property
    def shape(self) -> tuple[int, ...]:
        return tuple(space.num_quad_points for space in self.basespaces)

    @property
    def num_quad_points(self) -> tuple[int, ...]:
        return self.shape

    @property
    def dim(self) -> int:
        return int(np.prod([space.dim for space in self.basespaces], dtype=np.int64))

    @property
    def num_dofs(self) -> tuple[int, ...]:
        return tuple(space.num_dofs for space in self.basespaces)

    @property
    def complex_data(self) -> bool:
        return any(s.complex_data for s in self.basespaces)

    def mesh(self, kind: str = "quadrature", N=None, broadcast: bool = True):
        N = tuple(self.basespaces[ax].num_quad_points if N is None else N[ax] for ax in range(len(self)))
        mesh = []
        for ax, space in enumerate(self.basespaces):
            X = np.asarray(space.mesh(kind, N[ax]))
            mesh.append(self.broadcast_to_ndims(X, ax) if broadcast else X)
        return tuple(mesh)

    def broadcast_to_ndims(self, x, axis: int = 0):
        s = [np.newaxis] * len(self)
        s[axis] = slice(None)
        return x[tuple(s)]

    def get_orthogonal(self) -> "TensorProductSpace":
        return TensorProductSpace([s.get_orthogonal() for s in self.basespaces], system=self.system,
                                  name=self.name + "o")

    # ---- plans ----------------------------------------------------------------------------------
    def _axis_specs(self, op: int, shape, dtype: int, N, k):
        d = len(self)
        lead = len(shape) - d
        assert lead >= 0, f"array rank {len(shape)} < space dimension {d}"
        specs = [None] * lead
        for ax, space in enumerate(self.basespaces):
            specs.append(space.axis_spec(op, shape[lead + ax], dtype,
                                         None if N is None else N[ax], 0 if k is None else k[ax],
                                         inner=int(np.prod(shape[lead + ax + 1:], dtype=np.int64))))
        return specs

    def _plan(self, op: int, x, N=None, k=None) -> Plan:
        dtype = jfx_dtype(x.dtype)
        key = (op, dtype, tuple(x.shape), N, k)
        plan = self._plans.get(key)
        if plan is None:
            plan = Plan(op, dtype, tuple(x.shape), self._axis_specs(op, tuple(x.shape), dtype, N, k))
            self._plans[key] = plan
        return plan

    def _resolve_N(self, N):
        if N is None:
            return None
        return tuple(self.basespaces[ax].num_quad_points if N[ax] is None else int(N[ax]) for ax in range(len(self)))

    # ---- transforms (tensorproductspace.py:330-460) -----------------------------------------
    def backward(self, c, N=None):
        c, _ = as_jfx_array(c, self.complex_data)
        return self._plan(L.OP_BACKWARD, c, N=self._resolve_N(N))(c)

    def forward(self, u):
        u, _ = as_jfx_array(u, self.complex_data)
        return self._plan(L.OP_FORWARD, u)(u)

    def scalar_product(self, u):
        u, _ = as_jfx_array(u, self.complex_data)
        return self._plan(L.OP_SCALAR_PRODUCT, u)(u)

    def backward_primitive(self, c, k, N=None):
        c, _ = as_jfx_array(c, self.complex_data)
        k = tuple(int(v) for v in k)
        if not any(k):
            return self.backward(c, N)
        return self._plan(L.OP_BACKWARD_PRIMITIVE, c, N=self._resolve_N(N), k=k)(c)

    def evaluate_mesh(self, c, kind: str = "quadrature", N=None):
        kind = getattr(kind, "value", kind)
        if kind == "quadrature":
            return self.backward(c, N)
        for ax, space in enumerate(self.basespaces):
            axis = ax - len(self)
            c = space.evaluate_mesh(c, kind, None if N is None else N[ax], axis=axis)
        return c

    def to_orthogonal(self, c):
        for ax, space in enumerate(self.basespaces):
            c = space.to_orthogonal(c) if not hasattr(space, "_to_orthogonal_axis") else space._to_orthogonal_axis(c, ax - len(self))
        return c

    def from_orthogonal(self, c):
        for ax, space in enumerate(self.basespaces):
            c = space.from_orthogonal(c) if not hasattr(space, "_from_orthogonal_axis") else space._from_orthogonal_axis(c, ax - len(self))
        return c


def TensorProduct(*basespaces: OrthogonalSpace, system=None, name: str = "T") -> TensorProductSpace:
    """Factory (tensorproductspace.py:507-557): deep-copies the factor spaces."""
    spaces = []
    for s in basespaces:
        plans, s._plans = s._plans, {}
        try:
            spaces.append(copy.deepcopy(s))
        finally:
            s._plans = plans
    return TensorProductSpace(spaces, system, name)
