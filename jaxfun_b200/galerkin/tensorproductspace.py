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
from ..engine import Plan, as_jfx_array, device_key, device_scope, jfx_dtype
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
        key = (op, dtype, tuple(x.shape), N, k, device_key(x))                  # one plan per device
        plan = self._plans.get(key)
        if plan is None:
            with device_scope(x):
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

    def _metric(self, like):
        """sqrt(det g) sampled on the quadrature mesh, as an array like `like` (None for Cartesian systems)."""
        from .orthogonal import metric_weight
        if self.system is None:
            return None
        key = ("sg", getattr(like, "device", "host"), str(like.dtype))
        if key not in self._plans:
            sg = metric_weight(self.system, self.mesh())
            if sg is not None and not np.isscalar(sg):
                sg = np.ascontiguousarray(np.broadcast_to(sg, self.shape))
                if not isinstance(like, np.ndarray):
                    import torch
                    sg = torch.from_numpy(sg).to(device=like.device, dtype=like.dtype)
            self._plans[key] = sg
        return self._plans[key]

    def scalar_product(self, u):
        u, _ = as_jfx_array(u, self.complex_data)
        sg = self._metric(u)                     # tensorproductspace.py:376-379: u * sg before the separable products
        if sg is not None:
            u = u * sg
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

    def evaluate(self, x, c, group=None):
        """Expansion evaluated at scattered points (tensorproductspace.py:263-321): x is [n_pts, d] in true coordinates,
        result [n_pts] = einsum("pi,pj(,pk),ij(k)->p", C_0, C_1, (C_2), c) with C_ax = basis values of axis ax at the points.

        The last axis is a table pass of the engine ([.., N_last] -> [.., n_pts]); every other axis is reduced by
        `jfx_point_contract` with its per-point basis values.  With `group` (torch.distributed) the coefficient block `c` is
        this rank's slab of a spectral array sharded along axis 0; the partial sums are all-reduced (the `psum` of :300-304)."""
        offset = 0
        if group is not None:                                   # this rank's modes of axis 0
            import torch.distributed as dist
            offset = dist.get_rank(group) * c.shape[0]
        y = self._evaluate_partial(x, c, offset)
        if group is not None:
            dist.all_reduce(y, group=group)
        return y

    def _evaluate_partial(self, x, c, mode_offset: int = 0):
        """Contribution of the coefficient block c = C[mode_offset : mode_offset + c.shape[0]] (axis 0) to evaluate(x, C)."""
        import ctypes as C
        import torch
        from ..engine import current_stream_ptr
        d = len(self)
        x = np.atleast_2d(np.asarray(x, dtype=float))
        assert x.shape[1] == d, f"points must be [n_pts, {d}]"
        c, is_host = as_jfx_array(c, self.complex_data)
        if is_host:
            raise L.JfxError(-3, "evaluate needs a CUDA tensor; jaxfun_b200 has no CPU fallback")
        npts = x.shape[0]
        Cs = []
        for ax, space in enumerate(self.basespaces):
            X = np.asarray(space.map_reference_domain(x[:, ax]), dtype=float)
            Cs.append(np.asarray(space.eval_basis_functions(X)))          # the Vandermonde, as the reference (:267-270)
        Cs[0] = Cs[0][:, mode_offset:]
        for ax in range(d):
            assert c.shape[ax] <= Cs[ax].shape[1], f"axis {ax}: {c.shape[ax]} coefficients exceed N"
            Cs[ax] = np.ascontiguousarray(Cs[ax][:, :c.shape[ax]])
        # last axis: [.., N_last] x C_last^T -> [.., n_pts]
        y = self.basespaces[-1]._run(L.OP_APPLY, c, -1, table=Cs[-1], cache=False)
        lib = L.load()
        dtype = jfx_dtype(y.dtype)
        wdt = torch.float64 if dtype in (L.F64, L.C128) else torch.float32
        for ax in range(d - 2, -1, -1):
            w = Cs[ax]
            wc = bool(np.iscomplexobj(w))                      # Fourier axis: complex basis values, complex data
            wt = torch.from_numpy(w).to(device=y.device, dtype=(y.dtype if wc else wdt)).contiguous()
            outer = int(np.prod(y.shape[:ax], dtype=np.int64))
            out = torch.empty(tuple(y.shape[:ax]) + (npts,), dtype=y.dtype, device=y.device)
            L.check(lib.jfx_point_contract(C.c_void_p(current_stream_ptr()), C.c_void_p(y.data_ptr()), C.c_void_p(wt.data_ptr()),
                                           C.c_void_p(out.data_ptr()), outer, int(y.shape[ax]), npts, dtype, int(wc)))
            y = out
        return y

    def to_orthogonal(self, c):
        for ax, space in enumerate(self.basespaces):
            c = space.to_orthogonal(c) if not hasattr(space, "_to_orthogonal_axis") else space._to_orthogonal_axis(c, ax - len(self))
        return c

    def from_orthogonal(self, c):
        for ax, space in enumerate(self.basespaces):
            c = space.from_orthogonal(c) if not hasattr(space, "_from_orthogonal_axis") else space._from_orthogonal_axis(c, ax - len(self))
        return c


def _sample_boundary_value(v, symbols, grids, shape):
    """Boundary datum on the quadrature mesh of the other axes: number | sympy expression in x, y, z | callable | samples."""
    import sympy as sp
    if callable(v) and not isinstance(v, sp.Basic):
        return np.broadcast_to(np.asarray(v(*grids)), shape).copy()
    if isinstance(v, np.ndarray):
        assert tuple(v.shape) == tuple(shape), f"boundary samples have shape {v.shape}, the mesh of the other axes {shape}"
        return v
    e = sp.sympify(v)
    if not e.free_symbols:
        return np.full(shape, complex(e) if e.has(sp.I) else float(e))
    by_name = {str(sym): sym for sym in e.free_symbols}
    unknown = set(by_name) - set(symbols)
    assert not unknown, f"boundary value {e} depends on {sorted(unknown)}; the other axes are {symbols}"
    f = sp.lambdify([by_name.get(n, sp.Symbol(n)) for n in symbols], e, modules="numpy")
    return np.broadcast_to(np.asarray(f(*grids)), shape).copy()


class DirectSumTPS:
    """Tensor product with ONE or TWO DirectSum factors (2-D / 3-D; two: `_init_two_directions`) whose boundary values may depend on the other coordinates
    (`DirectSumTPS`, tensorproductspace.py:575-851; the case of `examples/poisson2D_periodic.py`).

    The lift is a fixed element of the orthogonal tensor-product space: every boundary datum g_b(other coordinates) is
    projected onto the other factors (`project1D` = forward transform of its samples, inner.py:1027-1046) and multiplied
    by the lifting function B_b along the DirectSum axis (tensorproductspace.py:817-830).  Its coefficients `lift` are
    built once on the host; transforms are the homogeneous tensor product's engine plans plus that constant:
    `backward(c) = hom.backward(c) + orthogonal.backward(lift)` (cached), `forward(u) = hom.from_orthogonal(
    orthogonal.forward(u) - lift)`."""

    def __init__(self, basespaces, system=None, name: str = "DSTPS") -> None:
        from .composite import DirectSum
        idx = [i for i, s in enumerate(basespaces) if isinstance(s, DirectSum)]
        if len(idx) == 2 and (len(basespaces) == 2 or (len(basespaces) == 3 and idx == [1, 2])):
            self._init_two_directions(list(basespaces), system, name)
            return
        if len(idx) == 2 and len(basespaces) == 3:
            raise ValueError("DirectSum cannot be the first space in a 3D tensor product.")   # tensorproductspace.py:612-615
        if len(idx) != 1:
            raise NotImplementedError("DirectSumTPS: one or two inhomogeneous directions are built")
        b, d = idx[0], len(basespaces)
        if d == 3 and b == 0:
            raise ValueError("DirectSum cannot be the first space in a 3D tensor product.")   # tensorproductspace.py:612-615
        D = basespaces[b]
        self.bc_axis, self.name, self.system = b, name, system
        self.basespaces = list(basespaces)
        self.hom = TensorProduct(*[s.a if i == b else s for i, s in enumerate(basespaces)], system=system, name=name + "0")
        self.orthogonal = TensorProduct(*[s.orthogonal if i == b else s.get_orthogonal() for i, s in enumerate(basespaces)],
                                        system=system, name=name + "o")
        others = [s for i, s in enumerate(self.hom.basespaces) if i != b]
        names = ["x", "y", "z"][:d]
        symbols = [n for i, n in enumerate(names) if i != b]
        grids = np.meshgrid(*[np.asarray(s.mesh(), dtype=float) for s in others], indexing="ij")
        shape = tuple(g.shape for g in grids)[0]
        lift = None
        for j, v in enumerate(D.raw_vals):
            gh = _sample_boundary_value(v, symbols, grids, shape)
            for ax, s in enumerate(others):                    # project1D along every other axis, then to orthogonal
                T = np.asarray(s._dense_table(L.OP_FORWARD, s.dim, s.num_quad_points, 0))
                gh = np.moveaxis(np.tensordot(T, gh, axes=(1, ax)), 0, ax)
                if not s.is_orthogonal:
                    gh = np.moveaxis(np.tensordot(np.asarray(s.S).T, gh, axes=(1, ax)), 0, ax)
            row = np.zeros(D.N)
            row[:D.S_bc.shape[1]] = D.S_bc[j]
            term = np.expand_dims(gh, b) * row.reshape([-1 if i == b else 1 for i in range(d)])
            lift = term if lift is None else lift + term
        if not self.orthogonal.complex_data:
            assert np.abs(np.imag(lift)).max() == 0 if np.iscomplexobj(lift) else True
            lift = np.real(lift)
        self.lift = np.ascontiguousarray(lift)
        self._cache: dict = {}

    def _init_two_directions(self, basespaces, system, name) -> None:
        """Two factors carry inhomogeneous boundary values (tensorproductspace.py:620-668, 689-747): a 2-D product of two
        DirectSums, or a 3-D product whose LAST two factors are DirectSums behind a plain first factor (the reference refuses a
        DirectSum in front, :612-615).  Writing (y, z) for the two inhomogeneous directions and x for the optional first one,
        the lift is the transfinite interpolant of the boundary data, point by point in x,

            sum_b B^y_b(y) [hom. part of g_b](x, z) + sum_c [hom. part of h_c](x, y) B^z_c(z) + sum_bc C_bc(x) B^y_b(y) B^z_c(z),

        where "hom. part" is the projection onto the OTHER direction's composite space after removing that direction's own
        lift of the corner values C_bc = (boundary functional c of z)(g_b) — the `projected_bcs` of the reference (Dirichlet:
        the value at the corner, Neumann: the derivative / df^nd) — and, in 3-D, everything is finally projected onto the first
        factor (`project` onto the product of the other spaces, :704-737, 739-747).  Boundary data must be numbers or SymPy
        expressions in the coordinates of the other axes (named x, y[, z] by position), as in the reference."""
        import sympy as sp
        d = len(basespaces)
        lead = d - 2                                           # 0: (D, D);  1: (plain, D, D)
        S0 = basespaces[0] if lead else None
        Dx, Dy = basespaces[lead], basespaces[lead + 1]
        self.bc_axis, self.name, self.system = (lead, lead + 1), name, system
        self.basespaces = list(basespaces)
        self.hom = TensorProduct(*([S0] if lead else []), Dx.a, Dy.a, system=system, name=name + "0")
        self.orthogonal = TensorProduct(*([S0.get_orthogonal()] if lead else []), Dx.orthogonal, Dy.orthogonal, system=system,
                                        name=name + "o")
        names_xyz = ["x", "y", "z"][:d]
        sym = {0: sp.Symbol(names_xyz[lead], real=True), 1: sp.Symbol(names_xyz[lead + 1], real=True)}
        sym_lead = sp.Symbol(names_xyz[0], real=True) if lead else None
        from .composite import _BC_ORDER, ordered_bc_names
        # samples of the first factor's coordinate: everything below carries a leading axis of that length (1 in 2-D)
        S0h = self.hom.basespaces[0] if lead else None
        xq = np.asarray(S0h.mesh(), dtype=float) if lead else np.zeros(1)
        nq0 = xq.shape[0]

        def as_expr(v, other):
            e = sp.sympify(v)
            allowed = {str(sym[other])} | ({str(sym_lead)} if lead else set())
            if lead:
                extra = {str(f) for f in e.free_symbols} - allowed
                assert not extra, f"boundary value {e} may only depend on {sorted(allowed)}"
                return e.xreplace({f: (sym_lead if str(f) == str(sym_lead) else sym[other]) for f in e.free_symbols})
            extra = {str(f) for f in e.free_symbols} - allowed
            assert not extra, f"boundary value {e} may only depend on {sym[other]}"
            return e.xreplace({f: sym[other] for f in e.free_symbols})

        def on_lead(e):
            """Samples over the first factor's mesh of an expression in its coordinate (a constant in 2-D)."""
            e = sp.sympify(e)
            if not e.free_symbols:
                return np.full(nq0, complex(e) if e.has(sp.I) else float(e))
            return np.broadcast_to(np.asarray(sp.lambdify(sym_lead, e, modules="numpy")(xq)), (nq0,)).copy()

        def functional(D, side, kind, expr, var):
            """Boundary functional (side, kind) of direction D applied to expr(var): value, or derivative / df^nd — sampled
            over the first factor's mesh."""
            a, b_ = (float(v) for v in D.a.domain)
            z = a if side == "left" else b_
            nd = _BC_ORDER[kind]
            if kind not in ("D",) and kind[0] != "N":
                raise NotImplementedError("two inhomogeneous directions: Dirichlet / Neumann conditions")
            df = 2.0 / (b_ - a)
            return on_lead((expr.diff(var, nd) / df**nd if nd else expr).subs(var, z))

        Ds, names = {0: Dx, 1: Dy}, {ax: ordered_bc_names(Ds_.bcs) for ax, Ds_ in ((0, Dx), (1, Dy))}
        data = {ax: [as_expr(v, 1 - ax) for v in Ds[ax].raw_vals] for ax in (0, 1)}
        # corner values from the first direction's data, as the reference (projected_bcs[0]): C[b][c] = functional_c (g_b)
        C = np.array([[functional(Dy, sc, kc, g, sym[1]) for (sc, kc) in names[1]] for g in data[0]])      # [nb, nc, nq0]
        # consistency of the data at the corners: the second direction's data must give the same numbers
        C2 = np.array([[functional(Dx, sb, kb, h, sym[0]) for h in data[1]] for (sb, kb) in names[0]])
        if not np.allclose(C, C2, rtol=1e-9, atol=1e-11):
            raise ValueError(f"boundary data of the two directions disagree at the corners:\n{C}\nvs\n{C2}")
        rows = {}
        for ax, D in Ds.items():
            R = np.zeros((D.S_bc.shape[0], D.N))
            R[:, :D.S_bc.shape[1]] = D.S_bc
            rows[ax] = R                                          # lifting functions of direction ax in orthogonal coefficients
        lift = np.einsum("bcq,bi,cj->qij", C, rows[0], rows[1])    # corner block, per sample of the first coordinate
        for ax in (0, 1):
            o = 1 - ax
            Do = Ds[o]
            orth_o, comp_o = Do.orthogonal, Do.a
            xo = np.asarray(comp_o.mesh(), dtype=float)
            Vo = np.asarray(orth_o.eval_basis_functions(np.asarray(orth_o.map_reference_domain(xo), dtype=float)))
            Tf = np.asarray(comp_o._dense_table(L.OP_FORWARD, comp_o.dim, comp_o.num_quad_points, 0))
            St = np.asarray(comp_o.S).T
            for b, g in enumerate(data[ax]):
                if lead:
                    f = sp.lambdify((sym_lead, sym[o]), g, modules="numpy")
                    vals = f(xq[:, None], xo[None, :])
                else:
                    vals = sp.lambdify(sym[o], g, modules="numpy")(xo)[None, :] if g.free_symbols else complex(g) if g.has(sp.I) else float(g)
                samples = np.broadcast_to(np.asarray(vals, dtype=complex if g.has(sp.I) else float), (nq0, xo.shape[0])).copy()
                corner = C[b] if ax == 0 else C[:, b]              # [n_other_bcs, nq0]
                samples = samples - np.einsum("cq,cj,pj->qp", corner, rows[o], Vo)   # minus the other direction's lift of the corners
                orth_coeffs = np.einsum("jm,mp,qp->qj", St, Tf, samples)             # homogeneous part, orthogonal coefficients
                lift = lift + (np.einsum("i,qj->qij", rows[0][b], orth_coeffs) if ax == 0
                               else np.einsum("qi,j->qij", orth_coeffs, rows[1][b]))
        if lead:
            # project1D along the first factor (forward transform of the samples), then to its orthogonal coefficients
            T0 = np.asarray(S0h._dense_table(L.OP_FORWARD, S0h.dim, S0h.num_quad_points, 0))
            lift = np.tensordot(T0, lift, axes=(1, 0))
            if not S0h.is_orthogonal:
                lift = np.tensordot(np.asarray(S0h.S).T, lift, axes=(1, 0))
        else:
            lift = lift[0]
        if not self.orthogonal.complex_data and np.iscomplexobj(lift):
            assert np.abs(lift.imag).max() < 1e-14 * max(1.0, np.abs(lift).max())
            lift = lift.real
        self.lift = np.ascontiguousarray(lift)
        self.corner_values = C if lead else C[:, :, 0]
        self._cache = {}

    # ---- bookkeeping forwarded to the homogeneous product --------------------------------------------------------
    def __len__(self) -> int:
        return len(self.basespaces)
    @property
    def dims(self) -> int:
        return len(self.basespaces)
    @property
    def shape(self):
        return self.hom.shape
    @property
    def num_dofs(self):
        return self.hom.num_dofs
    @property
    def dim(self) -> int:
        return self.hom.dim
    def mesh(self, kind: str = "quadrature", N=None, broadcast: bool = True):
        return self.hom.mesh(kind, N, broadcast)
    def get_homogeneous(self):
        return self.hom
    def get_orthogonal(self):
        return self.orthogonal

    # ---- the constant lift ----------------------------------------------------------------------------------------------
    def _lift_like(self, x):
        if isinstance(x, np.ndarray):
            return self.lift.astype(x.dtype) if np.iscomplexobj(x) or not np.iscomplexobj(self.lift) else self.lift
        import torch
        key = ("lift", x.device, x.dtype)
        t = self._cache.get(key)
        if t is None:
            t = self._cache[key] = torch.from_numpy(self.lift).to(device=x.device, dtype=x.dtype)
        return t

    def _physical_lift(self, like, k, N):
        key = ("phys", getattr(like, "device", "host"), like.dtype, k, N)
        u = self._cache.get(key)
        if u is None:
            lc = self._lift_like(like)
            u = self.orthogonal.backward(lc, N) if k is None else self.orthogonal.backward_primitive(lc, k, N)
            self._cache[key] = u
        return u

    def _coeff_dtype(self, c):
        c, _ = as_jfx_array(c, self.orthogonal.complex_data)
        return c

    # ---- transforms (tensorproductspace.py:781-851) ----------------------------------------------------------------------
    def to_orthogonal(self, c):
        c = self._coeff_dtype(c)
        return self.hom.to_orthogonal(c) + self._lift_like(c)

    def from_orthogonal(self, a):
        a = self._coeff_dtype(a)
        return self.hom.from_orthogonal(a - self._lift_like(a))

    def backward(self, c, N=None):
        c = self._coeff_dtype(c)
        N = None if N is None else tuple(N)
        return self.hom.backward(c, N) + self._physical_lift(c, None, N)

    def backward_primitive(self, c, k, N=None):
        c = self._coeff_dtype(c)
        k, N = tuple(int(v) for v in k), None if N is None else tuple(N)
        return self.hom.backward_primitive(c, k, N) + self._physical_lift(c, k, N)

    def forward(self, u):
        return self.from_orthogonal(self.orthogonal.forward(u))

    def scalar_product(self, u):
        raise RuntimeError("Scalar product requires homogeneous test space (call on get_homogeneous())")

    def evaluate_mesh(self, c, kind: str = "quadrature", N=None):
        return self.orthogonal.evaluate_mesh(self.to_orthogonal(c), kind, N)

    def evaluate(self, x, c):
        """Expansion (homogeneous part + boundary lift) at scattered points x [n_pts, d] — `DirectSumTPS.evaluate`,
        tests/galerkin/test_tensorproductspace_more.py:83-99: the orthogonal product evaluates the lifted coefficients."""
        return self.orthogonal.evaluate(x, self.to_orthogonal(c))


def TensorProduct(*basespaces: OrthogonalSpace, system=None, name: str = "T") -> TensorProductSpace:
    """Factory (tensorproductspace.py:507-557): deep-copies the factor spaces; a DirectSum factor makes it a DirectSumTPS."""
    from .composite import DirectSum
    if any(isinstance(s, DirectSum) for s in basespaces):
        return DirectSumTPS(list(basespaces), system, name)
    spaces = []
    for s in basespaces:
        plans, s._plans = s._plans, {}
        try:
            spaces.append(copy.deepcopy(s))
        finally:
            s._plans = plans
    if system is not None:
        for s in spaces:
            s.system = None       # a factor sees the 1-D sub-system, whose sg is 1 (coordinates.py:1254); the product applies sg
    return TensorProductSpace(spaces, system, name)
