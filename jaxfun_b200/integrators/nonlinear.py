"""Pseudo-spectral nonlinear terms — the engine-side counterpart of
`jaxfun.integrators.nonlinear.NonlinearCompiler` (`src/jaxfun/integrators/nonlinear.py:34-298`) and
`BaseIntegrator.nonlinear_rhs / nonlinear_rhs_scalar_product` (`integrators/base.py:230-248`).

The reference walks a SymPy tree every call: leaves call `space.backward` /
`space.backward_primitive`, interior nodes are pointwise Add / Mul / Pow / Function, and the
result goes through `testspace.forward` (or `scalar_product`).  Here the same tree is compiled ONCE
into

    leaves   : a list of derivative-order tuples            -> backward_primitive plans
    program  : postfix bytecode over the leaf values        -> one pointwise kernel
    final    : forward | scalar_product plan

and handed to `jfx_nonlinear_create`; one call = one `jfx_nonlinear_execute` on the device.

Expressions are written with the space's field symbol, e.g. for KdV on a Fourier space V:

    u, (x,) = field(V)                    # u = Function('u')(x)
    N = NonlinearTerm(V, -u * u.diff(x))  # the integrator stores -(rhs nonlinear part), base.py:165-166

Derivatives of products are expanded symbolically first (nonlinear.py:57-61), repeated
sub-expressions are evaluated once per point (the reference memoises per node, nonlinear.py:244-254),
and coordinate-only factors become mesh-sampled static arrays (nonlinear.py:219-242).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import sympy as sp

from .. import _lib as L
from ..engine import _fill_plan_desc, current_stream_ptr, jfx_dtype, require_device

try:
    import torch
except Exception:  # pragma: no cover
    torch = None

_COORDS = sp.symbols("x y z", real=True)


def field(space, name: str = "u"):
    """Return (u(x[,y[,z]]), coordinates) — the symbolic stand-in for the trial field on `space`."""
    d = space.dims
    xs = _COORDS[:d]
    return sp.Function(name)(*xs), xs


class CompiledExpression:
    """SymPy expression -> (leaves, postfix program, constants, statics)."""

    def __init__(self, expr, u, coords):
        self.u, self.coords = u, tuple(coords)
        self.leaves: list[tuple[int, ...]] = []
        self.program: list[tuple[int, int]] = []
        self.consts: list[complex] = []
        self.statics: list[sp.Expr] = []
        expr = sp.expand(sp.sympify(expr).doit())
        self._emit(expr)
        if len(self.program) > L.JFX_MAX_PROGRAM:
            raise ValueError(f"nonlinear expression too long ({len(self.program)} > {L.JFX_MAX_PROGRAM} ops)")
        if len(self.leaves) > L.JFX_MAX_LEAVES:
            raise ValueError(f"too many distinct derivative leaves ({len(self.leaves)} > {L.JFX_MAX_LEAVES})")

    # -- helpers ------------------------------------------------------------------------------
    def _leaf(self, orders: tuple[int, ...]) -> int:
        if orders not in self.leaves:
            self.leaves.append(orders)
        return self.leaves.index(orders)

    def _const(self, value: complex) -> int:
        value = complex(value)
        if value not in self.consts:
            self.consts.append(value)
        if len(self.consts) > 32:
            raise ValueError("too many distinct constants in nonlinear expression")
        return self.consts.index(value)

    def _push(self, op: int, arg: int = 0) -> None:
        self.program.append((op, int(arg)))

    def _contains_field(self, node) -> bool:
        return node.has(self.u.func)

    def _emit(self, node) -> None:
        node = sp.sympify(node)
        zero = (0,) * len(self.coords)
        if node == self.u:
            self._push(L.PW_LEAF, self._leaf(zero))
        elif isinstance(node, sp.Derivative):
            if node.expr == self.u:
                unknown = [v for v in node.variables if v not in self.coords]
                if unknown:
                    raise ValueError(f"Only spatial derivatives are supported, got: {unknown}")
                orders = tuple(node.variables.count(c) for c in self.coords)
                self._push(L.PW_LEAF, self._leaf(orders))
            else:  # product / chain rule: expand symbolically first (nonlinear.py:59-61)
                self._emit(sp.expand(node.doit()))
        elif not self._contains_field(node):
            if node.free_symbols:
                if not node.free_symbols.issubset(set(self.coords)):
                    raise ValueError(f"Unsupported nonlinear term with unresolved symbols: {node.free_symbols}")
                if node not in self.statics:
                    self.statics.append(node)
                self._push(L.PW_STATIC, self.statics.index(node))
            else:
                self._push(L.PW_CONST, self._const(complex(node)))
        elif isinstance(node, sp.Add):
            for i, a in enumerate(node.args):
                self._emit(a)
                if i:
                    self._push(L.PW_ADD)
        elif isinstance(node, sp.Mul):
            for i, a in enumerate(node.args):
                self._emit(a)
                if i:
                    self._push(L.PW_MUL)
        elif isinstance(node, sp.Pow):
            base, exp = node.args
            if not isinstance(exp, sp.Number):
                raise ValueError(f"Only numeric exponents are supported: {node}")
            self._emit(base)
            if exp.is_integer:
                self._push(L.PW_POWI, int(exp))
            else:
                self._push(L.PW_POWR, self._const(float(exp)))
        elif isinstance(node, sp.Abs):
            self._emit(node.args[0])
            self._push(L.PW_ABS)
        elif isinstance(node, sp.conjugate):
            self._emit(node.args[0])
            self._push(L.PW_CONJ)
        elif isinstance(node, sp.Function) and node.func.__name__ in L.FN and len(node.args) == 1:
            self._emit(node.args[0])
            self._push(L.PW_FUNC, L.FN[node.func.__name__])
        else:
            raise ValueError(f"Unsupported nonlinear term node: {node} ({type(node).__name__})")


class NonlinearTerm:
    """F(uh) = T(E(backward_primitive(uh, k_0), ...)) on the GPU (integrators/base.py:230-248)."""

    def __init__(self, space, expr, final: str = "forward", N=None, testspace=None, u=None, coords=None):
        require_device()
        self._lib = L.load()
        self.space = space
        self.testspace = space if testspace is None else testspace
        if u is None:
            u, coords = field(space)
        self.compiled = CompiledExpression(expr, u, coords)
        self._ctor = dict(expr=expr, N=N, testspace=testspace, u=u, coords=coords)
        self.final = {"forward": L.OP_FORWARD, "scalar_product": L.OP_SCALAR_PRODUCT}[final]
        self.N = N
        self._cache = {}

    def with_final(self, final: str) -> "NonlinearTerm":
        """The same term with another final transform ('forward' | 'scalar_product'), base.py:230-248."""
        c = self._ctor
        return NonlinearTerm(self.space, c["expr"], final=final, N=c["N"], testspace=c["testspace"], u=c["u"], coords=c["coords"])

    def with_resolution(self, N=None, final=None) -> "NonlinearTerm":
        """The same term evaluated at physical resolution N (None: keep this term's) and / or with another final transform —
        what `BaseIntegrator.nonlinear_rhs(uh, N)` forwards to the evaluator (base.py:230-248)."""
        c = self._ctor
        fin = final if final is not None else ("forward" if self.final == L.OP_FORWARD else "scalar_product")
        return NonlinearTerm(self.space, c["expr"], final=fin, N=(c["N"] if N is None else N), testspace=c["testspace"],
                             u=c["u"], coords=c["coords"])

    def _spaces(self, space):
        return list(space.basespaces) if hasattr(space, "basespaces") else [space]

    def _physical_shape(self, spaces):
        if self.N is None:
            return tuple(s.num_quad_points for s in spaces)
        N = self.N if isinstance(self.N, (tuple, list)) else (self.N,)
        return tuple(s.num_quad_points if n is None else int(n) for s, n in zip(spaces, N))

    def _build(self, shape, dtype):
        spaces, tspaces = self._spaces(self.space), self._spaces(self.testspace)
        d = len(spaces)
        lead = len(shape) - d
        assert lead >= 0
        phys = self._physical_shape(spaces)
        keep, leaf_descs = [], []
        for orders in self.compiled.leaves:
            specs = [None] * lead
            op = L.OP_BACKWARD_PRIMITIVE if any(orders) else L.OP_BACKWARD
            for ax, s in enumerate(spaces):
                inner = int(np.prod(shape[lead + ax + 1:], dtype=np.int64))
                specs.append(s.axis_spec(op, shape[lead + ax], dtype, phys[ax], orders[ax], inner=inner))
            desc, k = _fill_plan_desc(op, dtype, shape, specs)
            keep.append(k)
            leaf_descs.append(desc)
        pshape = tuple(shape[:lead]) + phys
        specs = [None] * lead
        for ax, s in enumerate(tspaces):
            inner = int(np.prod(pshape[lead + ax + 1:], dtype=np.int64))
            specs.append(s.axis_spec(self.final, pshape[lead + ax], dtype, inner=inner))
        fdesc, k = _fill_plan_desc(self.final, dtype, pshape, specs)
        keep.append(k)

        nd = L.NonlinearDesc()
        nd.abi_version = L.JFX_ABI_VERSION
        nd.n_leaves = len(leaf_descs)
        for i, dsc in enumerate(leaf_descs):
            nd.leaves[i] = C.pointer(dsc)
        nd.final_transform = C.pointer(fdesc)
        nd.n_program = len(self.compiled.program)
        for i, (op, arg) in enumerate(self.compiled.program):
            nd.program[i].op, nd.program[i].arg = op, arg
        nd.n_consts = len(self.compiled.consts)
        for i, cval in enumerate(self.compiled.consts):
            nd.consts[i][0], nd.consts[i][1] = cval.real, cval.imag
        # statics: sampled once on the quadrature mesh (nonlinear.py:219-242)
        static_arrays = []
        if self.compiled.statics:
            mesh = self.space.mesh(N=phys) if hasattr(self.space, "basespaces") else (self.space.mesh(N=phys[0]),)
            tdt = {L.F32: torch.float32, L.F64: torch.float64, L.C64: torch.complex64, L.C128: torch.complex128}[dtype]
            for i, sx in enumerate(self.compiled.statics):
                vals = sp.lambdify(self.compiled.coords, sx, modules="numpy")(*mesh)
                vals = np.broadcast_to(np.asarray(vals), pshape).copy()
                t = torch.from_numpy(vals).to(device="cuda", dtype=tdt).contiguous()
                static_arrays.append(t)
                nd.statics[i] = t.data_ptr()
        nd.n_statics = len(static_arrays)
        handle = C.c_void_p()
        L.check(self._lib.jfx_nonlinear_create(C.byref(nd), C.byref(handle)))
        ws = C.c_size_t()
        L.check(self._lib.jfx_nonlinear_workspace_bytes(handle, C.byref(ws)))
        so = (C.c_int64 * L.JFX_MAX_DIMS)()
        nd_out = C.c_int()
        L.check(self._lib.jfx_nonlinear_shape_out(handle, so, C.byref(nd_out)))
        entry = {
            "handle": handle, "ws_bytes": int(ws.value), "shape_out": tuple(int(so[i]) for i in range(len(shape))),
            "statics": static_arrays, "launches": int(self._lib.jfx_nonlinear_launches(handle)), "ws": None,
        }
        return entry

    def __call__(self, uh, out=None):
        if not (torch is not None and isinstance(uh, torch.Tensor) and uh.is_cuda):
            raise L.JfxError(-3, "NonlinearTerm needs a CUDA tensor; jaxfun_b200 has no CPU fallback")
        if getattr(self.space, "complex_data", False) and not uh.is_complex():
            uh = uh.to(torch.complex128 if uh.dtype == torch.float64 else torch.complex64)
        uh = uh.contiguous()
        dtype = jfx_dtype(uh.dtype)
        key = (tuple(uh.shape), dtype)
        e = self._cache.get(key)
        if e is None:
            e = self._cache[key] = self._build(tuple(uh.shape), dtype)
        if e["ws"] is None:
            e["ws"] = torch.empty(max(e["ws_bytes"], 1), dtype=torch.uint8, device=uh.device)
        if out is None:
            out = torch.empty(e["shape_out"], dtype=uh.dtype, device=uh.device)
        L.check(self._lib.jfx_nonlinear_execute(e["handle"], C.c_void_p(current_stream_ptr()), C.c_void_p(uh.data_ptr()),
                                                C.c_void_p(out.data_ptr()), C.c_void_p(e["ws"].data_ptr())))
        return out

    def launches(self, uh) -> int:
        key = (tuple(uh.shape), jfx_dtype(uh.dtype))
        return self._cache[key]["launches"] if key in self._cache else 0

    def __del__(self):
        try:
            for e in self._cache.values():
                self._lib.jfx_nonlinear_destroy(e["handle"])
            self._cache = {}
        except Exception:
            pass
