"""Time-integrator base for diagonal-operator spectral problems — the stage-arithmetic part of
`jaxfun.integrators.base.BaseIntegrator` (`src/jaxfun/integrators/base.py:230-344`).

Scope: `nonlinear_rhs`, `nonlinear_rhs_scalar_product`, `linear_rhs`,
`total_rhs`, `solve`.  The reference derives mass / linear operators from a SymPy weak form
(`inner`, out of scope); here the semi-discrete system is given directly in coefficient space

        d(uh)/dt = Ldiag * uh  +  N(uh)             (mass matrix already divided out)

with `Ldiag` a diagonal operator (an array shaped like uh — Fourier symbols such as (i k)^3, or any
diagonalised operator) and N a `NonlinearTerm`.  All stage arithmetic runs on the device through
`jfx_axpby_diag`; every nonlinear evaluation is one `jfx_nonlinear_execute`.
"""
from __future__ import annotations

import ctypes as C

from .. import _lib as L
from ..engine import current_stream_ptr, jfx_dtype

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


def axpby_diag(terms, out=None):
    """out = sum_t alpha_t * coeff_t (.) x_t   with terms = [(alpha, coeff_or_None, x), ...] on the device."""
    lib = L.load()
    n_terms = len(terms)
    x0 = terms[0][2]
    if out is None:
        out = torch.empty_like(x0)
    coeff_arr = (C.c_void_p * n_terms)()
    x_arr = (C.c_void_p * n_terms)()
    alpha = (C.c_double * n_terms)()
    ccomplex = any(c is not None and c.is_complex() for _, c, _ in terms)
    if ccomplex and not x0.is_complex():
        raise TypeError("complex diagonal coefficients need a complex state array")
    # the kernel reads every coefficient array in ONE element type: the state's complex type when any coefficient is complex,
    # else the real type of the state's precision (ETDRK4 / IMEX coefficients arrive as float64 / complex128 from numpy: a
    # complex64 / float32 state must not have them read as if they were single precision)
    real_dt = {torch.float32: torch.float32, torch.float64: torch.float64,
               torch.complex64: torch.float32, torch.complex128: torch.float64}[x0.dtype]
    want = x0.dtype if ccomplex else real_dt
    keep = []
    for i, (a, c, x) in enumerate(terms):
        assert x.shape == x0.shape and x.dtype == x0.dtype and x.is_contiguous()
        if c is not None:
            if c.is_complex() and not ccomplex:
                raise TypeError("mixed real / complex coefficient handling failed")   # unreachable: ccomplex covers it
            if c.dtype != want:
                c = c.to(want)
            c = c.contiguous()
            assert c.shape == x0.shape
            keep.append(c)
        coeff_arr[i] = c.data_ptr() if c is not None else None
        x_arr[i] = x.data_ptr()
        alpha[i] = float(a)
    L.check(lib.jfx_axpby_diag(C.c_void_p(current_stream_ptr()), n_terms, coeff_arr, alpha, x_arr,
                               C.c_void_p(out.data_ptr()), x0.numel(), jfx_dtype(x0.dtype), int(ccomplex)))
    return out


class BaseIntegrator:
    def __init__(self, space, linear_diag=None, nonlinear=None, forcing=None):
        self.trialspace = self.testspace = space
        self.Ldiag = linear_diag
        self.nonlinear = nonlinear
        self.forcing = forcing
        self.has_nonlinear = nonlinear is not None

    # base.py:230-236: N (padded physical resolution) is forwarded to the evaluator; None = the test space's own shape
    def nonlinear_rhs(self, uh, N=None):
        if not self.has_nonlinear:
            return torch.zeros_like(uh)
        return self._nonlinear_for(N)(uh)

    # base.py:238-248
    def nonlinear_rhs_scalar_product(self, uh, N=None):
        if not self.has_nonlinear:
            return torch.zeros_like(uh)
        return self._nonlinear_for(N, final="scalar_product")(uh)

    def _nonlinear_for(self, N=None, final=None):
        """The NonlinearTerm evaluated at physical resolution N (None: the term's own) with the requested final transform."""
        term = self.nonlinear
        if N is None and final is None:
            return term
        if not hasattr(term, "with_resolution"):
            return term
        key = (None if N is None else (tuple(N) if isinstance(N, (tuple, list)) else int(N)), final)
        cache = self.__dict__.setdefault("_nl_variants", {})
        t = cache.get(key)
        if t is None:
            t = cache[key] = term.with_resolution(N, final)
        return t

    # base.py:250-255
    def linear_rhs(self, uh):
        terms = [(1.0, self.Ldiag, uh)]
        if self.forcing is not None:
            terms.append((1.0, None, self.forcing))
        return axpby_diag(terms)

    # base.py:257-260
    def total_rhs(self, uh, N=None):
        if not self.has_nonlinear:
            return self.linear_rhs(uh)
        terms = [(1.0, self.Ldiag, uh), (1.0, None, self.nonlinear_rhs(uh, N))]
        if self.forcing is not None:
            terms.append((1.0, None, self.forcing))
        return axpby_diag(terms)

    def setup(self, dt: float) -> None:
        ...

    def step(self, u_hat, dt: float, N=None):
        raise NotImplementedError

    # base.py:269-344 (no snapshots/progress bar: the loop itself)
    def solve(self, u0, dt: float, steps: int, N=None, graph: bool = False):
        """`graph=True` captures ONE step (its nonlinear evaluations and fused diagonal combinations: only kernel launches,
        no allocation outside the capture pool, no host synchronisation) into a CUDA graph after two eager warm-up steps and
        replays it — what the reference gets from `lax.fori_loop` inside `jit` (base.py:329): one launch per step instead of
        ~25 for ETDRK4."""
        self.setup(dt)
        u = u0
        done = 0
        if graph and steps > 3 and torch is not None and getattr(u0, "is_cuda", False):
            side = torch.cuda.Stream(device=u0.device)
            side.wait_stream(torch.cuda.current_stream(u0.device))
            with torch.cuda.stream(side):                      # warm-up on the capture stream: plans, tables, workspaces
                buf = u0
                for _ in range(2):
                    buf = self.step(buf, dt, N)
                    done += 1
                static_in = buf.clone()
            torch.cuda.current_stream(u0.device).wait_stream(side)
            torch.cuda.synchronize(u0.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                static_out = self.step(static_in, dt, N)
            for _ in range(steps - done):
                g.replay()
                static_in.copy_(static_out)
            u = static_in.clone()
        else:
            for _ in range(steps):
                u = self.step(u, dt, N)
        if not bool(torch.isfinite(torch.view_as_real(u) if u.is_complex() else u).all()):
            raise FloatingPointError("time integration diverged")  # base.py:332-334
        return u
