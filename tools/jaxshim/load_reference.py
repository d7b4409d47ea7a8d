"""Mount the reference's transform-path source files on the numpy stand-in for jax.

Only works where /root/reference exists (this build container).  The package __init__ files of
`jaxfun`, `jaxfun.galerkin` and `jaxfun.integrators` are NOT executed (they pull in flax / optax /
the PINN stack); the individual modules of the transform path are executed unmodified from where
they lie.  `jaxfun.la` (flax-based matrix classes, not on the transform path) is replaced by a
permissive stub.
"""
import importlib
import os
import sys
import types

REF_SRC = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))


class _Anything(types.ModuleType):
    """Module whose every attribute is a harmless placeholder class."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (), {"__init__": lambda self, *a, **k: None,
                              "__class_getitem__": classmethod(lambda c, i: c)})
        setattr(self, name, cls)
        return cls


def _namespace(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    m.__package__ = name
    sys.modules[name] = m
    return m


def mount():
    if "jaxfun.galerkin.orthogonal" in sys.modules:
        return sys.modules["jaxfun"]
    if not os.path.isdir(REF_SRC):
        raise RuntimeError("/root/reference is not available: golden vectors can only be regenerated in the build container")
    if HERE not in sys.path:
        sys.path.insert(0, HERE)  # makes `import jax` resolve to the stand-in
    import jax  # noqa: F401
    assert "jaxshim" in jax.__file__, "a real jax is importable: use it instead of the stand-in"
    root = os.path.join(REF_SRC, "jaxfun")
    jf = _namespace("jaxfun", root)
    _namespace("jaxfun.galerkin", os.path.join(root, "galerkin"))
    _namespace("jaxfun.integrators", os.path.join(root, "integrators"))
    for stub in ("jaxfun.la", "jaxfun.la.matrixprotocol", "jaxfun.la.diamatrix", "jaxfun.la.matrix",
                 "jaxfun.la.tpmatrix", "jaxfun.la.operators", "flax", "flax.nnx"):
        sys.modules[stub] = _Anything(stub)
    sys.modules["jaxfun.la"].__path__ = []
    sys.modules["flax"].__path__ = []
    sys.modules["flax"].nnx = sys.modules["flax.nnx"]
    for mod in ("jaxfun.typing", "jaxfun.utils", "jaxfun.coordinates", "jaxfun.basespace",
                "jaxfun.galerkin.orthogonal", "jaxfun.galerkin.Jacobi", "jaxfun.galerkin.Legendre",
                "jaxfun.galerkin.Chebyshev", "jaxfun.galerkin.ChebyshevU", "jaxfun.galerkin.Ultraspherical",
                "jaxfun.galerkin.Fourier"):
        importlib.import_module(mod)
    return jf


if __name__ == "__main__":
    mount()
    from jaxfun.galerkin.Chebyshev import Chebyshev
    import numpy as np
    C = Chebyshev(8)
    c = np.arange(8.0)
    print(C.backward(c))
    print(C.forward(C.backward(c)))
