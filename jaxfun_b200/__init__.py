"""jaxfun_b200 — B200-native tensor-product spectral transforms (drop-in for the transform hot
path of spectralDNS/jaxfun).  Everything numerical runs in libjfx.so (CUDA, sm_100a) through the
C ABI in include/jfx.h; importing the package loads that library and fails if it is missing."""
from . import _lib

_lib.load()  # fail loudly when the CUDA extension has not been built

from . import galerkin  # noqa: E402
from .engine import PinnedArray, Plan, device_count, require_device  # noqa: E402,F401
from .galerkin import TensorProduct, TensorProductSpace  # noqa: E402,F401
from .galerkin.composite import Composite, DirectSum, FunctionSpace  # noqa: E402,F401
from .galerkin.Chebyshev import Chebyshev  # noqa: E402,F401
from .galerkin.ChebyshevU import ChebyshevU  # noqa: E402,F401
from .galerkin.Fourier import Fourier  # noqa: E402,F401
from .galerkin.Jacobi import Jacobi  # noqa: E402,F401
from .galerkin.Legendre import Legendre  # noqa: E402,F401
from .galerkin.Ultraspherical import Ultraspherical  # noqa: E402,F401

__all__ = ["galerkin", "Plan", "PinnedArray", "TensorProduct", "TensorProductSpace", "Chebyshev",
           "ChebyshevU", "Composite", "DirectSum", "FunctionSpace", "Fourier", "Jacobi", "Legendre", "Ultraspherical", "device_count",
           "require_device"]
