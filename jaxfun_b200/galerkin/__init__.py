from . import (  # noqa: F401
    Chebyshev as Chebyshev,
    ChebyshevU as ChebyshevU,
    Fourier as Fourier,
    Jacobi as Jacobi,
    Legendre as Legendre,
    Ultraspherical as Ultraspherical,
    orthogonal as orthogonal,
)
from .tensorproductspace import TensorProduct as TensorProduct, TensorProductSpace as TensorProductSpace  # noqa: F401
from .composite import Composite as Composite, FunctionSpace as FunctionSpace  # noqa: F401,E402
