from .base import BaseIntegrator as BaseIntegrator, axpby_diag as axpby_diag  # noqa: F401
from .etdrk4 import ETDRK4 as ETDRK4, etd_coefficients as etd_coefficients  # noqa: F401
from .nonlinear import NonlinearTerm as NonlinearTerm, field as field  # noqa: F401
from .rk4 import RK4 as RK4  # noqa: F401
