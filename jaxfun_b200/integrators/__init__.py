from .base import BaseIntegrator as BaseIntegrator, axpby_diag as axpby_diag  # noqa: F401
from .etdrk4 import ETDRK4 as ETDRK4, etd_coefficients as etd_coefficients  # noqa: F401
from .nonlinear import NonlinearTerm as NonlinearTerm, field as field  # noqa: F401
from .rk4 import RK4 as RK4  # noqa: F401
from .imex_rk import BackwardEuler as BackwardEuler, IMEXRungeKutta as IMEXRungeKutta  # noqa: F401
from .tableau import ARS222 as ARS222, ARS443 as ARS443, IMEX_EULER as IMEX_EULER, ButcherTableau as ButcherTableau, IMEXTableau as IMEXTableau  # noqa: F401
