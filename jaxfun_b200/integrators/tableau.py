"""Butcher tableaux of the IMEX Runge-Kutta schemes (`jaxfun.integrators.tableau`,
`src/jaxfun/integrators/tableau.py:10-127`): an explicit and a diagonally-implicit tableau with the same
abscissae.  Coefficients restated from the literature (Ascher, Ruuth, Spiteri, Appl. Numer. Math. 25
(1997): IMEX Euler (1,1,1), ARS(2,2,2), ARS(4,4,3)); the reference's larger ARK tables are constants and
can be passed in the same form."""
from __future__ import annotations

from dataclasses import dataclass

_TOL = 1e-10


@dataclass(frozen=True)
class ButcherTableau:
    A: tuple
    b: tuple
    c: tuple

    @property
    def stages(self) -> int:
        return len(self.b)


@dataclass(frozen=True)
class IMEXTableau:
    explicit: ButcherTableau
    implicit: ButcherTableau
    name: str = "imex"

    def __post_init__(self):
        s = self.explicit.stages
        assert self.implicit.stages == s, "explicit / implicit tableaux need the same number of stages"
        for i in range(s):
            for j in range(i, s):
                assert abs(self.explicit.A[i][j]) < _TOL, "explicit tableau must be strictly lower triangular"
            for j in range(i + 1, s):
                assert abs(self.implicit.A[i][j]) < _TOL, "implicit tableau must be lower triangular (DIRK)"

    @property
    def stages(self) -> int:
        return self.explicit.stages

    @property
    def explicit_is_stiffly_accurate(self) -> bool:
        e = self.explicit
        return all(abs(e.A[-1][j] - e.b[j]) < _TOL for j in range(e.stages))

    @property
    def implicit_is_stiffly_accurate(self) -> bool:
        m = self.implicit
        return all(abs(m.A[-1][j] - m.b[j]) < _TOL for j in range(m.stages))

    @property
    def is_stiffly_accurate(self) -> bool:
        return self.explicit_is_stiffly_accurate and self.implicit_is_stiffly_accurate

    @property
    def distinct_diagonal_coeffs(self) -> tuple:
        out = []
        for i in range(self.stages):
            a = self.implicit.A[i][i]
            if a != 0.0 and a not in out:
                out.append(a)
        return tuple(out)


IMEX_EULER = IMEXTableau(
    explicit=ButcherTableau(A=((0.0, 0.0), (1.0, 0.0)), b=(1.0, 0.0), c=(0.0, 1.0)),
    implicit=ButcherTableau(A=((0.0, 0.0), (0.0, 1.0)), b=(0.0, 1.0), c=(0.0, 1.0)),
    name="IMEX_EULER")


def _ars222() -> IMEXTableau:
    g = 1.0 - 2.0 ** -0.5
    d = 1.0 - 1.0 / (2.0 * g)
    return IMEXTableau(
        explicit=ButcherTableau(A=((0.0, 0.0, 0.0), (g, 0.0, 0.0), (d, 1.0 - d, 0.0)), b=(d, 1.0 - d, 0.0), c=(0.0, g, 1.0)),
        implicit=ButcherTableau(A=((0.0, 0.0, 0.0), (0.0, g, 0.0), (0.0, 1.0 - g, g)), b=(0.0, 1.0 - g, g), c=(0.0, g, 1.0)),
        name="ARS222")


ARS222 = _ars222()

ARS443 = IMEXTableau(
    explicit=ButcherTableau(
        A=((0.0, 0.0, 0.0, 0.0, 0.0),
           (1 / 2, 0.0, 0.0, 0.0, 0.0),
           (11 / 18, 1 / 18, 0.0, 0.0, 0.0),
           (5 / 6, -5 / 6, 1 / 2, 0.0, 0.0),
           (1 / 4, 7 / 4, 3 / 4, -7 / 4, 0.0)),
        b=(1 / 4, 7 / 4, 3 / 4, -7 / 4, 0.0), c=(0.0, 1 / 2, 2 / 3, 1 / 2, 1.0)),
    implicit=ButcherTableau(
        A=((0.0, 0.0, 0.0, 0.0, 0.0),
           (0.0, 1 / 2, 0.0, 0.0, 0.0),
           (0.0, 1 / 6, 1 / 2, 0.0, 0.0),
           (0.0, -1 / 2, 1 / 2, 1 / 2, 0.0),
           (0.0, 3 / 2, -3 / 2, 1 / 2, 1 / 2)),
        b=(0.0, 3 / 2, -3 / 2, 1 / 2, 1 / 2), c=(0.0, 1 / 2, 2 / 3, 1 / 2, 1.0)),
    name="ARS443")
