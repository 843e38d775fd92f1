"""basix stand-in: only make_quadrature on the interval (Gauss-Jacobi(0,0) = Gauss-Legendre on [0,1])."""
import enum

import numpy as np


class CellType(enum.Enum):
    interval = 1
    tetrahedron = 4


class QuadratureType(enum.Enum):
    default = 0


def make_quadrature(cell, degree, rule=QuadratureType.default):
    assert cell == CellType.interval
    m = (degree + 2) // 2
    x, w = np.polynomial.legendre.leggauss(m)
    return (0.5 * (x + 1.0)).reshape(-1, 1), 0.5 * w
