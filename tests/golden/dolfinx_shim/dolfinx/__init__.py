"""dolfinx stand-in (serial, affine intervals and tetrahedra) built on oracle/pi_oracle.py."""
import numpy as np

from . import common, cpp, fem, geometry, la, mesh  # noqa: F401

default_scalar_type = np.float64
has_petsc = False
