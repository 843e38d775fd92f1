"""B200-native reduction-operator assembly path of FEniCSx_ii (Π, Πᵀ, ΠᵀMΠ, MΠ).

Exports the reference's names for the hot path (reference ``__init__.py:1-32``); everything
outside the path (UFL rewriting, form assembly, solver) stays with the reference / dolfinx.
"""

from .assembly import (
    apply_matrix_replacer,
    apply_vector_replacer,
    assign_LG_map,
    average_coefficient,
    average_coefficients,
    restrict_columns,
    restrict_rows,
    restrict_rows_and_columns,
    restrict_vector,
)
from .interpolation import clear_interpolation_cache, create_interpolation_matrix
from .quadrature import Quadrature
from .restriction_operators import Circle, Disk, MappedRestriction, PointwiseTrace, ReductionOperator
from .sparse import DeviceCSR, MatrixCSR

__all__ = [
    "Circle",
    "DeviceCSR",
    "Disk",
    "MappedRestriction",
    "MatrixCSR",
    "PointwiseTrace",
    "Quadrature",
    "ReductionOperator",
    "apply_matrix_replacer",
    "apply_vector_replacer",
    "assign_LG_map",
    "average_coefficient",
    "average_coefficients",
    "clear_interpolation_cache",
    "create_interpolation_matrix",
    "restrict_columns",
    "restrict_rows",
    "restrict_rows_and_columns",
    "restrict_vector",
]
