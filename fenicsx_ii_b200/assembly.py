"""Block products and coefficient averaging on the device.

Reference entry points (file:line under /root/reference/src/fenicsx_ii/):
  apply_matrix_replacer cases [0], [1], [0,1]   matrix_assembler.py:175-258
  assign_LG_map                                  assembly.py:19-44
  average_coefficients (Π build + Π·u)           assembly.py:47-68
The UFL rewriting that decides *which* case applies (ufl_operations.py) and the dolfinx form
assembly that produces A are out of scope and stay with stock dolfinx; these functions take the
assembled A (scipy CSR, DeviceCSR, or a PETSc Mat) and the spaces / operators.
"""

from __future__ import annotations

import numpy as np

from .interpolation import create_interpolation_matrix
from .sparse import DeviceCSR, MatrixCSR

__all__ = ["assign_LG_map", "average_coefficient", "restrict_columns", "restrict_rows", "restrict_rows_and_columns"]


def _as_device(A) -> DeviceCSR:
    if isinstance(A, DeviceCSR):
        return A
    if isinstance(A, MatrixCSR):
        return A.device
    if hasattr(A, "getValuesCSR"):  # petsc4py Mat
        indptr, indices, data = A.getValuesCSR()
        return DeviceCSR.from_host(indptr, indices, data, A.getSize())
    return DeviceCSR.from_scipy(A)


def _pi(space, target, red_op, tol):
    A, _, _ = create_interpolation_matrix(space, target, red_op, tol=tol, use_petsc=False, cache=True)
    return A.device


def restrict_rows(A, test_parent_space, test_space, test_red_op, tol: float = 1e-8, pi: DeviceCSR | None = None) -> DeviceCSR:
    """Case ``[0]`` (matrix_assembler.py:175-199): ``D = Πᵀ A`` with Π = Π(test)."""
    K = pi if pi is not None else _pi(test_parent_space, test_space, test_red_op, tol)
    return K.transpose().matmult(_as_device(A))


def restrict_columns(A, trial_parent_space, trial_space, trial_red_op, tol: float = 1e-8, pi: DeviceCSR | None = None) -> DeviceCSR:
    """Case ``[1]`` (matrix_assembler.py:201-223): ``D = A Π`` with Π = Π(trial)."""
    K = pi if pi is not None else _pi(trial_parent_space, trial_space, trial_red_op, tol)
    return _as_device(A).matmult(K)


def restrict_rows_and_columns(A, test_parent_space, test_space, test_red_op, trial_parent_space, trial_space,
                              trial_red_op, tol: float = 1e-8, pi_test: DeviceCSR | None = None,
                              pi_trial: DeviceCSR | None = None, association: str = "reference") -> DeviceCSR:
    """Case ``[0,1]`` (matrix_assembler.py:225-258): ``D = Pᵀ A K``.

    ``association="reference"`` evaluates ``(Pᵀ A) K`` left to right as the reference does;
    ``"north_star"`` evaluates ``Pᵀ (A K)`` (the row-sharded form of SURVEY §8e).  Same matrix and
    pattern; floating-point rounding order differs."""
    P = pi_test if pi_test is not None else _pi(test_parent_space, test_space, test_red_op, tol)
    Kt = pi_trial if pi_trial is not None else _pi(trial_parent_space, trial_space, trial_red_op, tol)
    Ad = _as_device(A)
    Pt = P.transpose()
    if association == "reference":
        return Pt.matmult(Ad).matmult(Kt)
    return Pt.matmult(Ad.matmult(Kt))


def average_coefficient(u_values: np.ndarray, V, K, red_op, tol: float = 1e-8) -> np.ndarray:
    """``average_coefficients`` for one coefficient (assembly.py:57-67): build Π(V -> K, red_op) and
    return ``Π u``."""
    A, _, _ = create_interpolation_matrix(V, K, red_op, tol=tol, use_petsc=False)
    return A.device.mult(np.asarray(u_values, dtype=np.float64))


def assign_LG_map(C, row_map, col_map, row_bs: int, col_bs: int):
    """assembly.py:19-44 for PETSc matrices (requires petsc4py)."""
    from petsc4py import PETSc

    grow = row_map.local_to_global(np.arange(row_map.size_local + row_map.num_ghosts, dtype=np.int32)).astype(PETSc.IntType)
    gcol = col_map.local_to_global(np.arange(col_map.size_local + col_map.num_ghosts, dtype=np.int32)).astype(PETSc.IntType)
    C.setLGMap(PETSc.LGMap().create(grow, bsize=row_bs, comm=row_map.comm),
               PETSc.LGMap().create(gcol, bsize=col_bs, comm=col_map.comm))
