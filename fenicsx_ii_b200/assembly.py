"""Block products and coefficient averaging on the device, behind the reference's entry points.

Reference entry points (file:line under /root/reference/src/fenicsx_ii/):
  average_coefficients(form)                     assembly.py:47-68
  assign_LG_map(C, row_map, col_map, bs, bs)     assembly.py:19-44
  apply_matrix_replacer(a, get_matrix, post, M)  matrix_assembler.py:130-264  (cases [], [0], [1], [0,1])
  apply_vector_replacer(form, get_vector, ...)   vector_assembler.py:120-185  (cases [], [0]: Πᵀ b)
The UFL rewriting that splits a form and decides which case applies (``ufl_operations.apply_replacer``,
``get_replaced_argument_indices``) and the dolfinx assembly that produces ``A`` / ``b`` are out of scope and stay with the
reference / stock dolfinx: they are passed in (or imported from ``fenicsx_ii`` when that package is installed).  What
runs here instead of PETSc is Π itself and the products ``Πᵀ A``, ``A Π``, ``(Πᵀ A) Π``, ``Π u`` and ``Πᵀ b``.

Objects are duck-typed on the attributes the reference reads: averaged coefficients have ``parent_coefficient``,
``restriction_operator``, ``function_space`` and ``x`` (``x.array``, optionally ``x.scatter_forward()``); averaged
arguments have ``parent_space``, ``restriction_operator``, ``ufl_function_space()`` and ``number()``.
Matrices may be petsc4py ``Mat`` (anything with ``getValuesCSR``), scipy sparse matrices, ``DeviceCSR`` or ``MatrixCSR``;
PETSc in gives PETSc AIJ out (one ``createAIJ(csr=...)`` per product, LG maps assigned as the reference does), anything
else gives a device-resident ``MatrixCSR``.
"""

from __future__ import annotations

import typing

import numpy as np

from .interpolation import create_interpolation_matrix
from .sparse import DeviceCSR, MatrixCSR

__all__ = ["apply_matrix_replacer", "apply_vector_replacer", "assign_LG_map", "average_coefficient", "average_coefficients",
           "restrict_columns", "restrict_rows", "restrict_rows_and_columns", "restrict_vector"]


def _is_petsc_mat(A) -> bool:
    return hasattr(A, "getValuesCSR") and hasattr(A, "getSize")


def _as_device(A) -> DeviceCSR:
    if isinstance(A, DeviceCSR):
        return A
    if isinstance(A, MatrixCSR):
        return A.device
    if _is_petsc_mat(A):  # petsc4py Mat: assembled AIJ, sorted unique columns
        indptr, indices, data = A.getValuesCSR()
        return DeviceCSR.from_host(indptr, indices, data, A.getSize())
    return DeviceCSR.from_scipy(A)


def _pi(space, target, red_op, tol, cache: bool = True) -> DeviceCSR:
    """Π(space -> target, red_op) on the device.  The reference rebuilds it for every integral of every block
    (matrix_assembler.py:177,203,229,239); here it is memoised per (space, target, operator, tol) object identity —
    see ``clear_interpolation_cache`` for when to drop it (moved meshes, changed radius functions).  Products never
    modify it."""
    A, _, _ = create_interpolation_matrix(space, target, red_op, tol=tol, use_petsc=False, cache=cache)
    return A.device


# ---- low-level products (explicit Π or spaces) ---------------------------------------------------------
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


def restrict_vector(b, test_parent_space, test_space, test_red_op, tol: float = 1e-8, pi: DeviceCSR | None = None) -> np.ndarray:
    """``z = Πᵀ b`` (vector_assembler.py:161-181), b and z as numpy arrays."""
    K = pi if pi is not None else _pi(test_parent_space, test_space, test_red_op, tol)
    return K.mult_transpose(np.asarray(b, dtype=np.float64))


def average_coefficient(u_values: np.ndarray, V, K, red_op, tol: float = 1e-8) -> np.ndarray:
    """One coefficient of ``average_coefficients`` (assembly.py:57-67) on plain arrays: build Π(V -> K, red_op) and
    return ``Π u``."""
    A, _, _ = create_interpolation_matrix(V, K, red_op, tol=tol, use_petsc=False)
    return A.device.mult(np.asarray(u_values, dtype=np.float64))


# ---- the reference's entry points -------------------------------------------------------------------------
def _is_averaged_coefficient(c) -> bool:
    return hasattr(c, "parent_coefficient") and (hasattr(c, "restriction_operator") or hasattr(c, "_restriction_operator"))


def _restriction_operator(obj):
    op = getattr(obj, "restriction_operator", None)
    return op if op is not None else getattr(obj, "_restriction_operator")


def _vector_array(x) -> np.ndarray:
    """The numpy view of a dolfinx ``la.Vector`` (``x.array``) or of a PETSc ``Vec`` (``getArray`` / ``array``)."""
    if hasattr(x, "array"):
        return np.asarray(x.array)
    return np.asarray(x.getArray())


def average_coefficients(form, tol: float = 1e-8):
    """Perform the average operation on all coefficients of a given form (assembly.py:47-68).

    For every ``AveragedCoefficient`` ``c`` of ``form.coefficients()`` the interpolation matrix from the parent
    coefficient's space to ``c.function_space`` is built on the device and ``c.x`` is overwritten with ``Π u``
    (the reference's ``K.mult(u.x.petsc_vec, coeff.x.petsc_vec)``); ``c.x.scatter_forward()`` is called on every
    coefficient, as the reference does.

    Note:
        This updates the underlying vector of each coefficient.
    """
    for coeff in form.coefficients():
        if _is_averaged_coefficient(coeff):
            u = coeff.parent_coefficient
            K, _, _ = create_interpolation_matrix(u.function_space, coeff.function_space, _restriction_operator(coeff), tol=tol,
                                                  use_petsc=False)
            src = _vector_array(u.x)
            dst = _vector_array(coeff.x)
            n = K.shape[0]
            dst[:n] = K.device.mult(np.ascontiguousarray(src[: K.shape[1]], dtype=np.float64))
        x = getattr(coeff, "x", None)
        if x is not None and hasattr(x, "scatter_forward"):
            x.scatter_forward()


def assign_LG_map(C, row_map, col_map, row_bs: int, col_bs: int):
    """Assign a Local to Global map (LG-map) to a PETSc matrix (assembly.py:19-44; requires petsc4py)."""
    from petsc4py import PETSc

    grow = row_map.local_to_global(np.arange(row_map.size_local + row_map.num_ghosts, dtype=np.int32)).astype(PETSc.IntType)
    gcol = col_map.local_to_global(np.arange(col_map.size_local + col_map.num_ghosts, dtype=np.int32)).astype(PETSc.IntType)
    C.setLGMap(PETSc.LGMap().create(grow, bsize=row_bs, comm=row_map.comm),
               PETSc.LGMap().create(gcol, bsize=col_bs, comm=col_map.comm))


def _default_apply_replacer(a):
    try:  # the reference's own UFL rewriting, when that package (and UFL) is installed
        from fenicsx_ii.ufl_operations import apply_replacer  # type: ignore

        return apply_replacer(a)
    except ImportError:
        return list(a) if isinstance(a, (list, tuple)) else [a]


def _default_replaced_indices(form) -> list[int]:
    """ufl_operations.py:419-425 on duck-typed arguments: the numbers of the arguments that carry a restriction."""
    return [arg.number() for arg in form.arguments() if hasattr(arg, "parent_space") and
            (hasattr(arg, "restriction_operator") or hasattr(arg, "_restriction_operator"))]


def _wrap_like(A_in, D: DeviceCSR, row_space, col_space, col_map=None):
    """PETSc in -> PETSc AIJ out with the LG maps of assembly.py:19-44; otherwise a device-resident MatrixCSR."""
    M = MatrixCSR(D)
    if not _is_petsc_mat(A_in):
        return M
    P = M.to_petsc()
    assign_LG_map(P, row_space.dofmap.index_map, col_map if col_map is not None else col_space.dofmap.index_map,
                  row_space.dofmap.index_map_bs, col_space.dofmap.index_map_bs)
    return P


def _accumulate(matrix, D):
    if matrix is None:
        return D
    matrix.axpy(1.0, D)  # PETSc Mat.axpy, or MatrixCSR.axpy (device union of the patterns)
    return matrix


def apply_matrix_replacer(a, get_matrix: typing.Callable, post_operation: typing.Callable = lambda *args, **kwargs: None,
                          matrix=None, *, apply_replacer: typing.Callable | None = None,
                          get_replaced_argument_indices: typing.Callable | None = None, tol: float = 1e-8):
    """matrix_assembler.py:130-264.  Given a bi-linear form, split it (``apply_replacer``), obtain the matrix of every
    part from ``get_matrix``, apply ``post_operation`` to it, replace the rows and/or columns of averaged arguments by
    multiplying with the interpolation matrices — on the device — and sum the parts into one matrix, which is returned.

    Args:
        a: The bi-linear form to process (any object with ``arguments()``; a list of already split forms is accepted
            when the reference's UFL machinery is not installed).
        get_matrix: A function that takes a form and returns its assembled matrix (PETSc ``Mat``, scipy, ``DeviceCSR``
            or ``MatrixCSR``).
        post_operation: A function applied to each assembled matrix (e.g. ``A.assemble()``).
        matrix: An optional matrix to which the result is added.  If None, a new matrix is created.
        apply_replacer, get_replaced_argument_indices: (extension) the UFL-side helpers of
            ``fenicsx_ii.ufl_operations``; default: that module when importable, else duck-typed equivalents.
    """
    if hasattr(a, "arguments") and len(a.arguments()) != 2:
        raise ValueError("The form has more than two arguments, cannot assemble a matrix.")
    split = apply_replacer or _default_apply_replacer
    indices_of = get_replaced_argument_indices or _default_replaced_indices
    for avg_form in split(a):
        A = get_matrix(avg_form)
        post_operation(A)
        test_arg, trial_arg = avg_form.arguments()
        replacement_indices = list(indices_of(avg_form))
        if replacement_indices == []:
            D = A if _is_petsc_mat(A) else (A if isinstance(A, MatrixCSR) else MatrixCSR(_as_device(A)))
        elif replacement_indices == [0]:
            # replace rows: interpolation matrix (transposed) on the left
            test_space, trial_space = test_arg.parent_space, trial_arg.ufl_function_space()
            K = _pi(test_space, test_arg.ufl_function_space(), _restriction_operator(test_arg), tol)
            D = _wrap_like(A, K.transpose().matmult(_as_device(A)), test_space, trial_space)
        elif replacement_indices == [1]:
            # replace columns: interpolation matrix on the right
            trial_space, test_space = trial_arg.parent_space, test_arg.ufl_function_space()
            Pi, _, imap_trial = create_interpolation_matrix(trial_space, trial_arg.ufl_function_space(),
                                                            _restriction_operator(trial_arg), tol=tol, use_petsc=False, cache=True)
            D = _wrap_like(A, _as_device(A).matmult(Pi.device), test_space, trial_space, col_map=imap_trial)
        elif replacement_indices == [0, 1]:
            # rows first (Pᵀ A), then columns ((Pᵀ A) K): the reference's association
            test_space, trial_space = test_arg.parent_space, trial_arg.parent_space
            P = _pi(test_space, test_arg.ufl_function_space(), _restriction_operator(test_arg), tol)
            Z = P.transpose().matmult(_as_device(A))
            K = _pi(trial_space, trial_arg.ufl_function_space(), _restriction_operator(trial_arg), tol)
            D = _wrap_like(A, Z.matmult(K), test_space, trial_space)
        else:
            raise ValueError(f"Unexpected replacement indices {replacement_indices}")
        matrix = _accumulate(matrix, D)
    assert matrix is not None
    return matrix


def apply_vector_replacer(form, get_vector: typing.Callable, post_assembly: typing.Callable = lambda *args, **kwargs: None,
                          vec=None, *, apply_replacer: typing.Callable | None = None,
                          get_replaced_argument_indices: typing.Callable | None = None, tol: float = 1e-8):
    """vector_assembler.py:120-185.  Linear forms: parts whose test function is averaged are assembled on the
    restricted space and mapped back with ``z = Πᵀ b`` (device SpMV with the transposed interpolation matrix).
    ``get_vector`` may return a numpy array (numpy out) or a PETSc ``Vec`` (PETSc ``Vec`` out)."""
    if hasattr(form, "arguments") and len(form.arguments()) != 1:
        raise ValueError("The form must have exactly one argument.")
    split = apply_replacer or _default_apply_replacer
    indices_of = get_replaced_argument_indices or _default_replaced_indices
    for avg_form in split(form):
        b = get_vector(avg_form)
        post_assembly(b)
        replacement_indices = list(indices_of(avg_form))
        test_arg = avg_form.arguments()[0]
        if replacement_indices == []:
            z = b
        elif replacement_indices == [0]:
            K = _pi(test_arg.parent_space, test_arg.ufl_function_space(), _restriction_operator(test_arg), tol)
            is_numpy = isinstance(b, np.ndarray)
            z_arr = K.mult_transpose(np.ascontiguousarray(b if is_numpy else _vector_array(b), dtype=np.float64)[: K.shape[0]])
            if is_numpy:
                z = z_arr
            else:
                from petsc4py import PETSc

                z = PETSc.Vec().createWithArray(z_arr)
        else:
            raise ValueError(f"Unexpected replacement indices {replacement_indices}")
        if vec is None:
            vec = z
        elif isinstance(vec, np.ndarray):
            vec = vec + np.asarray(z)
        else:
            vec.axpy(1.0, z)
    assert vec is not None
    return vec
