"""Shared workload builders and the tie-aware parity comparator."""

from __future__ import annotations

import numpy as np

from fenicsx_ii_b200 import synthetic as sy
from fenicsx_ii_b200.mesh import functionspace
from oracle import pi_oracle as orc


def make_operator(op_name, gamma, radius=None, degree=None):
    from fenicsx_ii_b200 import Circle, Disk, PointwiseTrace

    if op_name == "trace":
        return PointwiseTrace(gamma)
    if op_name == "mapped":
        from fenicsx_ii_b200 import MappedRestriction
        from tests.golden_cases import mapped_shift

        return MappedRestriction(gamma, mapped_shift)
    if op_name == "circle":
        return Circle(gamma, radius, degree=degree)
    return Disk(gamma, radius, degree=degree)


def oracle_pi(V, K, gamma, op_name, radius=None, degree=None, cells_K=None, return_coo=False, cell_radius=None):
    """Oracle Π for the same inputs create_interpolation_matrix receives."""
    mesh = V.mesh
    deg = V.element.degree
    rad = radius
    if cell_radius is not None:
        nip = K.element.interpolation_points.shape[0]
        ck = np.arange(gamma.geometry.dofmap.shape[0]) if cells_K is None else np.asarray(cells_K)
        per_row = np.repeat(cell_radius[ck], nip)
        rad = lambda x: per_row  # noqa: E731
    return orc.assemble_pi(
        mesh.geometry.x, mesh.geometry.dofmap, V.dofmap.list, deg, gamma.geometry.x, gamma.geometry.dofmap,
        K.dofmap.list, K.element.interpolation_points, op_name, radius=rad, quad_degree=degree, cells_K=cells_K,
        n_rows=K.num_dofs, return_coo=return_coo,
    )


def assert_csr_parity(got, ref, rtol=1e-12):
    """Bit-exact indptr/indices; values within rtol * max|row| (north_star: 1e-12 relative; the
    row scale is used because relative error of a near-zero basis value is ill-posed)."""
    indptr, indices, data = got
    np.testing.assert_array_equal(indptr, ref["indptr"])
    np.testing.assert_array_equal(indices, ref["indices"])
    n = indptr.shape[0] - 1
    rows = np.repeat(np.arange(n), np.diff(indptr))
    scale = np.zeros(n)
    np.maximum.at(scale, rows, np.abs(ref["data"]))
    err = np.abs(data - ref["data"])
    bound = rtol * np.maximum(scale[rows], 1e-300)
    bad = err > bound
    assert not bad.any(), f"{bad.sum()} values differ; max err {err.max():.3e}, worst ratio {(err / bound).max():.3e}"


def small_cfg(nx=8, ny=8, nz=16, degree=1):
    mesh = sy.kuhn_box(nx, ny, nz)
    V = functionspace(mesh, ("Lagrange", 1)) if degree == 1 else sy.kuhn_p2_space(mesh)
    return mesh, V


def expand_blocks_csr(indptr, indices, data, bs: int):
    """Scalar CSR of the blocked matrix the reference fills for bs > 1 (interpolation.py:214-248 with
    utils.unroll_dofmap): entry (r, c, v) -> rows r*bs+b hold columns c*bs+bc for ALL bc, v on bc == b, 0.0 elsewhere."""
    n = indptr.shape[0] - 1
    counts = np.diff(indptr)
    new_ptr = np.zeros(n * bs + 1, dtype=np.int64)
    new_ptr[1:] = np.cumsum(np.repeat(counts * bs, bs))
    new_idx = np.empty(new_ptr[-1], dtype=np.int32)
    new_val = np.zeros(new_ptr[-1], dtype=np.float64)
    for r in range(n):
        c = indices[indptr[r]:indptr[r + 1]].astype(np.int64)
        v = data[indptr[r]:indptr[r + 1]]
        cols = (c[:, None] * bs + np.arange(bs)[None, :]).reshape(-1)
        for b in range(bs):
            lo = new_ptr[r * bs + b]
            new_idx[lo:lo + cols.shape[0]] = cols
            vals = np.zeros((c.shape[0], bs))
            vals[:, b] = v
            new_val[lo:lo + cols.shape[0]] = vals.reshape(-1)
    return new_ptr, new_idx, new_val
