"""CPU: host-side mirror of the reference interface (names, validation, quadrature contract)."""

import numpy as np
import pytest

import fenicsx_ii_b200 as fx
from fenicsx_ii_b200 import synthetic as sy
from fenicsx_ii_b200.mesh import functionspace
from tests.helpers import small_cfg


def test_exports_reference_names():
    for name in ("Circle", "Disk", "PointwiseTrace", "MappedRestriction", "ReductionOperator", "Quadrature",
                 "create_interpolation_matrix"):
        assert hasattr(fx, name)


def test_operator_validation_errors_match_reference():
    """restriction_operators.py:99-106,229-233: same exceptions for the same misuse (host-side, no GPU)."""
    from fenicsx_ii_b200 import Circle, Disk

    mesh, V = small_cfg(2, 2, 2)
    gamma = sy.straight_line(3)
    with pytest.raises(ValueError):
        Circle(gamma, 0.1, degree=2)
    with pytest.raises(ValueError):
        Circle(mesh, 0.1)
    with pytest.raises(ValueError):
        Disk(mesh, 0.1)
    gamma.geometry.cmap.degree = 2
    with pytest.raises(NotImplementedError):
        Circle(gamma, 0.1)
    with pytest.raises(NotImplementedError):
        Disk(gamma, 0.1)


def test_compute_quadrature_contract_shapes():
    """ReductionOperator.compute_quadrature: points [Nr,nq,3] (trace: [Nr,3]), weights [Nr,nq], scales [Nr]
    (restriction_operators.py:17-47, quadrature.py:11-20)."""
    gamma = sy.polyline_network(10, 0.1, n_lines=2, lo=0.3, hi=0.7)
    X = np.array([[0.25], [0.75]])
    cells = np.arange(4, dtype=np.int32)
    q = fx.PointwiseTrace(gamma).compute_quadrature(cells, X)
    assert q.points.shape == (8, 3) and q.weights.shape == (8, 1) and q.scales.shape == (8,)
    c = fx.Circle(gamma, 0.05, degree=5)
    q = c.compute_quadrature(cells, X)
    assert c.num_points == 3 and q.points.shape == (8, 3, 3) and q.weights.shape == (8, 3) and q.scales.shape == (8,)
    np.testing.assert_allclose(q.weights.sum(axis=1) / q.scales, 1.0, rtol=1e-14)
    np.testing.assert_allclose(np.linalg.norm(q.points - q.points.mean(axis=1, keepdims=True) * 0 - fx.restriction_operators
                                              .get_physical_points(gamma, cells, X)[:, None, :], axis=2), 0.05, rtol=1e-12)
    d = fx.Disk(gamma, lambda x: 0.02 + 0 * x[0], degree=3)
    q = d.compute_quadrature(cells, X)
    assert d.num_points == 9 and q.points.shape == (8, 9, 3)
    np.testing.assert_allclose(q.weights.sum(axis=1) / q.scales, 1.0, rtol=1e-14)
    assert str(fx.PointwiseTrace(gamma)).startswith("PointwiseTrace(")


def test_functionspace_standins():
    mesh, V = small_cfg(3, 3, 3, degree=2)
    assert V.dofmap.list.shape == (162, 10) and V.num_dofs == 7**3
    V2 = functionspace(mesh, ("Lagrange", 2))
    assert V2.num_dofs == V.num_dofs  # generic edge numbering agrees in count with the closed form
    gamma = sy.straight_line(5)
    assert functionspace(gamma, ("Quadrature", 10)).element.interpolation_points.shape == (6, 1)
    assert functionspace(gamma, ("Lagrange", 2)).num_dofs == 5 + 4
    V3 = functionspace(mesh, ("Lagrange", 3))
    assert V3.dofmap.list.shape == (162, 20) and V3.num_dofs == 10**3  # vertices + 2 per edge + 1 per face
    with pytest.raises(NotImplementedError):
        functionspace(mesh, ("Lagrange", 4))


def test_gamma_description_skips_identity_arrays():
    """`device_rows(..., args_only=True)` + `DeviceRows._gamma_desc`: all cells of Γ in order and the cell-local
    numbering of Quadrature / DG spaces are passed as NULL (generated on the device), subsets and continuous
    spaces as arrays; the record matches fxii_rows_desc field by field (no GPU needed)."""
    from fenicsx_ii_b200 import _lib
    from fenicsx_ii_b200.interpolation import DeviceRows, device_rows

    gamma = sy.polyline_network(12, 0.1, n_lines=2, lo=0.3, hi=0.7)
    op = fx.Circle(gamma, 0.05, degree=7)
    Kq = functionspace(gamma, ("Quadrature", 4))
    args = device_rows(Kq, op, args_only=True)
    assert args[2] is None and args[4] is None  # cells_k, row_dofs: identity
    d, keep, n_point_rows, nq = DeviceRows._gamma_desc(*args)
    nip = Kq.element.interpolation_points.shape[0]
    assert (d.cells_k, d.row_dofs) == (None, None) and d.n_cells_k == 12 and d.nip == nip
    assert n_point_rows == 12 * nip and nq == op.num_points == d.nq and d.n_rows_total == Kq.num_dofs
    assert d.kind == _lib.OP_CIRCLE and d.radius_per_row == 0 and d.gx == keep[0].ctypes.data

    # the same cells given explicitly are recognised as "all cells"
    args = device_rows(Kq, op, np.arange(12, dtype=np.int32), args_only=True)
    assert args[2] is None and args[4] is None
    # a subset keeps both arrays
    sub = np.array([1, 4, 5, 9], dtype=np.int32)
    args = device_rows(Kq, op, sub, args_only=True)
    np.testing.assert_array_equal(args[2], sub)
    np.testing.assert_array_equal(args[4], Kq.dofmap.list[sub].reshape(-1))
    d, keep, n_point_rows, _ = DeviceRows._gamma_desc(*args)
    assert d.n_cells_k == 4 and n_point_rows == 4 * nip and d.cells_k == keep[2].ctypes.data
    # continuous K: all cells, but shared dofs -> explicit row numbering
    Kp = functionspace(gamma, ("Lagrange", 1))
    args = device_rows(Kp, fx.PointwiseTrace(gamma), args_only=True)
    assert args[2] is None
    np.testing.assert_array_equal(args[4], Kp.dofmap.list.reshape(-1))
    # generic operators have no device description
    assert device_rows(Kq, fx.MappedRestriction(gamma, lambda x: x), args_only=True) is None


def test_geometry_token_follows_in_place_moves():
    """The Ω handle and the Π cache are keyed on a fingerprint of the coordinates (interpolation._geometry_token)."""
    from fenicsx_ii_b200.interpolation import _cache_key, _geometry_token

    mesh, V = small_cfg(3, 3, 3)
    gamma = sy.straight_line(5)
    K = functionspace(gamma, ("DG", 0))
    t0, k0 = _geometry_token(mesh), _cache_key(V, K, None, 1e-8, None)
    assert _geometry_token(mesh) == t0 and _cache_key(V, K, None, 1e-8, None) == k0
    mesh.geometry.x[:, 1] += 0.25
    assert _geometry_token(mesh) != t0 and _cache_key(V, K, None, 1e-8, None) != k0
    mesh.geometry.x[:, 1] -= 0.25
    assert _geometry_token(mesh) == t0
    gamma.geometry.x[:, 2] *= 0.5
    assert _cache_key(V, K, None, 1e-8, None) != k0


def test_block_rows_are_refused_where_blocks_of_cells_are_not_blocks_of_rows():
    """interpolation.device_rows_block returns None — before touching the device — for a K with shared dofs, for a
    non-contiguous block of cells and for an operator that overrides the plugin contract; assemble_blocks_to_host then
    builds Π in one general call."""
    from fenicsx_ii_b200 import Circle, PointwiseTrace
    from fenicsx_ii_b200.interpolation import device_rows_block

    gamma = sy.straight_line(9)
    cells = np.arange(2, 6, dtype=np.int32)
    assert device_rows_block(functionspace(gamma, ("Lagrange", 1)), PointwiseTrace(gamma), cells) is None       # shared dofs
    assert device_rows_block(functionspace(gamma, ("DG", 1)), PointwiseTrace(gamma), np.array([1, 3, 4], dtype=np.int32)) is None

    class MyCircle(Circle):
        def compute_quadrature(self, cells, reference_points):  # a generic operator now: host points
            return super().compute_quadrature(cells, reference_points)

    assert device_rows_block(functionspace(gamma, ("Quadrature", 4)), MyCircle(gamma, 0.1, degree=5), cells) is None
    assert device_rows_block(functionspace(gamma, ("Quadrature", 4)), Circle(gamma, 0.1, degree=5), np.zeros(0, dtype=np.int32)) is None
