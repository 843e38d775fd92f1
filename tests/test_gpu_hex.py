"""GPU: hexahedral meshes (Q1 geometry and Q1 space) — the Newton pull-back of non-affine cells
(interpolation_utils.py:36-38 with a non-affine cmap; tests/test_interpolation_matrix.py:31-34,102-108 run the reference
on hexahedra).  Ownership, colliding-set size and CSR pattern bit-exact against the oracle, values 1e-12, on boxes (the
trilinear map is affine: Newton lands in one step) and on tapered cells (planar faces, genuinely non-affine map), with
points on faces / edges / vertices, through the fused and the COO path."""

import numpy as np
import pytest

from fenicsx_ii_b200 import Circle, Disk, PointwiseTrace, create_interpolation_matrix
from fenicsx_ii_b200 import synthetic as sy
from fenicsx_ii_b200._lib import FxiiError
from fenicsx_ii_b200.mesh import functionspace
from oracle import pi_oracle as orc
from tests.helpers import assert_csr_parity, oracle_pi

pytestmark = pytest.mark.gpu


def check(V, K, gamma, op, name, materialize, **kw):
    A, _, _ = create_interpolation_matrix(V, K, op, materialize_coo=materialize)
    ref = oracle_pi(V, K, gamma, name, **kw)
    cell, ncoll = A.rows.diagnostics()
    np.testing.assert_array_equal(cell, ref["cell"])
    np.testing.assert_array_equal(ncoll, np.minimum(ref["n_colliding"], 255))
    assert_csr_parity((A.indptr, A.indices, A.data), ref)
    assert A.stats["fused"] == (0 if materialize else A.stats["fused"])
    return A


@pytest.mark.parametrize("materialize", [False, True])
@pytest.mark.parametrize("taper", [0.0, 0.3, -0.25])
@pytest.mark.parametrize("kind", ["trace", "circle", "disk"])
def test_hex_parity(cuda, kind, taper, materialize):
    mesh = sy.hex_box(6, 5, 7, taper=taper)
    V = functionspace(mesh, ("Lagrange", 1))
    assert V.dofmap.list.shape[1] == 8
    gamma = sy.polyline_network(90, 1 / 6, n_lines=4, lo=0.3, hi=0.7)
    K = functionspace(gamma, ("Quadrature", 4) if kind != "trace" else ("DG", 1))
    if kind == "trace":
        A = check(V, K, gamma, PointwiseTrace(gamma), "trace", materialize)
        # an isoparametric Q1 space contains the linear functions: the trace of one is exact, also on tapered cells
        x = mesh.geometry.x
        u = 0.3 * x[:, 0] - 1.7 * x[:, 1] + 0.5 * x[:, 2] + 2.0
        x0 = orc.physical_points(gamma.geometry.x, gamma.geometry.dofmap, np.arange(gamma.geometry.dofmap.shape[0]),
                                 K.element.interpolation_points)
        np.testing.assert_allclose(A.mult(u), 0.3 * x0[:, 0] - 1.7 * x0[:, 1] + 0.5 * x0[:, 2] + 2.0, rtol=0, atol=1e-13)
    elif kind == "circle":
        A = check(V, K, gamma, Circle(gamma, 0.07, degree=9), "circle", materialize, radius=0.07, degree=9)
    else:
        A = check(V, K, gamma, Disk(gamma, 0.06, degree=3), "disk", materialize, radius=0.06, degree=3)
    np.testing.assert_allclose(A.mult(np.ones(V.num_dofs)), 1.0, rtol=0, atol=1e-13)


def test_hex_points_on_faces_edges_vertices(cuda):
    """The demo's line x = y = 0.5 in a 4x4x8 box mesh lies on cell EDGES (4 colliding hexahedra, 8 at the nodes that
    coincide with mesh vertices); lowest cell index wins; trilinear functions are reproduced exactly on boxes."""
    mesh = sy.hex_box(4, 4, 8)
    V = functionspace(mesh, ("Lagrange", 1))
    gamma = sy.straight_line(9)
    for K in (functionspace(gamma, ("Quadrature", 10)), functionspace(gamma, ("Lagrange", 1))):
        for materialize in (False, True):
            A = check(V, K, gamma, PointwiseTrace(gamma), "trace", materialize)
            _, ncoll = A.rows.diagnostics()
            assert ncoll.min() >= 4 and (ncoll.max() == 8) == (K.element.family == "Lagrange")
            x = mesh.geometry.x
            q = (1.0 + x[:, 0]) * (2.0 - x[:, 1]) * (0.5 + x[:, 2])  # trilinear: in Q1 on boxes
            x0 = orc.physical_points(gamma.geometry.x, gamma.geometry.dofmap, np.arange(8), K.element.interpolation_points)
            ref = ((1.0 + x0[:, 0]) * (2.0 - x0[:, 1]) * (0.5 + x0[:, 2])).reshape(8, -1)
            got = A.mult(q)[K.dofmap.list]
            np.testing.assert_allclose(got, ref, rtol=0, atol=1e-13)
    check(V, functionspace(gamma, ("Quadrature", 10)), gamma, Circle(gamma, 0.25, degree=5), "circle", False, radius=0.25, degree=5)


def test_hex_points_outside_are_reported(cuda):
    mesh = sy.hex_box(3, 3, 3, taper=0.2)
    V = functionspace(mesh, ("Lagrange", 1))
    gamma = sy.line_mesh(np.array([[0.5, 0.5, 0.5], [0.5, 0.5, 1.4]]))  # leaves the mesh at z = 1
    K = functionspace(gamma, ("DG", 1))
    from fenicsx_ii_b200._lib import PointsNotFound

    with pytest.raises(PointsNotFound):
        create_interpolation_matrix(V, K, PointwiseTrace(gamma))


def test_unsupported_spaces_are_rejected(cuda):
    """P2-sized dofmaps on hexahedra, or Q1-sized ones on tetrahedra, are refused by the C ABI (FXII_E_UNSUPPORTED)."""
    from fenicsx_ii_b200.interpolation import DeviceMesh, DeviceSpace

    hexm = sy.hex_box(2, 2, 2)
    dm = DeviceMesh(hexm.geometry.x, hexm.geometry.dofmap, 1e-8)
    with pytest.raises(FxiiError):
        DeviceSpace(dm, np.zeros((8, 10), dtype=np.int32), 2, 100)
    tet = sy.kuhn_box(2, 2, 2)
    dt = DeviceMesh(tet.geometry.x, tet.geometry.dofmap, 1e-8)
    with pytest.raises(FxiiError):
        DeviceSpace(dt, np.zeros((48, 8), dtype=np.int32), 1, 27)
