"""GPU parity of Π against the CPU oracle, through the public API / C ABI.

Bar (north_star): cell ownership, sparsity pattern and CSR indices bit-exact; values within
1e-12 of the row scale."""

import numpy as np
import pytest

from fenicsx_ii_b200 import create_interpolation_matrix
from fenicsx_ii_b200 import synthetic as sy
from fenicsx_ii_b200.mesh import functionspace
from oracle import pi_oracle as orc
from tests.helpers import assert_csr_parity, make_operator, oracle_pi, small_cfg

pytestmark = pytest.mark.gpu


def run_case(V, K, gamma, op_name, radius=None, degree=None, cells_K=None, cell_radius=None, expect_fused=None):
    """Both device paths against the oracle: the default one (rows reduced inside the locate kernel
    when K is cell-local) and the forced COO path (whose COO is compared entry by entry)."""
    rad = cell_radius if cell_radius is not None else radius
    op = make_operator(op_name, gamma, rad, degree)
    ref = oracle_pi(V, K, gamma, op_name, radius, degree, cells_K=cells_K, return_coo=True, cell_radius=cell_radius)
    A = None
    for materialize in (True, False):
        A, imap_K, imap_V = create_interpolation_matrix(V, K, op, cells_K=cells_K, materialize_coo=materialize)
        cell, ncoll = A.rows.diagnostics()
        np.testing.assert_array_equal(cell, ref["cell"])
        np.testing.assert_array_equal(ncoll.astype(np.int64), np.minimum(ref["n_colliding"], 255))
        if materialize:
            assert A.stats["fused"] == 0
            cols, vals = A.rows.coo()
            np.testing.assert_array_equal(cols, ref["cols"])
        elif expect_fused is not None:
            assert bool(A.stats["fused"]) == expect_fused
        assert_csr_parity((A.indptr, A.indices, A.data), ref)
        assert A.shape == (K.num_dofs, V.num_dofs)
        assert imap_K is K.dofmap.index_map and imap_V is V.dofmap.index_map
    return A, ref


@pytest.mark.parametrize("op_name,radius,degree", [("trace", None, None), ("circle", 0.2, 5), ("circle", 0.05, 5),
                                                   ("disk", 0.1, 3), ("disk", 0.07, 7)])
def test_config1_straight_line_ties(cuda, op_name, radius, degree):
    """Config 1 geometry (scaled down): the line x=y=0.5 lies on mesh edges, so every centre collides
    with several tetrahedra; the lowest cell index must win on both sides."""
    mesh, V = small_cfg(8, 8, 16)
    gamma = sy.straight_line(17)
    K = functionspace(gamma, ("Quadrature", 10))
    A, ref = run_case(V, K, gamma, op_name, radius, degree, expect_fused=True)
    assert (ref["n_colliding"] > 1).any()
    np.testing.assert_allclose(np.asarray(A.to_scipy().sum(axis=1)).ravel(), 1.0, atol=1e-13)


@pytest.mark.parametrize("kspace", [("Quadrature", 10), ("DG", 1), ("DG", 2), ("Lagrange", 1), ("Lagrange", 2)])
@pytest.mark.parametrize("op_name,radius,degree", [("trace", None, None), ("circle", 0.03, 7)])
def test_network_generic_position(cuda, kspace, op_name, radius, degree):
    """Config 2 shape: polyline network in generic position; continuous K exercises first-visit-wins."""
    mesh, V = small_cfg(12, 12, 12)
    gamma = sy.polyline_network(400, 1 / 12, n_lines=5)
    K = functionspace(gamma, kspace)
    # continuous K shares dofs between Γ cells: the fused path must detect it and fall back
    A, ref = run_case(V, K, gamma, op_name, radius, degree, expect_fused=kspace[0] != "Lagrange")


@pytest.mark.parametrize("op_name,degree", [("circle", 15), ("circle", 126), ("disk", 7)])
def test_p2_vessel_tree_per_segment_radius(cuda, op_name, degree):
    """Config 3/4 shape: P2, vessel tree with per-segment radii, 64-point circle / 49-point disk."""
    mesh, V = small_cfg(10, 10, 10, degree=2)
    h = 1 / 10
    gamma, seg_r = sy.vessel_tree(300, h / 2, r_min=0.2 * h, r_max=1.5 * h, branch_len=40)
    K = functionspace(gamma, ("Quadrature", 10))
    run_case(V, K, gamma, op_name, None, degree, cell_radius=seg_r, expect_fused=True)


def test_cells_subset_and_callable_radius(cuda):
    mesh, V = small_cfg(8, 8, 8)
    gamma = sy.polyline_network(120, 1 / 8, n_lines=3, lo=0.25, hi=0.75)
    K = functionspace(gamma, ("DG", 1))
    cells_K = np.arange(5, 97, 3, dtype=np.int32)
    rad = lambda x: 0.02 + 0.1 * x[2] ** 2  # noqa: E731
    A, ref = run_case(V, K, gamma, "circle", rad, 9, cells_K=cells_K)
    # rows of cells outside cells_K are empty
    touched = np.zeros(K.num_dofs, dtype=bool)
    touched[K.dofmap.list[cells_K].ravel()] = True
    assert (np.diff(A.indptr)[~touched] == 0).all()


def test_empty_cells(cuda):
    mesh, V = small_cfg(4, 4, 4)
    gamma = sy.straight_line(5)
    K = functionspace(gamma, ("Quadrature", 4))
    from fenicsx_ii_b200 import PointwiseTrace

    A, _, _ = create_interpolation_matrix(V, K, PointwiseTrace(gamma), cells_K=np.zeros(0, dtype=np.int32))
    assert A.device.nnz == 0 and A.shape == (K.num_dofs, V.num_dofs)


def test_point_outside_raises(cuda):
    from fenicsx_ii_b200 import PointwiseTrace
    from fenicsx_ii_b200._lib import PointsNotFound

    mesh, V = small_cfg(4, 4, 4)
    nodes = np.array([[0.5, 0.5, 0.5], [0.5, 0.5, 1.5]])
    gamma = sy.line_mesh(nodes)
    K = functionspace(gamma, ("DG", 1))
    with pytest.raises(PointsNotFound):
        create_interpolation_matrix(V, K, PointwiseTrace(gamma))


def test_mapped_restriction_host_points(cuda):
    """Generic operators go through compute_quadrature on the host (FXII_OP_POINTS path)."""
    from fenicsx_ii_b200 import MappedRestriction

    mesh, V = small_cfg(6, 6, 6)
    gamma = sy.polyline_network(60, 1 / 6, n_lines=2, lo=0.3, hi=0.6)
    K = functionspace(gamma, ("DG", 1))
    shift = lambda x: x + np.array([[0.05], [0.02], [-0.03]])  # noqa: E731
    op = MappedRestriction(gamma, shift)
    A, _, _ = create_interpolation_matrix(V, K, op)
    q = op.compute_quadrature(np.arange(60, dtype=np.int32), K.element.interpolation_points)
    from oracle import pi_oracle as orc

    ref = orc.assemble_pi(mesh.geometry.x, mesh.geometry.dofmap, V.dofmap.list, 1, gamma.geometry.x, gamma.geometry.dofmap,
                          K.dofmap.list, K.element.interpolation_points, "points", n_rows=K.num_dofs,
                          points=q.points, weights=q.weights, scales=q.scales)
    assert_csr_parity((A.indptr, A.indices, A.data), ref)
    u = 2.0 * mesh.geometry.x[:, 0] - mesh.geometry.x[:, 2] + 0.3
    np.testing.assert_allclose(A.mult(u), 2.0 * q.points[:, 0] - q.points[:, 2] + 0.3, rtol=1e-12, atol=1e-13)


def test_run_to_run_bitwise_reproducible(cuda):
    mesh, V = small_cfg(8, 8, 8, degree=2)
    gamma, seg_r = sy.vessel_tree(200, 1 / 16, branch_len=50)
    K = functionspace(gamma, ("Lagrange", 1))
    from fenicsx_ii_b200 import Disk

    op = Disk(gamma, seg_r, degree=5)
    A1, _, _ = create_interpolation_matrix(V, K, op)
    A2, _, _ = create_interpolation_matrix(V, K, op)
    np.testing.assert_array_equal(A1.indices, A2.indices)
    assert A1.data.tobytes() == A2.data.tobytes()


@pytest.mark.parametrize("nq_degree", [3, 4, 11, 17, 31, 33, 47, 65, 95, 129])
def test_fused_team_sizes(cuda, nq_degree):
    """Circle rules with 2..65 points exercise every team shape of the fused kernel (sub-warp teams,
    one warp, several warps, block sizes 128/192/256)."""
    mesh, V = small_cfg(6, 6, 6)
    gamma = sy.polyline_network(40, 1 / 6, n_lines=2, lo=0.3, hi=0.7)
    K = functionspace(gamma, ("DG", 1))
    run_case(V, K, gamma, "circle", 0.08, nq_degree, expect_fused=True)


def test_fused_p2_disk_many_points(cuda):
    mesh, V = small_cfg(6, 6, 6, degree=2)
    gamma = sy.polyline_network(30, 1 / 6, n_lines=2, lo=0.35, hi=0.65)
    K = functionspace(gamma, ("Quadrature", 2))
    run_case(V, K, gamma, "disk", 0.12, 11, expect_fused=True)  # 121 points per row
    run_case(V, K, gamma, "disk", 0.12, 17, expect_fused=False)  # 289 points per row: COO path


@pytest.mark.parametrize("mode", [0, 2])
@pytest.mark.parametrize("op_name,radius,degree,p", [("trace", None, None, 1), ("circle", 0.05, 15, 1), ("disk", 0.06, 5, 2)])
def test_repeated_builds_on_one_handle(cuda, op_name, radius, degree, p, mode):
    """Build 1 reads the nnz back before allocating; build 2+ know it (one host sync).
    mode 0: every build is ONE cooperative kernel (stats.fused == 3).  mode 2: separate kernels,
    builds 2+ replayed as one CUDA graph launch.  All builds must be bitwise identical."""
    from fenicsx_ii_b200.interpolation import device_rows, device_space

    mesh, V = small_cfg(8, 8, 8, degree=p)
    gamma = sy.polyline_network(200, 1 / 8, n_lines=4, lo=0.25, hi=0.75)
    K = functionspace(gamma, ("Quadrature", 6))
    op = make_operator(op_name, gamma, radius, degree)
    ds = device_space(V)
    rows = device_rows(K, op)
    rows.set_mode(mode)
    results, modes = [], []
    for _ in range(4):
        csr, st = rows.assemble(ds)
        results.append(csr.to_host())
        modes.append(st["fused"])
    assert modes == ([3, 3, 3, 3] if mode == 0 else [1, 2, 2, 2])
    for r in results[1:]:
        np.testing.assert_array_equal(r[0], results[0][0])
        np.testing.assert_array_equal(r[1], results[0][1])
        assert r[2].tobytes() == results[0][2].tobytes()
    ref = oracle_pi(V, K, gamma, op_name, radius, degree)
    assert_csr_parity(results[-1], ref)


def test_cooperative_and_multi_kernel_paths_agree(cuda):
    """The single cooperative launch and the multi-kernel fused path must give the same bytes
    (P2 space, subset of Γ cells so that matrix rows have gaps, per-segment radius)."""
    from fenicsx_ii_b200.interpolation import device_rows, device_space

    mesh, V = small_cfg(6, 6, 6, degree=2)
    gamma = sy.polyline_network(150, 1 / 6, n_lines=3, lo=0.3, hi=0.7)
    K = functionspace(gamma, ("DG", 1))
    op = make_operator("circle", gamma, 0.07, 9)
    cells_K = np.arange(5, 140, 3, dtype=np.int32)
    ds = device_space(V)
    out = {}
    for mode in (0, 2):
        rows = device_rows(K, op, cells_K)
        rows.set_mode(mode)
        rows.assemble(ds)
        csr, st = rows.assemble(ds)  # second build: nnz hint known
        assert st["fused"] == (3 if mode == 0 else 2)
        out[mode] = csr.to_host()
    np.testing.assert_array_equal(out[0][0], out[2][0])
    np.testing.assert_array_equal(out[0][1], out[2][1])
    assert out[0][2].tobytes() == out[2][2].tobytes()


def test_single_tetrahedron_mesh(cuda):
    """Smallest possible Ω: the LBVH is one leaf."""
    from fenicsx_ii_b200.mesh import Mesh

    x = np.array([[0.0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]])
    mesh = Mesh(x, np.array([[0, 1, 2, 3]]), tdim=3)
    V = functionspace(mesh, ("Lagrange", 1))
    gamma = sy.line_mesh(np.array([[0.1, 0.1, 0.1], [0.2, 0.25, 0.3], [0.3, 0.2, 0.35]]))
    K = functionspace(gamma, ("DG", 2))
    A, ref = run_case(V, K, gamma, "trace", expect_fused=True)
    assert (ref["cell"] == 0).all()
    run_case(V, K, gamma, "circle", 0.03, 9, expect_fused=True)


@pytest.mark.parametrize("direction", [(0, 0, 1), (0, 0, -1), (1, 0, 0), (1e-9, 0, 1), (0.3, -0.5, 0.8)])
def test_tangent_directions_including_ez(cuda, direction):
    """rotation_matrix special case: the axis falls back to ez when ez x n vanishes (quadrature.py:46-51)."""
    mesh, V = small_cfg(6, 6, 6)
    d = np.array(direction, dtype=float)
    d /= np.linalg.norm(d)
    start = np.array([0.41, 0.47, 0.39])
    nodes = start[None, :] + np.linspace(0, 0.2, 5)[:, None] * d[None, :] * np.array([1.0, 1.0, 1.0])
    gamma = sy.line_mesh(nodes)
    K = functionspace(gamma, ("Quadrature", 4))
    run_case(V, K, gamma, "circle", 0.11, 11, expect_fused=True)
    run_case(V, K, gamma, "disk", 0.11, 4, expect_fused=True)


def test_points_on_domain_boundary_and_vertices(cuda):
    """Γ running along a boundary edge of Ω and through mesh vertices: maximal tie sets."""
    mesh, V = small_cfg(4, 4, 4)
    nodes = np.array([[0.0, 0.0, 0.0], [0.0, 0.0, 0.25], [0.0, 0.0, 0.5], [0.25, 0.25, 0.75], [0.5, 0.5, 1.0]])
    gamma = sy.line_mesh(nodes)
    K = functionspace(gamma, ("Lagrange", 1))  # dofs at mesh vertices, shared between segments
    A, ref = run_case(V, K, gamma, "trace", expect_fused=False)
    assert ref["n_colliding"].max() >= 6


def test_circle_leaving_the_domain_raises(cuda):
    from fenicsx_ii_b200 import Circle
    from fenicsx_ii_b200._lib import PointsNotFound

    mesh, V = small_cfg(4, 4, 4)
    gamma = sy.line_mesh(np.array([[0.5, 0.05, 0.2], [0.5, 0.05, 0.8]]))
    K = functionspace(gamma, ("DG", 1))
    with pytest.raises(PointsNotFound):
        create_interpolation_matrix(V, K, Circle(gamma, 0.2, degree=9))
    try:
        create_interpolation_matrix(V, K, Circle(gamma, 0.2, degree=9))
    except AssertionError as e:  # PointsNotFound is an AssertionError, like the reference's assert
        assert "points lie in no cell" in str(e)


def test_blocked_spaces_and_cache(cuda):
    """tests/test_interpolation_matrix.py:254-301 (vector spaces, bs=3): Π_blocked = Π ⊗ I_3 with the
    reference's full 3x3 blocks (explicit zeros off the diagonal); plus the Π cache."""
    import scipy.sparse as sp

    from fenicsx_ii_b200 import Circle, clear_interpolation_cache

    mesh, _ = small_cfg(6, 6, 6)
    gamma = sy.polyline_network(50, 1 / 6, n_lines=2, lo=0.3, hi=0.7)
    V = functionspace(mesh, ("Lagrange", 1))
    K = functionspace(gamma, ("DG", 1))
    V3 = functionspace(mesh, ("Lagrange", 1, (3,)))
    K3 = functionspace(gamma, ("DG", 1, (3,)))
    op = Circle(gamma, 0.05, degree=7)
    A, _, _ = create_interpolation_matrix(V, K, op)
    A3, _, _ = create_interpolation_matrix(V3, K3, op)
    assert A3.shape == (3 * K.num_dofs, 3 * V.num_dofs) == (K3.num_dofs, V3.num_dofs)
    S, S3 = A.to_scipy(), A3.to_scipy()
    np.testing.assert_array_equal(S3.toarray(), sp.kron(S, sp.identity(3)).toarray())
    assert S3.nnz == 9 * S.nnz  # full blocks, zeros kept
    # Π u for a vector field interpolates component-wise
    u = np.stack([mesh.geometry.x[:, 2], 2 * mesh.geometry.x[:, 2] + 1, np.ones(V.num_dofs)], axis=1).reshape(-1)
    r = A3.mult(u).reshape(-1, 3)
    np.testing.assert_allclose(r[:, 1], 2 * r[:, 0] + 1, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(r[:, 2], 1.0, rtol=1e-12)
    # cache: same objects -> same matrix object; different operator -> rebuilt
    clear_interpolation_cache()
    B1, _, _ = create_interpolation_matrix(V, K, op, cache=True)
    B2, _, _ = create_interpolation_matrix(V, K, op, cache=True)
    B3, _, _ = create_interpolation_matrix(V, K, Circle(gamma, 0.05, degree=7), cache=True)
    assert B1 is B2 and B3 is not B1
    np.testing.assert_array_equal(B3.indices, B1.indices)
    clear_interpolation_cache()


@pytest.mark.parametrize("degree", [1, 2])
def test_unstructured_delaunay_mesh(cuda, degree):
    """Irregular Ω (Delaunay tetrahedra of random points, incl. badly shaped cells): the LBVH, the
    quantised boxes and the collision tests must not depend on mesh structure."""
    from scipy.spatial import Delaunay

    from fenicsx_ii_b200.mesh import Mesh

    rng = np.random.default_rng(7)
    pts = rng.uniform(0, 1, size=(1500, 3))
    corners = np.array([[i, j, k] for i in (0, 1) for j in (0, 1) for k in (0, 1)], dtype=float)
    pts = np.vstack([corners, pts])
    tets = Delaunay(pts).simplices.astype(np.int32)
    vol = np.abs(np.linalg.det(pts[tets[:, 1:]] - pts[tets[:, :1]])) / 6
    tets = tets[vol > 1e-9]  # drop exactly degenerate slivers (a FEM mesh has none)
    mesh = Mesh(pts, tets, tdim=3)
    V = functionspace(mesh, ("Lagrange", degree))
    gamma = sy.polyline_network(120, 0.04, n_lines=4, lo=0.2, hi=0.8)
    K = functionspace(gamma, ("Quadrature", 6))
    A, ref = run_case(V, K, gamma, "trace", expect_fused=True)
    np.testing.assert_allclose(np.asarray(A.to_scipy().sum(axis=1)).ravel(), 1.0, atol=1e-9)
    run_case(V, K, gamma, "circle", 0.07, 15, expect_fused=True)
    run_case(V, K, gamma, "disk", 0.05, 4, expect_fused=True)


def test_build_delivered_to_host_arrays(cuda):
    """create_interpolation_matrix(..., host_out=...) — fxii_pi_assemble_host — gives the same bytes as the
    device-resident build + download, on the first call (no nnz hint) and on later ones (speculated download)."""
    import torch

    from fenicsx_ii_b200 import create_interpolation_matrix

    mesh, V = small_cfg(8, 8, 8, degree=1)
    gamma = sy.polyline_network(300, 1 / 8, n_lines=4, lo=0.25, hi=0.75)
    K = functionspace(gamma, ("Quadrature", 4))
    for op_name, radius, degree in (("trace", None, None), ("circle", 0.06, 9)):
        op = make_operator(op_name, gamma, radius, degree)
        A0, _, _ = create_interpolation_matrix(V, K, op)
        ref = (A0.indptr.copy(), A0.indices.copy(), A0.data.copy())
        cap = ref[1].shape[0] + 17
        ip = torch.empty(K.num_dofs + 1, dtype=torch.int64).pin_memory().numpy()
        ix = torch.empty(cap, dtype=torch.int32).pin_memory().numpy()
        va = torch.empty(cap, dtype=torch.float64).pin_memory().numpy()
        op2 = make_operator(op_name, gamma, radius, degree)  # fresh operator: no nnz hint on the first call
        for _ in range(3):
            ix[:] = -7
            A, _, _ = create_interpolation_matrix(V, K, op2, host_out=(ip, ix, va))
            assert A.shape == A0.shape
            np.testing.assert_array_equal(A.indptr, ref[0])
            np.testing.assert_array_equal(A.indices, ref[1])
            assert A.data.tobytes() == ref[2].tobytes()
        # the device matrix is rebuilt from the host arrays on demand
        y = A.mult(np.ones(V.num_dofs))
        np.testing.assert_allclose(y, A0.mult(np.ones(V.num_dofs)), rtol=0, atol=0)
        small = (ip, ix[:5], va[:5])
        with pytest.raises(ValueError):
            create_interpolation_matrix(V, K, make_operator(op_name, gamma, radius, degree), host_out=small)


def test_host_delivery_of_a_generic_operator(cuda):
    """Operators without a device descriptor (uploaded points) take fxii_pi_assemble_host instead of the one-shot call."""
    import torch

    from fenicsx_ii_b200 import MappedRestriction, create_interpolation_matrix

    mesh, V = small_cfg(6, 6, 6, degree=1)
    gamma = sy.polyline_network(60, 1 / 6, n_lines=2, lo=0.3, hi=0.7)
    K = functionspace(gamma, ("DG", 1))
    op = MappedRestriction(gamma, lambda x: x + 0.01)
    A0, _, _ = create_interpolation_matrix(V, K, op)
    ip = torch.empty(K.num_dofs + 1, dtype=torch.int64).pin_memory().numpy()
    ix = torch.empty(A0.indices.shape[0], dtype=torch.int32).pin_memory().numpy()
    va = torch.empty(A0.indices.shape[0], dtype=torch.float64).pin_memory().numpy()
    for _ in range(2):
        A, _, _ = create_interpolation_matrix(V, K, op, host_out=(ip, ix, va))
        np.testing.assert_array_equal(A.indptr, A0.indptr)
        np.testing.assert_array_equal(A.indices, A0.indices)
        assert A.data.tobytes() == A0.data.tobytes()


def test_continuous_k_remembers_the_general_path(cuda):
    """Continuous K (dofs shared between point rows): the first build tries the fused kernel, finds the conflict and
    reruns the COO path; later builds on the handle — and later calls with the same K — go there directly."""
    from fenicsx_ii_b200 import create_interpolation_matrix
    from fenicsx_ii_b200.interpolation import device_rows, device_space

    mesh, V = small_cfg(6, 6, 6, degree=1)
    gamma = sy.polyline_network(80, 1 / 6, n_lines=2, lo=0.3, hi=0.7)
    K = functionspace(gamma, ("Lagrange", 1))
    op = make_operator("circle", gamma, 0.05, 7)
    ds = device_space(V)
    rows = device_rows(K, op)
    out = []
    for _ in range(3):
        csr, st = rows.assemble(ds)
        assert st["fused"] == 0
        out.append(csr.to_host())
    for r in out[1:]:
        np.testing.assert_array_equal(r[0], out[0][0])
        np.testing.assert_array_equal(r[1], out[0][1])
        assert r[2].tobytes() == out[0][2].tobytes()
    A1, _, _ = create_interpolation_matrix(V, K, op)
    assert K.__dict__.get("_fxii_general_path") is True
    A2, _, _ = create_interpolation_matrix(V, K, op)
    np.testing.assert_array_equal(A1.indices, A2.indices)
    assert A1.data.tobytes() == A2.data.tobytes() == out[0][2].tobytes()
    ref = oracle_pi(V, K, gamma, "circle", 0.05, 7)
    assert_csr_parity((A2.indptr, A2.indices, A2.data), ref)


# ---- P3 (tests/test_interpolation_matrix.py:174-247 of the reference: V = ("Lagrange", 3)) -------------------------------
def _p3_dof_coordinates(V):
    mesh = V.mesh
    P = np.einsum("nv,cvd->cnd", orc.p3_nodes(), mesh.geometry.x[mesh.geometry.dofmap])
    coords = np.zeros((V.num_dofs, 3))
    coords[V.dofmap.list.reshape(-1)] = P.reshape(-1, 3)
    return coords


@pytest.mark.parametrize("op_name", ["trace", "circle", "disk"])
def test_p3_parity_with_oracle(cuda, op_name):
    mesh = sy.kuhn_box(4, 3, 5)
    V = functionspace(mesh, ("Lagrange", 3))
    gamma = sy.polyline_network(40, 0.2, n_lines=3, lo=0.3, hi=0.7)
    K = functionspace(gamma, ("Quadrature", 4))
    radius, degree = (None, None) if op_name == "trace" else ((0.06, 7) if op_name == "circle" else (0.05, 3))
    op = make_operator(op_name, gamma, radius, degree)
    A, _, _ = create_interpolation_matrix(V, K, op)
    ref = oracle_pi(V, K, gamma, op_name, radius, degree)
    cell, ncoll = A.rows.diagnostics()
    np.testing.assert_array_equal(cell, ref["cell"])
    assert_csr_parity((A.indptr, A.indices, A.data), ref)


@pytest.mark.parametrize("case", [1, 2, 3])
@pytest.mark.parametrize("family,degree", [("Lagrange", 2), ("DG", 2), ("Quadrature", 3)])
def test_p3_circle_trace_of_polynomials(cuda, family, degree, case):
    """The reference's own `test_circle_trace` (tests/test_interpolation_matrix.py:174-247), cases with a polynomial f:
    V = P3 represents f exactly, so Π applied to the nodal interpolant of f equals the interpolant of the circle
    average of f in K — x2 -> x2, x0²+x1² -> R², x2(x0²+x1²) -> x2 R² for a line along x2 through the origin."""
    from fenicsx_ii_b200 import Circle
    from fenicsx_ii_b200.mesh import Mesh

    mesh = sy.kuhn_box(4, 4, 4)
    mesh = Mesh(mesh.geometry.x * 2.0 - np.array([1.0, 1.0, 0.0]), mesh.geometry.dofmap, tdim=3)  # [-1,1]^2 x [0,2]
    V = functionspace(mesh, ("Lagrange", 3))
    nodes = np.zeros((9, 3))
    nodes[:, 2] = np.linspace(0.1, 1.9, 9)
    gamma = sy.line_mesh(nodes)
    K = functionspace(gamma, (family, degree))
    R = 0.53
    f = {1: lambda x: x[:, 2], 2: lambda x: x[:, 0] ** 2 + x[:, 1] ** 2, 3: lambda x: x[:, 2] * (x[:, 0] ** 2 + x[:, 1] ** 2)}[case]
    pif = {1: lambda z: z, 2: lambda z: np.full_like(z, R * R), 3: lambda z: z * R * R}[case]
    A, _, _ = create_interpolation_matrix(V, K, Circle(gamma, R, degree=6))
    u = f(_p3_dof_coordinates(V))
    b = A.to_scipy() @ u
    # K's interpolation points in physical space (cell-local dofs for DG / Quadrature; shared for Lagrange)
    xr = np.asarray(K.element.interpolation_points).reshape(-1)
    z = nodes[gamma.geometry.dofmap[:, 0], 2][:, None] * (1 - xr)[None, :] + nodes[gamma.geometry.dofmap[:, 1], 2][:, None] * xr[None, :]
    expect = np.zeros(K.num_dofs)
    expect[K.dofmap.list.reshape(-1)] = pif(z.reshape(-1))
    np.testing.assert_allclose(b, expect, rtol=1e-12, atol=1e-12)


def test_p3_parity_on_an_unstructured_mesh(cuda):
    """P3 on a Delaunay mesh: the edges of a cell run with and against their reference direction, so the two dofs of an
    edge appear in both orders in the dofmap — ownership, pattern and values must still match the oracle."""
    from scipy.spatial import Delaunay

    from fenicsx_ii_b200.mesh import Mesh

    rng = np.random.default_rng(5)
    g = np.linspace(0.0, 1.0, 5)
    pts = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3)
    inner = (pts > 0).all(axis=1) & (pts < 1).all(axis=1)
    pts[inner] += rng.uniform(-0.06, 0.06, size=(int(inner.sum()), 3))
    tets = Delaunay(pts).simplices.astype(np.int32)
    vol = np.einsum("ij,ij->i", np.cross(pts[tets[:, 1]] - pts[tets[:, 0]], pts[tets[:, 2]] - pts[tets[:, 0]]), pts[tets[:, 3]] - pts[tets[:, 0]])
    tets = np.ascontiguousarray(tets[np.abs(vol) > 1e-10])
    mesh = Mesh(pts, tets, tdim=3)
    V = functionspace(mesh, ("Lagrange", 3))
    flipped = (tets[:, 0] > tets[:, 1]).mean()
    assert 0.05 < flipped < 0.95  # both edge orientations occur
    gamma = sy.polyline_network(30, 0.2, n_lines=3, lo=0.25, hi=0.75)
    K = functionspace(gamma, ("DG", 1))
    op = make_operator("circle", gamma, 0.05, 7)
    A, _, _ = create_interpolation_matrix(V, K, op)
    ref = oracle_pi(V, K, gamma, "circle", 0.05, 7)
    cell, _ = A.rows.diagnostics()
    np.testing.assert_array_equal(cell, ref["cell"])
    assert_csr_parity((A.indptr, A.indices, A.data), ref)
