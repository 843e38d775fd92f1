"""CPU: pins the oracle (and the product's host-side quadrature) to the golden vectors produced by
the reference's own code (tests/golden/make_golden.py) and to the reference's test tables."""

import os

import numpy as np
import pytest

from fenicsx_ii_b200 import quadrature as pq
from fenicsx_ii_b200 import restriction_operators as pro
from oracle import pi_oracle as orc
from tests.golden_cases import CASES, GOLDEN_DIR, load_case

Q = np.load(os.path.join(GOLDEN_DIR, "quadrature_golden.npz"))


@pytest.mark.parametrize("n", range(1, 11))
def test_disk_rule_matches_reference_output(n):
    for impl in (orc.compute_disk_quadrature, pq.compute_disk_quadrature):
        p, w = impl(n)
        np.testing.assert_allclose(p, Q[f"disk_points_{n}"], rtol=0, atol=2e-16)
        np.testing.assert_allclose(w, Q[f"disk_weights_{n}"], rtol=0, atol=2e-16)


@pytest.mark.parametrize("n", range(1, 11))
def test_disk_rule_matches_bojanov_petrova_tables(n):
    """Port of the reference's tests/test_circle_quadrature.py:231-250 on its own tables (same
    tolerances: assert_allclose default rtol=1e-7, atol=1e-16)."""
    points, weights = orc.compute_disk_quadrature(n)
    _, qw = orc.make_quadrature_interval(2 * n - 1)
    m = len(qw)
    np.testing.assert_allclose(points[:, 0], np.repeat(Q[f"table_x_{n}"], m), atol=1e-16)
    coeff = Q[f"table_coeff_{n}"]
    for i in range(n):
        scale = 2 * np.sqrt(1 - np.cos((i + 1) * np.pi / (n + 1)) ** 2)
        np.testing.assert_allclose(weights[i * m : (i + 1) * m], scale * qw * coeff[i], atol=1e-16)


@pytest.mark.parametrize("n", range(1, 8))
def test_disk_rule_polynomial_exactness(n):
    """Port of tests/test_circle_quadrature.py:253-269 (sympy integrals, rtol=atol=1e-12)."""
    import sympy as sp

    points, weights = orc.compute_disk_quadrature(n)
    x, y = sp.symbols("x y")
    for m in range(n):
        f = lambda X: 2 + (X[0] ** (n - m) - 0.1) * (X[1] ** m - 0.2)  # noqa: E731
        exact = sp.integrate(f([x, y]), (y, -sp.sqrt(1 - x**2), sp.sqrt(1 - x**2)), (x, -1, 1))
        np.testing.assert_allclose(np.sum(weights * f(points.T)), float(exact), rtol=1e-12, atol=1e-12)


def test_rotation_matrix_matches_reference():
    for impl in (orc.rotation_matrix, pq.rotation_matrix):
        R = impl(Q["rot_normals"].copy())
        np.testing.assert_allclose(R, Q["rot_matrices"], rtol=0, atol=5e-16)
    # special cases quoted in SURVEY §8a row a6
    np.testing.assert_allclose(orc.rotation_matrix(np.array([[0.0, 0, 1]]))[0], np.eye(3), atol=1e-16)
    np.testing.assert_allclose(orc.rotation_matrix(np.array([[0.0, 0, -1]]))[0], np.diag([-1.0, -1, 1]), atol=1e-16)


@pytest.mark.parametrize("tag,kind,deg,radius", [("circle5", "circle", 5, 0.37), ("circle15", "circle", 15, 0.37),
                                                 ("circle126", "circle", 126, 0.37), ("circle9c", "circle", 9, "callable"),
                                                 ("disk3", "disk", 3, 0.21), ("disk7", "disk", 7, 0.21)])
def test_circle_disk_quadrature_matches_reference(tag, kind, deg, radius):
    x0, nn = Q["avg_x0"], Q["avg_normals"]
    R = 0.1 + x0[:, 2] ** 2 if radius == "callable" else np.full(x0.shape[0], radius)
    table, w = orc.circle_table(deg) if kind == "circle" else orc.disk_table(deg)
    p, wt, s = orc.averaging_points(x0, nn, R, table, w, kind)
    np.testing.assert_allclose(p, Q[f"{tag}_points"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(wt, Q[f"{tag}_weights"], rtol=1e-15, atol=0)
    np.testing.assert_allclose(s, Q[f"{tag}_scales"], rtol=1e-15, atol=0)
    # the product's host-side operator classes follow the same contract
    from fenicsx_ii_b200 import synthetic as sy

    line = sy.line_mesh(np.array([[0.0, 0, 0], [0, 0, 1.0]]))
    rad = (lambda x: 0.1 + x[2] ** 2) if radius == "callable" else radius
    op = pro.Circle(line, rad, degree=deg) if kind == "circle" else pro.Disk(line, rad, degree=deg)
    p2, w2, s2 = op.quadrature(x0.copy(), nn.copy())
    np.testing.assert_allclose(p2, Q[f"{tag}_points"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(w2, Q[f"{tag}_weights"], rtol=1e-15, atol=0)
    np.testing.assert_allclose(s2, Q[f"{tag}_scales"], rtol=1e-15, atol=0)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_pi_matches_reference_orchestration(name):
    """Oracle Π == Π built by the reference's unmodified interpolation.py on the dolfinx shim:
    bit-exact pattern and ownership, values to 1e-14."""
    mesh, V, gamma, K, op_name, radius, degree, cells_K, fx = load_case(name)
    bs = int(V.dofmap.bs)
    extra = {}
    if op_name == "mapped":  # generic operator: the oracle takes the mapped points (weights 1, scales 1)
        from tests.golden_cases import mapped_shift

        ck = np.arange(gamma.geometry.dofmap.shape[0]) if cells_K is None else cells_K
        phys = orc.physical_points(gamma.geometry.x, gamma.geometry.dofmap, ck, K.element.interpolation_points)
        pts = mapped_shift(phys.T).T
        extra = dict(points=pts, weights=np.ones((pts.shape[0], 1)), scales=np.ones(pts.shape[0]))
        op_name = "points"
    r = orc.assemble_pi(mesh.geometry.x, mesh.geometry.dofmap, V.dofmap.list, V.element.degree, gamma.geometry.x,
                        gamma.geometry.dofmap, K.dofmap.list, K.element.interpolation_points, op_name, radius=radius,
                        quad_degree=degree, cells_K=cells_K, n_rows=K.dofmap.index_map.size_local, **extra)
    if bs > 1:  # blocked spaces: the reference's matrix is the scalar one with every entry blown up to a bs x bs block
        from tests.helpers import expand_blocks_csr

        r["indptr"], r["indices"], r["data"] = expand_blocks_csr(r["indptr"], r["indices"], r["data"], bs)
    np.testing.assert_array_equal(r["indptr"], fx["indptr"])
    np.testing.assert_array_equal(r["indices"], fx["indices"])
    np.testing.assert_array_equal(r["cell"], fx["cell"])
    np.testing.assert_array_equal(r["n_colliding"], fx["n_colliding"])
    np.testing.assert_allclose(r["data"], fx["data"], rtol=0, atol=1e-14)


def test_first_visit_wins_leaves_explicit_zeros():
    """Continuous K: rows shared by two Γ cells get the union pattern, values from the first visit only
    (interpolation.py:208-217,242) — visible in the golden fixture as explicit zeros."""
    # (for the trace the second visit hits the same tetrahedron, so its zeros coincide with entries of the
    # first visit; the circle's points differ between the two segments sharing a vertex dof)
    _, _, _, _, _, _, _, _, fx = load_case("circle_network_p2k_firstvisit")
    assert (fx["data"] == 0.0).sum() > 0
    row_sums = np.add.reduceat(fx["data"], fx["indptr"][:-1][np.diff(fx["indptr"]) > 0])
    np.testing.assert_allclose(row_sums, 1.0, atol=1e-14)
