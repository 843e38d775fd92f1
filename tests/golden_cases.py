"""Rebuilds the inputs of the golden Π cases (same specs as tests/golden/make_golden.py, which
cannot be imported on the GPU box because it needs /root/reference)."""

from __future__ import annotations

import os

import numpy as np

from fenicsx_ii_b200 import synthetic as sy
from fenicsx_ii_b200.mesh import functionspace

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CASES = {
    "trace_line_on_edges": dict(n=(4, 4, 8), p=1, gamma=("line", 9), K=("Quadrature", 10), op=("trace",)),
    "circle_line_R02": dict(n=(4, 4, 8), p=1, gamma=("line", 9), K=("Quadrature", 10), op=("circle", 0.2, 5)),
    "disk_network_dg1": dict(n=(6, 6, 6), p=1, gamma=("network", 60, 1 / 6, 3), K=("DG", 1), op=("disk", 0.06, 3)),
    "trace_network_p1_firstvisit": dict(n=(6, 6, 6), p=1, gamma=("network", 60, 1 / 6, 3), K=("Lagrange", 1), op=("trace",)),
    "circle_network_p2k_firstvisit": dict(n=(5, 5, 5), p=1, gamma=("network", 40, 1 / 5, 2), K=("Lagrange", 2),
                                          op=("circle", 0.05, 7)),
    "circle_p2v_callable_radius": dict(n=(4, 4, 4), p=2, gamma=("network", 30, 1 / 4, 2), K=("Quadrature", 4),
                                       op=("circle", "callable", 9)),
    "disk_p2v_subset": dict(n=(4, 4, 4), p=2, gamma=("network", 30, 1 / 4, 2), K=("DG", 2), op=("disk", 0.05, 3)),
    # generic operator: points computed on the host by the operator and uploaded (restriction_operators.py:318-362)
    "mapped_network_dg1": dict(n=(5, 5, 5), p=1, gamma=("network", 40, 1 / 5, 2), K=("DG", 1), op=("mapped",)),
    # blocked (vector) spaces: the fixture is the scalar CSR of the bs x bs blocked matrix the reference fills
    "trace_blocked3_dg2": dict(n=(4, 4, 4), p=1, bs=3, gamma=("network", 24, 1 / 4, 2), K=("DG", 2), op=("trace",)),
    "circle_blocked2_p1k_firstvisit": dict(n=(4, 4, 4), p=1, bs=2, gamma=("network", 24, 1 / 4, 2), K=("Lagrange", 1),
                                           op=("circle", 0.05, 5)),
    # hexahedra (tests/test_interpolation_matrix.py:31-34): Q1 on boxes with the line on mesh edges, and on tapered
    # (non-affine, planar-faced) cells where the pull-back is a genuine Newton iteration
    "trace_hex_line_on_edges": dict(n=(4, 4, 8), p=1, hex=0.0, gamma=("line", 9), K=("Quadrature", 10), op=("trace",)),
    "circle_hex_tapered_dg1": dict(n=(5, 5, 5), p=1, hex=0.3, gamma=("network", 40, 1 / 5, 2), K=("DG", 1), op=("circle", 0.05, 7)),
    "disk_hex_tapered_p1k_firstvisit": dict(n=(5, 6, 4), p=1, hex=-0.2, gamma=("network", 40, 1 / 5, 2), K=("Lagrange", 1),
                                            op=("disk", 0.04, 3)),
    # P3 on tetrahedra (tests/test_interpolation_matrix.py:181 builds V = ("Lagrange", 3)): 20 dofs per cell, edge dofs
    # oriented by the dofmap
    "circle_p3v_quadrature": dict(n=(3, 3, 4), p=3, gamma=("network", 24, 1 / 3, 2), K=("Quadrature", 3), op=("circle", 0.07, 6)),
    "trace_p3v_p1k_firstvisit": dict(n=(3, 3, 3), p=3, gamma=("network", 20, 1 / 3, 2), K=("Lagrange", 1), op=("trace",)),
}


def callable_radius(x):
    return 0.03 + 0.05 * x[2] ** 2


def mapped_shift(x):
    """The mapping of the MappedRestriction case (vectorised: x[0] are the x coordinates, ...)."""
    return np.stack([x[0] + 0.013, x[1] - 0.007 + 0.02 * x[2], x[2] + 0.004])


def load_case(name):
    """(mesh, V, gamma, K, op_name, radius, degree, cells_K, fixture) — Γ comes from the fixture
    itself so the test does not depend on the generator's random stream."""
    spec = CASES[name]
    fx = np.load(os.path.join(GOLDEN_DIR, f"pi_{name}.npz"))
    mesh = sy.hex_box(*spec["n"], taper=spec["hex"]) if "hex" in spec else sy.kuhn_box(*spec["n"])
    bs = spec.get("bs", 1)
    if "hex" in spec:
        V = functionspace(mesh, ("Lagrange", 1))
    elif bs > 1:
        V = functionspace(mesh, ("Lagrange", spec["p"], (bs,)))
    else:
        V = functionspace(mesh, ("Lagrange", spec["p"])) if spec["p"] in (1, 3) else sy.kuhn_p2_space(mesh)
    from fenicsx_ii_b200.mesh import Mesh

    gamma = Mesh(fx["gamma_x"], fx["gamma_cells"], tdim=1, name="gamma")
    K = functionspace(gamma, spec["K"] + ((bs,),) if bs > 1 else spec["K"])
    op = spec["op"]
    radius = degree = None
    if op[0] not in ("trace", "mapped"):
        radius = callable_radius if op[1] == "callable" else op[1]
        degree = op[2]
    cells_K = fx["cells_K"] if bool(fx["has_cells_K"]) else None
    return mesh, V, gamma, K, op[0], radius, degree, cells_K, fx
