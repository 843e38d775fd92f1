#!/usr/bin/env python
"""Generates the golden fixtures of tests/golden/ by running the REFERENCE's own code.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

* quadrature_golden.npz — outputs of the reference's `quadrature.py` (compute_disk_quadrature,
  rotation_matrix) and `restriction_operators.py` (Circle.quadrature, Disk.quadrature) imported
  unmodified, plus the Bojanov-Petrova tables held by the reference's
  `tests/test_circle_quadrature.py` (its golden vectors for this path).
* pi_*.npz — Π built by the reference's unmodified `interpolation.create_interpolation_matrix`
  running on the numpy dolfinx shim (tests/golden/dolfinx_shim): inputs + CSR + ownership.

The reference files are imported from where they lie; nothing is copied into the repo.
"""

from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(HERE, "dolfinx_shim"))
sys.path.insert(0, ROOT)

# a namespace package `fenicsx_ii` pointing at the reference sources: submodules import without
# running the reference's __init__.py (which pulls in petsc4py / the solver)
pkg = types.ModuleType("fenicsx_ii")
pkg.__path__ = [os.path.join(REF, "src", "fenicsx_ii")]
sys.modules["fenicsx_ii"] = pkg

import dolfinx  # noqa: E402  (the shim)
from fenicsx_ii import quadrature as rq  # noqa: E402
from fenicsx_ii import restriction_operators as rro  # noqa: E402
from fenicsx_ii.interpolation import create_interpolation_matrix as ref_create  # noqa: E402

from fenicsx_ii_b200 import synthetic as sy  # noqa: E402
from fenicsx_ii_b200.mesh import functionspace  # noqa: E402


def load_reference_test_tables():
    spec = importlib.util.spec_from_file_location("ref_test_circle", os.path.join(REF, "tests", "test_circle_quadrature.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def quadrature_fixture():
    out = {}
    tables = load_reference_test_tables()
    for n in range(1, 11):
        p, w = rq.compute_disk_quadrature(n)
        out[f"disk_points_{n}"] = p
        out[f"disk_weights_{n}"] = w
        out[f"table_x_{n}"] = tables.reference_x_coordinates(n)
        out[f"table_coeff_{n}"] = tables.reference_coefficients(n)
    rng = np.random.default_rng(11)
    normals = rng.normal(size=(40, 3))
    normals[0] = [0, 0, 1]
    normals[1] = [0, 0, -1]
    normals[2] = [1, 0, 0]
    normals[3] = [0, 1e-10, 1]
    normals[4] = [1e-7, 0, -1]
    out["rot_normals"] = normals
    out["rot_matrices"] = rq.rotation_matrix(normals.copy())
    line = dolfinx.mesh.Mesh(np.array([[0.0, 0, 0], [0, 0, 1]]), np.array([[0, 1]]), 1)
    x0 = rng.uniform(-1, 1, size=(25, 3))
    nn = normals[:25] / np.linalg.norm(normals[:25], axis=1)[:, None]
    out["avg_x0"] = x0
    out["avg_normals"] = nn
    for deg in (5, 15, 126):
        c = rro.Circle(line, 0.37, degree=deg)
        p, w, s = c.quadrature(x0.copy(), nn.copy())
        out[f"circle{deg}_points"], out[f"circle{deg}_weights"], out[f"circle{deg}_scales"] = p, w, s
    c = rro.Circle(line, lambda x: 0.1 + x[2] ** 2, degree=9)
    p, w, s = c.quadrature(x0.copy(), nn.copy())
    out["circle9c_points"], out["circle9c_weights"], out["circle9c_scales"] = p, w, s
    for deg in (3, 7):
        d = rro.Disk(line, 0.21, degree=deg)
        p, w, s = d.quadrature(x0.copy(), nn.copy())
        out[f"disk{deg}_points"], out[f"disk{deg}_weights"], out[f"disk{deg}_scales"] = p, w, s
    np.savez_compressed(os.path.join(HERE, "quadrature_golden.npz"), **out)
    print("quadrature_golden.npz", len(out), "arrays")


def shim_spaces(mesh, V, gamma, K):
    m3 = dolfinx.mesh.Mesh(mesh.geometry.x, mesh.geometry.dofmap, 3)
    m1 = dolfinx.mesh.Mesh(gamma.geometry.x, gamma.geometry.dofmap, 1)
    bs = int(V.dofmap.bs)
    assert bs == int(K.dofmap.bs)
    Vs = dolfinx.fem.FunctionSpace(m3, V.element.degree, V.dofmap.list, V.dofmap.index_map.size_local, bs=bs)
    Ks = dolfinx.fem.FunctionSpace(m1, K.element.degree, K.dofmap.list, K.dofmap.index_map.size_local,
                                   interpolation_points=K.element.interpolation_points, family=K.element.family, bs=bs)
    return m3, m1, Vs, Ks


CASES = {
    # name: (mesh dims, V degree, gamma builder, K element, operator, kwargs)
    "trace_line_on_edges": dict(n=(4, 4, 8), p=1, gamma=("line", 9), K=("Quadrature", 10), op=("trace",)),
    "circle_line_R02": dict(n=(4, 4, 8), p=1, gamma=("line", 9), K=("Quadrature", 10), op=("circle", 0.2, 5)),
    "disk_network_dg1": dict(n=(6, 6, 6), p=1, gamma=("network", 60, 1 / 6, 3), K=("DG", 1), op=("disk", 0.06, 3)),
    "trace_network_p1_firstvisit": dict(n=(6, 6, 6), p=1, gamma=("network", 60, 1 / 6, 3), K=("Lagrange", 1), op=("trace",)),
    "circle_network_p2k_firstvisit": dict(n=(5, 5, 5), p=1, gamma=("network", 40, 1 / 5, 2), K=("Lagrange", 2),
                                          op=("circle", 0.05, 7)),
    "circle_p2v_callable_radius": dict(n=(4, 4, 4), p=2, gamma=("network", 30, 1 / 4, 2), K=("Quadrature", 4),
                                       op=("circle", "callable", 9)),
    "disk_p2v_subset": dict(n=(4, 4, 4), p=2, gamma=("network", 30, 1 / 4, 2), K=("DG", 2), op=("disk", 0.05, 3),
                            cells_K=np.arange(3, 27, 2, dtype=np.int32)),
    "mapped_network_dg1": dict(n=(5, 5, 5), p=1, gamma=("network", 40, 1 / 5, 2), K=("DG", 1), op=("mapped",)),
    # blocked (vector) spaces, tests/test_interpolation_matrix.py:254-301: bs x bs blocks with explicit zeros
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


def build_case(spec):
    mesh = sy.hex_box(*spec["n"], taper=spec["hex"]) if "hex" in spec else sy.kuhn_box(*spec["n"])
    bs = spec.get("bs", 1)
    if "hex" in spec:
        V = functionspace(mesh, ("Lagrange", 1))
    elif bs > 1:
        V = functionspace(mesh, ("Lagrange", spec["p"], (bs,)))
    else:
        V = functionspace(mesh, ("Lagrange", spec["p"])) if spec["p"] in (1, 3) else sy.kuhn_p2_space(mesh)
    g = spec["gamma"]
    if g[0] == "line":
        gamma = sy.straight_line(g[1])
    else:
        gamma = sy.polyline_network(g[1], g[2], n_lines=g[3], lo=0.3, hi=0.7)
    K = functionspace(gamma, spec["K"] + ((bs,),) if bs > 1 else spec["K"])
    return mesh, V, gamma, K


def pi_fixtures(only=None):
    for name, spec in CASES.items():
        if only and name not in only:
            continue
        mesh, V, gamma, K = build_case(spec)
        m3, m1, Vs, Ks = shim_spaces(mesh, V, gamma, K)
        op = spec["op"]
        if op[0] == "trace":
            red = rro.PointwiseTrace(m1)
        elif op[0] == "mapped":
            from tests.golden_cases import mapped_shift

            red = rro.MappedRestriction(m1, mapped_shift)
        elif op[0] == "circle":
            red = rro.Circle(m1, callable_radius if op[1] == "callable" else op[1], degree=op[2])
        else:
            red = rro.Disk(m1, op[1], degree=op[2])
        cells_K = spec.get("cells_K")
        A, imap_K, imap_V = ref_create(Vs, Ks, red, use_petsc=False, cells_K=cells_K)
        last = m3._shim_last
        np.savez_compressed(
            os.path.join(HERE, f"pi_{name}.npz"),
            indptr=A.indptr, indices=A.indices, data=A.data, cell=last.cell, n_colliding=last.n_colliding,
            gamma_x=gamma.geometry.x, gamma_cells=gamma.geometry.dofmap,
            cells_K=np.zeros(0, dtype=np.int32) if cells_K is None else cells_K,
            has_cells_K=np.array(cells_K is not None),
        )
        print(f"pi_{name}.npz rows={A.shape[0]} nnz={A.indices.shape[0]} ties={(last.n_colliding > 1).sum()}")


if __name__ == "__main__":
    only = set(sys.argv[1:])  # case names: regenerate just these Π fixtures
    if not only:
        quadrature_fixture()
    pi_fixtures(only)
