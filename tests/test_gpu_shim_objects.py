"""GPU: objects with the DOLFINX attribute surface (the numpy shim of tests/golden/dolfinx_shim — the very objects the
reference's own interpolation.py ran on when the golden fixtures were generated, with ``mesh.comm``, ``cmap`` as an
object with ``degree``, ``SimpleNamespace`` dofmaps and elements) go straight into
``fenicsx_ii_b200.create_interpolation_matrix`` and give the fixture's matrix: the product reads nothing but the
attributes dolfinx provides."""

import os
import sys

import numpy as np
import pytest

from fenicsx_ii_b200 import Circle, Disk, PointwiseTrace, create_interpolation_matrix
from tests.golden_cases import CASES, callable_radius, load_case
from tests.helpers import assert_csr_parity

pytestmark = pytest.mark.gpu
SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dolfinx_shim")


@pytest.fixture
def shim(monkeypatch):
    monkeypatch.syspath_prepend(SHIM)
    import dolfinx  # the shim

    yield dolfinx
    for name in [m for m in sys.modules if m.split(".")[0] in ("dolfinx", "ufl", "basix", "mpi4py")]:
        sys.modules.pop(name, None)


@pytest.mark.parametrize("name", ["circle_line_R02", "disk_network_dg1", "trace_network_p1_firstvisit", "circle_p2v_callable_radius",
                                  "disk_p2v_subset", "circle_blocked2_p1k_firstvisit"])
def test_shim_dolfinx_objects_pass_through(cuda, shim, name):
    mesh, V, gamma, K, op_name, radius, degree, cells_K, fx = load_case(name)
    m3 = shim.mesh.Mesh(mesh.geometry.x, mesh.geometry.dofmap, 3)
    m1 = shim.mesh.Mesh(gamma.geometry.x, gamma.geometry.dofmap, 1)
    bs = int(V.dofmap.bs)
    Vs = shim.fem.FunctionSpace(m3, V.element.degree, V.dofmap.list, V.dofmap.index_map.size_local, bs=bs)
    Ks = shim.fem.FunctionSpace(m1, K.element.degree, K.dofmap.list, K.dofmap.index_map.size_local,
                                interpolation_points=K.element.interpolation_points, family=K.element.family, bs=bs)
    assert type(Vs).__module__.startswith("dolfinx") and not hasattr(Vs, "num_dofs")  # not this repo's mesh.FunctionSpace
    op = {"trace": lambda: PointwiseTrace(m1), "circle": lambda: Circle(m1, radius, degree=degree),
          "disk": lambda: Disk(m1, radius, degree=degree)}[op_name]()
    A, imap_K, imap_V = create_interpolation_matrix(Vs, Ks, op, cells_K=cells_K)
    assert imap_K is Ks.dofmap.index_map and imap_V is Vs.dofmap.index_map
    assert A.shape == (Ks.dofmap.index_map.size_local * bs, Vs.dofmap.index_map.size_local * bs)
    assert_csr_parity((A.indptr, A.indices, A.data), dict(indptr=fx["indptr"], indices=fx["indices"], data=fx["data"]))
    cell, _ = A.rows.diagnostics()
    np.testing.assert_array_equal(cell, fx["cell"])
