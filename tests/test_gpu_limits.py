"""GPU: the limits of the path fail LOUDLY: a traversal that runs out of stack, a matrix row of the
general path that collects more COO entries than shared memory holds (sorted in global memory instead), index arrays out
of range, large pageable copies, moved meshes."""

import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys
sys.path.insert(0, %(root)r)
import numpy as np
from fenicsx_ii_b200 import PointwiseTrace, create_interpolation_matrix
from fenicsx_ii_b200 import synthetic as sy
from fenicsx_ii_b200._lib import FXII_E_UNSUPPORTED, FxiiError
from fenicsx_ii_b200.mesh import functionspace
mesh = sy.kuhn_box(8, 8, 8)
V = functionspace(mesh, ("Lagrange", 1))
gamma = sy.straight_line(9)   # on mesh edges: every point collides with many cells, the walk needs its stack
K = functionspace(gamma, ("DG", 1))
for materialize in (False, True):
    try:
        create_interpolation_matrix(V, K, PointwiseTrace(gamma), materialize_coo=materialize)
    except FxiiError as e:
        assert e.code == FXII_E_UNSUPPORTED and "stack" in str(e), e
    else:
        raise SystemExit("a traversal with a 2-entry stack did not report the overflow")
print("overflow reported")
'''


def test_traversal_stack_overflow_is_reported(cuda, tmp_path):
    """A variant of the library with a 2-entry traversal stack (-DFXII_KSTACK=2): the walk cannot hold the siblings of
    its path, and the build returns FXII_E_UNSUPPORTED instead of silently dropping subtrees."""
    from fenicsx_ii_b200.csrc.build import build_variant

    lib = build_variant(tmp_path / "libfxii_kstack2.so", "pi_assemble.cu", ["-DFXII_KSTACK=2"])
    env = dict(os.environ, FXII_LIB=str(lib))
    out = subprocess.run([sys.executable, "-c", SCRIPT % dict(root=ROOT)], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "overflow reported" in out.stdout, out.stdout + out.stderr


def test_general_path_rows_beyond_the_shared_memory_budget(cuda):
    """A K dof shared by so many point rows that its matrix row collects more than 16384 COO entries (the shared-memory
    sort budget of the general path): the row is sorted in a global-memory buffer instead — same pattern, same
    first-visit-wins values as the oracle; the short rows of the same matrix keep the shared-memory kernels."""
    from fenicsx_ii_b200 import Disk, create_interpolation_matrix
    from fenicsx_ii_b200 import synthetic as sy
    from fenicsx_ii_b200.mesh import FunctionSpace, DofMap, Element, functionspace
    from tests.helpers import assert_csr_parity, oracle_pi

    mesh = sy.kuhn_box(4, 4, 4)
    V = functionspace(mesh, ("Lagrange", 1))
    gamma = sy.polyline_network(120, 0.05, n_lines=1, lo=0.3, hi=0.7)
    base = functionspace(gamma, ("DG", 0))
    # dof 0 for the first 100 cells: 100 point rows x 49 disk points x 4 dofs = 19600 entries in one matrix row;
    # the other 20 cells keep dofs of their own (short rows)
    dofs = np.zeros((120, 1), dtype=np.int32)
    dofs[100:, 0] = np.arange(1, 21)
    K = FunctionSpace(gamma, Element("DG", 0, base.element.interpolation_points), DofMap(dofs, 21))
    A, _, _ = create_interpolation_matrix(V, K, Disk(gamma, 0.05, degree=7))
    assert A.stats["fused"] == 0
    ref = oracle_pi(V, K, gamma, "disk", 0.05, 7)
    assert_csr_parity((A.indptr, A.indices, A.data), ref)


def test_index_arrays_are_range_checked_on_the_device(cuda):
    """Cell node indices and dof indices outside their ranges are refused (device-side check, reported as the
    ValueError the host-side numpy checks used to raise) instead of being dereferenced."""
    from fenicsx_ii_b200 import synthetic as sy
    from fenicsx_ii_b200.interpolation import DeviceMesh, DeviceSpace

    mesh = sy.kuhn_box(4, 4, 4)
    x, cells = np.asarray(mesh.geometry.x), np.asarray(mesh.geometry.dofmap).copy()
    good = DeviceMesh(x, cells, 1e-8)
    for bad_value in (-1, x.shape[0]):
        bad = cells.copy()
        bad[17, 2] = bad_value
        with pytest.raises(ValueError, match="cell node index out of range"):
            DeviceMesh(x, bad, 1e-8)
    dofmap = cells.copy()
    DeviceSpace(good, dofmap, 1, x.shape[0])
    for bad_value in (-3, x.shape[0]):
        bad = dofmap.copy()
        bad[-1, 0] = bad_value
        with pytest.raises(ValueError, match="dof index out of range"):
            DeviceSpace(good, bad, 1, x.shape[0])


def test_large_pageable_uploads_take_the_staged_path_unchanged(cuda):
    """Host arrays above 32 MiB in pageable memory are uploaded through the threaded pinned staging pipeline
    (capi.cu: staged_upload); the device copy must equal the source byte for byte — odd sizes, several chunks per
    worker, and a pinned source (plain copy) beside it."""
    import torch

    from fenicsx_ii_b200.sparse import DeviceCSR

    rng = np.random.default_rng(7)
    n_rows, nnz = 1000, (40 << 20) // 8 + 12345  # values: 40 MiB + a ragged tail; indices: 20 MiB (plain path)
    indptr = np.linspace(0, nnz, n_rows + 1).astype(np.int64)
    indices = rng.integers(0, 1 << 20, size=nnz, dtype=np.int32)
    values = rng.standard_normal(nnz)
    a = DeviceCSR.from_host(indptr, indices, values, (n_rows, 1 << 20))
    gi, gj, gv = a.to_host()
    assert np.array_equal(gi, indptr) and np.array_equal(gj, indices) and np.array_equal(gv, values)
    pinned = torch.from_numpy(values).pin_memory().numpy()
    b = DeviceCSR.from_host(indptr, indices, pinned, (n_rows, 1 << 20))
    assert np.array_equal(b.to_host()[2], values)


def test_moved_mesh_is_not_served_from_the_caches(cuda):
    """the Ω handle (LBVH) and the Π cache are keyed on a fingerprint of the coordinates — after the mesh is
    moved in place, a cached build must not be reused."""
    from fenicsx_ii_b200 import PointwiseTrace, clear_interpolation_cache, create_interpolation_matrix
    from fenicsx_ii_b200 import synthetic as sy
    from fenicsx_ii_b200._lib import PointsNotFound
    from fenicsx_ii_b200.mesh import functionspace

    clear_interpolation_cache()
    mesh = sy.kuhn_box(4, 4, 4)
    V = functionspace(mesh, ("Lagrange", 1))
    gamma = sy.straight_line(5)  # x = y = 0.5, inside the unit cube
    K = functionspace(gamma, ("DG", 0))
    op = PointwiseTrace(gamma)
    A0, _, _ = create_interpolation_matrix(V, K, op, cache=True)
    assert A0.device.nnz > 0
    mesh.geometry.x[:, 0] += 10.0  # the line now lies outside the mesh
    with pytest.raises(PointsNotFound):
        create_interpolation_matrix(V, K, op, cache=True)
    mesh.geometry.x[:, 0] -= 10.0
    A1, _, _ = create_interpolation_matrix(V, K, op, cache=True)
    assert np.array_equal(A1.indices, A0.indices) and np.array_equal(A1.data, A0.data)
    clear_interpolation_cache()
