"""CPU: the PETSc branches (use_petsc=True -> MatrixCSR.to_petsc, assign_LG_map) against a recording stand-in for
petsc4py (tests/fake_petsc4py): the CSR is handed over in ONE createAIJ(size=, csr=) call with PETSc.IntType indices,
assembled, and the LG maps are created from the index maps exactly as assembly.py:19-44 does."""

import os
import sys

import numpy as np
import pytest

FAKE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fake_petsc4py")


@pytest.fixture
def fake_petsc(monkeypatch):
    for name in [m for m in sys.modules if m == "petsc4py" or m.startswith("petsc4py.")]:
        monkeypatch.delitem(sys.modules, name)
    monkeypatch.syspath_prepend(FAKE)
    from petsc4py import PETSc

    PETSc.CALLS.clear()
    yield PETSc
    for name in [m for m in sys.modules if m == "petsc4py" or m.startswith("petsc4py.")]:
        sys.modules.pop(name, None)


def test_to_petsc_hands_the_csr_over_in_one_call(fake_petsc):
    from fenicsx_ii_b200.sparse import MatrixCSR

    indptr = np.array([0, 2, 2, 5], dtype=np.int64)
    indices = np.array([1, 3, 0, 2, 3], dtype=np.int32)
    data = np.array([1.0, 0.0, 2.0, 3.0, 4.0])  # an explicit zero survives
    A = MatrixCSR(None, host=(indptr, indices, data), shape=(3, 4)).to_petsc()
    assert isinstance(A, fake_petsc.Mat) and A.assembled
    (name, kw), = fake_petsc.CALLS
    assert name == "Mat.createAIJ" and kw["size"] == (3, 4) and kw["nnz"] == 5
    assert kw["indptr_dtype"] == fake_petsc.IntType and kw["indices_dtype"] == fake_petsc.IntType
    ip, ij, iv = A.getValuesCSR()
    np.testing.assert_array_equal(ip, indptr)
    np.testing.assert_array_equal(ij, indices)
    np.testing.assert_array_equal(iv, data)


def test_assign_lg_map_follows_the_reference(fake_petsc):
    from fenicsx_ii_b200 import assign_LG_map
    from fenicsx_ii_b200.mesh import IndexMap

    C = fake_petsc.Mat().createAIJ(size=(6, 8), csr=(np.zeros(7, dtype=np.int32), np.zeros(0, dtype=np.int32), np.zeros(0)))
    fake_petsc.CALLS.clear()
    assign_LG_map(C, IndexMap(3), IndexMap(4), 2, 2)
    names = [c[0] for c in fake_petsc.CALLS]
    assert names == ["LGMap.create", "LGMap.create", "Mat.setLGMap"]
    assert fake_petsc.CALLS[0][1] == dict(n=3, bsize=2, dtype=np.dtype(fake_petsc.IntType))
    assert fake_petsc.CALLS[1][1] == dict(n=4, bsize=2, dtype=np.dtype(fake_petsc.IntType))
    assert fake_petsc.CALLS[2][1] == dict(rows=3, cols=4, rbs=2, cbs=2)
    np.testing.assert_array_equal(C.lgmap[0].indices, np.arange(3))


def test_to_petsc_without_petsc4py_is_an_error():
    from fenicsx_ii_b200.sparse import MatrixCSR

    for name in [m for m in sys.modules if m == "petsc4py" or m.startswith("petsc4py.")]:
        sys.modules.pop(name, None)
    A = MatrixCSR(None, host=(np.zeros(2, dtype=np.int64), np.zeros(0, dtype=np.int32), np.zeros(0)), shape=(1, 1))
    with pytest.raises(RuntimeError):
        A.to_petsc()
