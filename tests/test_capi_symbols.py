"""CPU: the C-ABI library loads, exports every function include/fxii.h declares, the ctypes table
covers each of them, and the product path refuses to run without a CUDA device."""

import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "fxii.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fxii_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_surface():
    names = declared_functions()
    for must in ("fxii_mesh_create", "fxii_space_create", "fxii_rows_create", "fxii_pi_assemble", "fxii_csr_transpose",
                 "fxii_spgemm", "fxii_csr_spmv", "fxii_csr_download"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from fenicsx_ii_b200 import _lib

    lib = _lib.load()
    raw = ctypes.CDLL(os.fspath(_lib.lib_path()))
    for name in declared_functions():
        assert hasattr(raw, name), f"libfxii.so does not export {name}"
        assert name in _lib.SIGNATURES, f"_lib.SIGNATURES lacks {name}"
    assert lib.fxii_version() >= 100


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import numpy as np

    from fenicsx_ii_b200 import PointwiseTrace, create_interpolation_matrix
    from fenicsx_ii_b200 import synthetic as sy
    from fenicsx_ii_b200._lib import FxiiError
    from fenicsx_ii_b200.mesh import functionspace

    mesh = sy.kuhn_box(2, 2, 2)
    V = functionspace(mesh, ("Lagrange", 1))
    gamma = sy.straight_line(3)
    K = functionspace(gamma, ("DG", 1))
    with pytest.raises(FxiiError):
        create_interpolation_matrix(V, K, PointwiseTrace(gamma))
    assert np.all(np.isfinite(mesh.geometry.x))


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under fenicsx_ii_b200/ may reference it."""
    pkg = os.path.join(ROOT, "fenicsx_ii_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
