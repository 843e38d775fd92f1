"""ctypes binding of the C ABI declared in include/fxii.h.

There is no CPU fallback: if ``libfxii.so`` is missing it is built with nvcc; if that fails, or if
a call is made without a CUDA device, the error is raised to the caller.
"""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
# FXII_LIB: load a differently tuned build of the same ABI (kernel A/B experiments)
_LIB_PATH = Path(os.environ["FXII_LIB"]) if os.environ.get("FXII_LIB") else _HERE / "csrc" / "libfxii.so"

FXII_OK = 0
FXII_E_INVALID = -1
FXII_E_CUDA = -2
FXII_E_NOT_FOUND = -3
FXII_E_UNSUPPORTED = -4

OP_TRACE, OP_CIRCLE, OP_DISK, OP_POINTS = 0, 1, 2, 3


class FxiiError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"fxii error {code}: {message}")
        self.code = code


class PointsNotFound(FxiiError, AssertionError):
    """Some points lie in no cell of the 3D mesh (the reference asserts at interpolation.py:78)."""


class PiStats(C.Structure):
    _fields_ = [
        ("n_points", C.c_int64),
        ("n_entries", C.c_int64),
        ("n_not_found", C.c_int64),
        ("n_ties", C.c_int64),
        ("n_fallback", C.c_int64),
        ("ms_frames", C.c_float),
        ("ms_locate", C.c_float),
        ("ms_csr", C.c_float),
        ("ms_csr_count", C.c_float),
        ("ms_csr_fill", C.c_float),
        ("fused", C.c_int32),
    ]


class RowsDesc(C.Structure):
    """fxii_rows_desc (include/fxii.h)."""

    _fields_ = [
        ("gx", C.c_void_p), ("n_gnodes", C.c_int64),
        ("gcells", C.c_void_p), ("n_gcells", C.c_int64),
        ("cells_k", C.c_void_p), ("n_cells_k", C.c_int64),
        ("xref", C.c_void_p), ("nip", C.c_int32),
        ("row_dofs", C.c_void_p), ("n_rows_total", C.c_int64),
        ("kind", C.c_int32), ("nq", C.c_int32),
        ("table", C.c_void_p), ("w", C.c_void_p), ("radius", C.c_void_p),
        ("radius_per_row", C.c_int32),
    ]


_vp = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int32
_dbl = C.c_double

# name -> (restype, argtypes); must list every symbol of include/fxii.h (tests check this)
SIGNATURES = {
    "fxii_version": (C.c_int, []),
    "fxii_last_error": (C.c_char_p, []),
    "fxii_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "fxii_set_device": (C.c_int, [C.c_int]),
    "fxii_sm_count": (C.c_int, [C.POINTER(C.c_int)]),
    "fxii_launch_count": (C.c_int64, []),
    "fxii_trim_memory": (C.c_int, []),
    "fxii_mesh_create": (C.c_int, [_vp, _i64, _vp, _i64, _dbl, _vp, C.POINTER(_vp)]),
    "fxii_mesh_create_cells": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _dbl, _vp, C.POINTER(_vp)]),
    "fxii_mesh_info": (C.c_int, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i32)]),
    "fxii_mesh_destroy": (C.c_int, [_vp]),
    "fxii_space_create": (C.c_int, [_vp, _vp, _i32, _i32, _i64, _vp, C.POINTER(_vp)]),
    "fxii_space_destroy": (C.c_int, [_vp]),
    "fxii_rows_create": (
        C.c_int,
        [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _i32, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _i32, _vp, C.POINTER(_vp)],
    ),
    "fxii_rows_create_points": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _vp, _i64, _vp, C.POINTER(_vp)]),
    "fxii_rows_destroy": (C.c_int, [_vp]),
    "fxii_rows_keep_points": (C.c_int, [_vp, _i32]),
    "fxii_rows_set_mode": (C.c_int, [_vp, _i32]),
    "fxii_rows_set_nnz_hint": (C.c_int, [_vp, _i64]),
    "fxii_rows_get_nnz_hint": (C.c_int64, [_vp]),
    "fxii_pi_assemble": (C.c_int, [_vp, _vp, _vp, C.POINTER(_vp), C.POINTER(PiStats)]),
    "fxii_pi_assemble_host": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, C.POINTER(_i64), C.POINTER(PiStats)]),
    "fxii_interpolation_matrix_host": (C.c_int, [_vp, C.POINTER(RowsDesc), _i64, _vp, _vp, _vp, _vp, _i64, C.POINTER(_i64),
                                               C.POINTER(PiStats), C.POINTER(_vp)]),
    "fxii_rows_download_diag": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "fxii_rows_download_coo": (C.c_int, [_vp, _vp, _vp, _vp]),
    "fxii_csr_upload": (C.c_int, [_i64, _i64, _i64, _vp, _vp, _vp, _vp, C.POINTER(_vp)]),
    "fxii_csr_from_device": (C.c_int, [_i64, _i64, _i64, _vp, _vp, _vp, _vp, C.POINTER(_vp)]),
    "fxii_device_alloc": (C.c_int, [_i64, _vp, C.POINTER(_vp)]),
    "fxii_device_free": (C.c_int, [_vp, _vp]),
    "fxii_alloc_base": (C.c_int, [_vp, C.POINTER(_vp)]),
    "fxii_ipc_export": (C.c_int, [_vp, _vp]),
    "fxii_ipc_open": (C.c_int, [_vp, C.POINTER(_vp)]),
    "fxii_ipc_close": (C.c_int, [_vp]),
    "fxii_copy_dev_async": (C.c_int, [_vp, _vp, _i64, _vp]),
    "fxii_copy_segments_async": (C.c_int, [_i32, _vp, _vp, _vp, _vp, _vp]),
    "fxii_csr_info": (C.c_int, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64)]),
    "fxii_csr_device_ptrs": (C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "fxii_csr_download": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "fxii_csr_download_async": (C.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "fxii_csr_destroy": (C.c_int, [_vp]),
    "fxii_csr_transpose": (C.c_int, [_vp, _vp, C.POINTER(_vp)]),
    "fxii_csr_transpose_window": (C.c_int, [_vp, _i64, _i64, _vp, C.POINTER(_vp)]),
    "fxii_spgemm": (C.c_int, [_vp, _vp, _i64, _i64, _vp, C.POINTER(_vp), C.POINTER(_i64)]),
    "fxii_spgemm_scaled": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp, C.POINTER(_vp), C.POINTER(_i64)]),
    "fxii_csr_wrap_device": (C.c_int, [_i64, _i64, _i64, _vp, _vp, _vp, C.POINTER(_vp)]),
    "fxii_csr_validate": (C.c_int, [_vp, _vp]),
    "fxii_csr_expand_blocks": (C.c_int, [_vp, _i32, _vp, C.POINTER(_vp)]),
    "fxii_csr_scale_rows": (C.c_int, [_vp, _vp, _vp]),
    "fxii_csr_spmv": (C.c_int, [_vp, _vp, _vp, _vp]),
    "fxii_csr_spmv_t": (C.c_int, [_vp, _vp, _vp, _vp]),
}

_lib = None


def lib_path() -> Path:
    return _LIB_PATH


def load(build_if_missing: bool = True):
    """Load libfxii.so (building it in-tree first if absent).  Raises if neither is possible."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        if not build_if_missing:
            raise FxiiError(FXII_E_INVALID, f"{_LIB_PATH} is missing; run `python -m fenicsx_ii_b200.csrc.build`")
        from .csrc.build import build

        build()
    lib = C.CDLL(os.fspath(_LIB_PATH))
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export the symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc: int):
    if rc == FXII_OK:
        return
    msg = load().fxii_last_error().decode("utf-8", "replace")
    if rc == FXII_E_NOT_FOUND:
        raise PointsNotFound(rc, msg)
    raise FxiiError(rc, msg)


def ptr(a: np.ndarray | None):
    """Host pointer of a C-contiguous array (None -> NULL)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def current_stream() -> int:
    """cudaStream_t of torch's current stream, so torch CUDA events bracket our kernels."""
    import torch

    # torch.cuda.current_stream() costs ~15 us of Python; the raw query is ~1 us
    return int(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))


def require_cuda():
    """The product path has no CPU fallback: fail loudly when there is no CUDA device."""
    lib = load()
    n = C.c_int(0)
    rc = lib.fxii_device_count(C.byref(n))
    if rc != FXII_OK or n.value < 1:
        raise FxiiError(FXII_E_CUDA, "no CUDA device available: fenicsx_ii_b200 runs only on a GPU (no CPU fallback)")
    return n.value
