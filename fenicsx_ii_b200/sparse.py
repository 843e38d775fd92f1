"""Device-resident CSR matrix handle and the block products on it.

``DeviceCSR`` wraps an ``fxii_csr*``.  Products replace the PETSc calls of the reference's
``apply_matrix_replacer`` (matrix_assembler.py:175-258): ``transpose()`` = ``Mat.transpose``,
``matmult()`` = ``Mat.matMult``, ``mult()`` = ``Mat.mult``.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class DeviceCSR:
    def __init__(self, handle: int):
        self._h = C.c_void_p(handle)
        n_rows, n_cols, nnz = C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(_lib.load().fxii_csr_info(self._h, C.byref(n_rows), C.byref(n_cols), C.byref(nnz)))
        self.shape = (n_rows.value, n_cols.value)
        self.nnz = nnz.value

    # -- construction ----------------------------------------------------------------------------
    @classmethod
    def from_host(cls, indptr, indices, data, shape, stream: int | None = None) -> "DeviceCSR":
        _lib.require_cuda()
        indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        indices = np.ascontiguousarray(indices, dtype=np.int32)
        data = np.ascontiguousarray(data, dtype=np.float64)
        assert indptr.shape[0] == shape[0] + 1 and indices.shape[0] == data.shape[0] == indptr[-1]
        out = C.c_void_p()
        s = _lib.current_stream() if stream is None else stream
        _lib.check(_lib.load().fxii_csr_upload(shape[0], shape[1], int(indptr[-1]), _lib.ptr(indptr), _lib.ptr(indices),
                                               _lib.ptr(data), s, C.byref(out)))
        return cls(out.value)

    @classmethod
    def from_scipy(cls, mat) -> "DeviceCSR":
        """Upload a scipy matrix.  Duplicate entries are summed and columns sorted first (the products assume
        PETSc's assembled-AIJ form: unique ascending columns); explicit zeros are kept."""
        mat = mat.tocsr()
        if not mat.has_canonical_format:
            mat = mat.copy()
            mat.sum_duplicates()
        return cls.from_host(mat.indptr, mat.indices, mat.data, mat.shape)

    @classmethod
    def wrap_device_arrays(cls, indptr_t, indices_t, values_t, shape) -> "DeviceCSR":
        """A matrix over torch CUDA tensors WITHOUT copying them (the gathered row shards of the multi-GPU
        product); the tensors are kept alive by the returned object."""
        import torch

        assert indptr_t.dtype == torch.int64 and indices_t.dtype == torch.int32 and values_t.dtype == torch.float64
        assert indptr_t.is_contiguous() and indices_t.is_contiguous() and values_t.is_contiguous()
        assert indptr_t.numel() == shape[0] + 1 and indices_t.numel() == values_t.numel()
        out = C.c_void_p()
        _lib.check(_lib.load().fxii_csr_wrap_device(shape[0], shape[1], int(indices_t.numel()), indptr_t.data_ptr(),
                                                    indices_t.data_ptr(), values_t.data_ptr(), C.byref(out)))
        m = cls(out.value)
        m._keepalive = (indptr_t, indices_t, values_t)
        return m

    def row_slice(self, r0: int, r1: int) -> "DeviceCSR":
        """Rows [r0, r1) as a matrix of their own, sharing this matrix's index and value arrays (only the row
        pointers are copied).  One small host read (the entry offsets of the two boundary rows)."""
        ptr, idx, val = self.device_tensors()
        e0, e1 = ptr[[r0, r1]].tolist()
        sub = DeviceCSR.wrap_device_arrays((ptr[r0:r1 + 1] - e0).contiguous(), idx[e0:e1], val[e0:e1], (r1 - r0, self.shape[1]))
        sub._keepalive = sub._keepalive + (self,)
        return sub

    @classmethod
    def vstack(cls, blocks, entry_offsets=None) -> "DeviceCSR":
        """Row-wise concatenation of matrices with the same number of columns, one copy kernel for all arrays (no host
        synchronisation).  `entry_offsets[i]`: what to add to block i's row pointers on the way — default: the entries
        of the blocks before it; pass zeros for blocks whose pointers are already in the concatenated numbering."""
        import torch

        blocks = list(blocks)
        assert blocks and all(b.shape[1] == blocks[0].shape[1] for b in blocks)
        assert 3 * len(blocks) <= 32, "at most 10 blocks per concatenation"
        n_rows, nnz = sum(b.shape[0] for b in blocks), sum(b.nnz for b in blocks)
        dev = torch.device("cuda", torch.cuda.current_device())
        ptr = torch.empty(n_rows + 1, dtype=torch.int64, device=dev)
        ptr[:1] = 0
        idx = torch.empty(nnz, dtype=torch.int32, device=dev)
        val = torch.empty(nnz, dtype=torch.float64, device=dev)
        dst, src, nbytes, add = [], [], [], []
        r = e = 0
        for i, b in enumerate(blocks):
            p0, p1, p2 = C.c_void_p(), C.c_void_p(), C.c_void_p()
            _lib.check(_lib.load().fxii_csr_device_ptrs(b._h, C.byref(p0), C.byref(p1), C.byref(p2)))
            shift = e if entry_offsets is None else int(entry_offsets[i])
            if b.shape[0]:
                dst.append(ptr.data_ptr() + 8 * (1 + r)), src.append((p0.value or 0) + 8), nbytes.append(8 * b.shape[0]), add.append(shift)
            if b.nnz:
                dst.append(idx.data_ptr() + 4 * e), src.append(p1.value or 0), nbytes.append(4 * b.nnz), add.append(0)
                dst.append(val.data_ptr() + 8 * e), src.append(p2.value or 0), nbytes.append(8 * b.nnz), add.append(0)
            r += b.shape[0]
            e += b.nnz
        k = len(dst)
        if k:
            _lib.check(_lib.load().fxii_copy_segments_async(k, (C.c_void_p * k)(*dst), (C.c_void_p * k)(*src), (C.c_int64 * k)(*nbytes),
                                                            (C.c_int64 * k)(*add), _lib.current_stream()))
        out = cls.wrap_device_arrays(ptr, idx, val, (n_rows, blocks[0].shape[1]))
        out._keepalive = out._keepalive + tuple(blocks)  # the copy is stream-ordered: sources stay alive with the result
        return out

    def validate(self) -> "DeviceCSR":
        """Raise unless indptr is monotone, columns are in range and strictly ascending in every row."""
        _lib.check(_lib.load().fxii_csr_validate(self._h, _lib.current_stream()))
        return self

    @classmethod
    def from_device_arrays(cls, indptr_t, indices_t, values_t, shape) -> "DeviceCSR":
        """Copy torch CUDA tensors (int64 / int32 / float64) into a new matrix, e.g. after the
        NCCL all-gather of row shards."""
        out = C.c_void_p()
        nnz = int(indices_t.numel())
        _lib.check(_lib.load().fxii_csr_from_device(shape[0], shape[1], nnz, indptr_t.data_ptr(), indices_t.data_ptr(),
                                                    values_t.data_ptr(), _lib.current_stream(), C.byref(out)))
        return cls(out.value)

    # -- access ----------------------------------------------------------------------------------
    def to_host(self):
        """(indptr int64, indices int32, data float64) numpy arrays."""
        indptr = np.empty(self.shape[0] + 1, dtype=np.int64)
        indices = np.empty(self.nnz, dtype=np.int32)
        data = np.empty(self.nnz, dtype=np.float64)
        _lib.check(_lib.load().fxii_csr_download(self._h, _lib.ptr(indptr), _lib.ptr(indices), _lib.ptr(data),
                                                 _lib.current_stream()))
        return indptr, indices, data

    def to_host_into(self, indptr, indices, data, sync: bool = True):
        """Download into caller-provided (e.g. pinned) buffers on torch's current stream.  ``sync=False``: the copies
        are only enqueued — synchronise the stream before reading the buffers and keep this matrix alive until then
        (with pinned buffers on a side stream the download overlaps the kernels of the main stream)."""
        fn = _lib.load().fxii_csr_download if sync else _lib.load().fxii_csr_download_async
        _lib.check(fn(self._h, indptr.ctypes.data, indices.ctypes.data, data.ctypes.data, _lib.current_stream()))

    def to_scipy(self):
        import scipy.sparse as sp

        indptr, indices, data = self.to_host()
        return sp.csr_matrix((data, indices, indptr), shape=self.shape)

    def device_tensors(self):
        """Zero-copy torch views (indptr, indices, values) of the device arrays; valid while this
        object lives."""
        import torch

        p0, p1, p2 = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _lib.check(_lib.load().fxii_csr_device_ptrs(self._h, C.byref(p0), C.byref(p1), C.byref(p2)))
        dev = torch.cuda.current_device()

        def view(ptr, n, typestr, dtype):
            if n == 0 or not ptr:
                return torch.empty(0, dtype=dtype, device=f"cuda:{dev}")

            class _Arr:
                __cuda_array_interface__ = dict(shape=(n,), typestr=typestr, data=(ptr, False), version=3, strides=None)

            holder = _Arr()
            t = torch.as_tensor(holder, device=f"cuda:{dev}")
            t._fxii_owner = self  # keep the matrix alive
            return t

        return (
            view(p0.value, self.shape[0] + 1, "<i8", torch.int64),
            view(p1.value, self.nnz, "<i4", torch.int32),
            view(p2.value, self.nnz, "<f8", torch.float64),
        )

    # -- products --------------------------------------------------------------------------------
    def transpose(self, col_begin: int = 0, col_end: int = -1) -> "DeviceCSR":
        """Aᵀ, or only its rows [col_begin, col_end) (the columns of A in that window) — the V-dof row
        shard a rank needs for the multi-GPU product."""
        out = C.c_void_p()
        if col_begin == 0 and col_end < 0:
            _lib.check(_lib.load().fxii_csr_transpose(self._h, _lib.current_stream(), C.byref(out)))
        else:
            _lib.check(_lib.load().fxii_csr_transpose_window(self._h, col_begin, col_end, _lib.current_stream(), C.byref(out)))
        return DeviceCSR(out.value)

    def matmult(self, other: "DeviceCSR", row_begin: int = 0, row_end: int = -1, return_ip: bool = False, row_scale=None):
        """``self[row_begin:row_end] @ other`` (Mat.matMult).  ``row_scale``: float64 torch CUDA tensor d of
        ``other.shape[0]`` entries — computes ``self @ diag(d) @ other`` without materialising ``diag(d) @ other``
        (the diagonal quadrature mass of Πᵀ(M_ΓΠ))."""
        out = C.c_void_p()
        ip = C.c_int64()
        if row_scale is None:
            _lib.check(_lib.load().fxii_spgemm(self._h, other._h, row_begin, row_end, _lib.current_stream(), C.byref(out),
                                               C.byref(ip)))
        else:
            import torch

            assert row_scale.is_cuda and row_scale.dtype == torch.float64 and row_scale.is_contiguous()
            assert row_scale.numel() == other.shape[0]
            _lib.check(_lib.load().fxii_spgemm_scaled(self._h, other._h, C.c_void_p(row_scale.data_ptr()), row_begin, row_end,
                                                      _lib.current_stream(), C.byref(out), C.byref(ip)))
        c = DeviceCSR(out.value)
        return (c, ip.value) if return_ip else c

    def expand_blocks(self, bs: int) -> "DeviceCSR":
        """Π of blocked spaces from the scalar Π: bs x bs blocks, value on the diagonal, explicit zeros elsewhere."""
        out = C.c_void_p()
        _lib.check(_lib.load().fxii_csr_expand_blocks(self._h, int(bs), _lib.current_stream(), C.byref(out)))
        return DeviceCSR(out.value)

    def scale_rows(self, d) -> "DeviceCSR":
        """diag(d)·A in place; `d`: numpy array, or a float64 torch CUDA tensor (used in place, no host round trip)."""
        if hasattr(d, "data_ptr") and getattr(d, "is_cuda", False):
            import torch

            assert d.dtype == torch.float64 and d.is_contiguous() and d.numel() == self.shape[0]
            _lib.check(_lib.load().fxii_csr_scale_rows(self._h, C.c_void_p(d.data_ptr()), _lib.current_stream()))
            return self
        d = np.ascontiguousarray(d, dtype=np.float64)
        assert d.shape == (self.shape[0],)
        _lib.check(_lib.load().fxii_csr_scale_rows(self._h, _lib.ptr(d), _lib.current_stream()))
        return self

    def mult(self, u) -> np.ndarray:
        u = np.ascontiguousarray(u, dtype=np.float64)
        assert u.shape == (self.shape[1],)
        y = np.empty(self.shape[0], dtype=np.float64)
        _lib.check(_lib.load().fxii_csr_spmv(self._h, _lib.ptr(u), _lib.ptr(y), _lib.current_stream()))
        return y

    def mult_transpose(self, u) -> np.ndarray:
        u = np.ascontiguousarray(u, dtype=np.float64)
        assert u.shape == (self.shape[0],)
        y = np.empty(self.shape[1], dtype=np.float64)
        _lib.check(_lib.load().fxii_csr_spmv_t(self._h, _lib.ptr(u), _lib.ptr(y), _lib.current_stream()))
        return y

    def add_scaled(self, alpha: float, other: "DeviceCSR") -> "DeviceCSR":
        """``self + alpha * other`` as a new matrix whose pattern is the UNION of both patterns, columns sorted
        (PETSc ``MatAXPY`` with ``DIFFERENT_NONZERO_PATTERN``; explicit zeros survive).  Computed on the device as
        ``[I  alpha I] @ [self; other]`` with the SpGEMM kernels, so every entry is ``1*a + alpha*b`` in that order."""
        import torch

        assert self.shape == other.shape
        m = self.shape[0]
        p0, i0, v0 = self.device_tensors()
        p1, i1, v1 = other.device_tensors()
        stacked = DeviceCSR.wrap_device_arrays(torch.cat([p0, p1[1:] + self.nnz]), torch.cat([i0, i1]), torch.cat([v0, v1]),
                                               (2 * m, self.shape[1]))
        dev = p0.device
        rows = torch.arange(m, dtype=torch.int32, device=dev)
        sel = DeviceCSR.wrap_device_arrays(torch.arange(m + 1, dtype=torch.int64, device=dev) * 2,
                                           torch.stack([rows, rows + m], dim=1).reshape(-1).contiguous(),
                                           torch.tensor([1.0, float(alpha)], dtype=torch.float64, device=dev).repeat(m), (m, 2 * m))
        return sel.matmult(stacked)

    def copy(self) -> "DeviceCSR":
        a, b, c = self.device_tensors()
        return DeviceCSR.from_device_arrays(a, b, c, self.shape)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.load().fxii_csr_destroy(h)
            except Exception:
                pass


class MatrixCSR:
    """What ``create_interpolation_matrix(..., use_petsc=False)`` returns: the surface of
    ``dolfinx.la.MatrixCSR`` the callers use (``indptr``, ``indices``, ``data``, ``to_scipy``,
    ``to_dense``) over a device-resident matrix.  Host arrays are downloaded on first access."""

    def __init__(self, device: "DeviceCSR | None", stats=None, rows=None, host=None, shape=None, dtype=np.float64):
        self.dtype = np.dtype(dtype)  # scalar type of `data` on the host (the device arithmetic is always float64)
        self._device = device
        self.stats = stats
        self.rows = rows  # DeviceRows of the build (per-point diagnostics), may be None
        self._host = host  # (indptr, indices, data) when the build delivered host arrays directly
        self._shape = tuple(shape) if shape is not None else None
        assert device is not None or (host is not None and shape is not None)

    @property
    def device(self) -> DeviceCSR:
        """The device-resident matrix (uploaded on first use when the build delivered host arrays)."""
        if self._device is None:
            a, b, c = self._host
            self._device = DeviceCSR.from_host(a, b, c, self._shape)
        return self._device

    def _arrays(self):
        if self._host is None:
            self._host = self._device.to_host()
        return self._host

    @property
    def shape(self):
        return self._shape if self._device is None else self._device.shape

    @property
    def indptr(self):
        return self._arrays()[0]

    @property
    def indices(self):
        return self._arrays()[1]

    @property
    def data(self):
        """Values in the matrix' scalar type: float64 as computed, or cast on the host (complex128 with a zero imaginary
        part for ``complex_dtype=True``; the mesh's float32 for single-precision meshes — computed in float64, rounded once)."""
        v = self._arrays()[2]
        return v if self.dtype == np.float64 else v.astype(self.dtype)

    def to_scipy(self):
        import scipy.sparse as sp

        a, b, _ = self._arrays()
        return sp.csr_matrix((self.data, b, a), shape=self.shape)

    def to_dense(self):
        return self.to_scipy().toarray()

    def mult(self, u):
        u = np.asarray(u)
        if np.iscomplexobj(u):  # a real matrix acting on a complex vector: real and imaginary parts separately
            return self.device.mult(np.ascontiguousarray(u.real)) + 1j * self.device.mult(np.ascontiguousarray(u.imag))
        return self.device.mult(u)

    def scatter_reverse(self):  # serial: nothing to accumulate
        return None

    def axpy(self, alpha: float, other) -> "MatrixCSR":
        """``self += alpha * other`` (what ``apply_matrix_replacer`` does with PETSc ``Mat.axpy`` to sum the parts of a
        form, matrix_assembler.py:172-258), on the device; the pattern becomes the union of both."""
        o = other.device if isinstance(other, MatrixCSR) else other
        self._device = self.device.add_scaled(alpha, o)
        self._host = None
        self.rows = None
        return self

    def to_petsc(self):
        """PETSc AIJ with the CSR handed over in one call (instead of per-entry setValuesLocal,
        interpolation.py:182-188,243-248)."""
        try:
            from petsc4py import PETSc
        except ImportError as e:
            raise RuntimeError("use_petsc=True needs petsc4py, which is not installed") from e
        a, b, _ = self._arrays()
        c = self.data
        if b.shape[0] >= 2**31 and PETSc.IntType().itemsize == 4:
            raise RuntimeError("matrix needs 64-bit PETSc indices")
        A = PETSc.Mat().createAIJ(size=self.shape, csr=(a.astype(PETSc.IntType), b.astype(PETSc.IntType), c))
        A.assemble()
        return A
