"""``create_interpolation_matrix`` — the hot entry point (reference: interpolation.py:15-251),
same signature and return triple, built on the CUDA path behind include/fxii.h.

Stage map (reference line -> here):
  :43-53   interpolation points, cells_K, red_op.compute_quadrature  -> DeviceRows (points are
           generated on the device for PointwiseTrace / Circle / Disk; generic operators run
           compute_quadrature on the host and upload the points)
  :66-68   determine_point_ownership                                 -> LBVH location kernel
  :116     evaluate_basis_function (pull_back + tabulation)          -> fused in the same kernel
  :140-163 SparsityPattern                                           -> COO -> CSR kernels
  :205-250 value insertion, first-visit-wins, assemble               -> COO -> CSR kernels
The MPI plumbing (:82-137) is the identity on a serial communicator and is not reproduced;
distributed 3D meshes raise NotImplementedError (no CPU fallback by design).
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .sparse import DeviceCSR, MatrixCSR

__all__ = ["DeviceMesh", "DeviceRows", "DeviceSpace", "PiStats", "clear_interpolation_cache", "create_interpolation_matrix",
           "device_space"]


class PiStats(dict):
    """Counters and per-stage device times of the last assembly (fxii_pi_stats)."""


def _check_indices(rc: int):
    """`_lib.check`, with the device-side index range checks reported as the ValueError the host checks used to raise."""
    try:
        _lib.check(rc)
    except _lib.FxiiError as e:
        if e.code == _lib.FXII_E_INVALID and "out of range" in str(e):
            raise ValueError(str(e)) from None
        raise


class DeviceMesh:
    """Ω on the device (affine tetrahedra or Q1 hexahedra): packed cell geometry + LBVH.  Built once per (mesh, tol) and cached on the
    mesh object — the reference rebuilds dolfinx's tree on every call (interpolation.py:66-68)."""

    def __init__(self, x: np.ndarray, cells: np.ndarray, padding: float, stream: int | None = None):
        _lib.require_cuda()
        x = np.ascontiguousarray(x, dtype=np.float64)
        cells = np.ascontiguousarray(cells, dtype=np.int32)
        if x.ndim != 2 or x.shape[1] != 3:
            raise ValueError("mesh.geometry.x must have shape (n, 3)")
        if cells.ndim != 2 or cells.shape[1] not in (4, 8):
            raise NotImplementedError("only affine tetrahedra (4 geometry nodes) and Q1 hexahedra (8 geometry nodes) are supported")
        # node indices are range-checked on the device (FXII_E_INVALID -> ValueError)
        self.n_cells, self.n_nodes, self.padding = cells.shape[0], x.shape[0], float(padding)
        out = C.c_void_p()
        s = _lib.current_stream() if stream is None else stream
        _check_indices(_lib.load().fxii_mesh_create_cells(_lib.ptr(x), x.shape[0], _lib.ptr(cells), cells.shape[0], cells.shape[1],
                                                          float(padding), s, C.byref(out)))
        self._h = out

    def info(self) -> dict:
        a, b, c, d = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int32()
        _lib.check(_lib.load().fxii_mesh_info(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return dict(n_cells=a.value, n_nodes=b.value, device_bytes=c.value, max_depth=d.value)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.load().fxii_mesh_destroy(h)
            except Exception:
                pass


class DeviceSpace:
    """V on the device: the cell dofmap next to its mesh."""

    def __init__(self, mesh: DeviceMesh, dofmap: np.ndarray, degree: int, n_dofs: int, stream: int | None = None):
        dofmap = np.ascontiguousarray(dofmap, dtype=np.int32)
        if dofmap.shape[0] != mesh.n_cells:
            raise ValueError("dofmap rows must equal the number of cells")
        # dof indices are range-checked on the device (FXII_E_INVALID -> ValueError)
        self.mesh, self.nd, self.degree, self.n_dofs = mesh, dofmap.shape[1], int(degree), int(n_dofs)
        out = C.c_void_p()
        s = _lib.current_stream() if stream is None else stream
        _check_indices(_lib.load().fxii_space_create(mesh._h, _lib.ptr(dofmap), dofmap.shape[1], int(degree), int(n_dofs), s,
                                                     C.byref(out)))
        self._h = out

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.load().fxii_space_destroy(h)
            except Exception:
                pass


class DeviceRows:
    """The Γ side of one Π build on the device (fxii_rows)."""

    def __init__(self, handle: C.c_void_p, n_point_rows: int, nq: int, n_rows_total: int):
        self._h = handle
        self.n_point_rows, self.nq, self.n_rows_total = n_point_rows, nq, n_rows_total
        self.nd = None

    @staticmethod
    def _gamma_desc(gx, gcells, cells_k, xref, row_dofs, n_rows_total, desc: dict):
        """fxii_rows_desc of a Γ-described build + the arrays it points into (keep them alive during the call).
        `cells_k` None: all cells of Γ in order; `row_dofs` None: rows 0..n_cells*nip-1 (the numbering of
        Quadrature / DG spaces).  Neither is then uploaded — the library generates them on the device."""
        gx = np.ascontiguousarray(gx, dtype=np.float64)
        gcells = np.ascontiguousarray(gcells, dtype=np.int32)
        xref = np.ascontiguousarray(np.asarray(xref, dtype=np.float64).reshape(-1))
        nip = xref.shape[0]
        if cells_k is None:
            n_cells_k = gcells.shape[0]
        else:
            cells_k = np.ascontiguousarray(cells_k, dtype=np.int32)
            n_cells_k = cells_k.shape[0]
        if row_dofs is not None:
            row_dofs = np.ascontiguousarray(np.asarray(row_dofs, dtype=np.int32).reshape(-1))
            assert row_dofs.shape[0] == n_cells_k * nip
        kind, nq = desc["kind"], int(desc["nq"])
        table = w = radius = None
        per_row = 0
        if kind != _lib.OP_TRACE:
            table = np.ascontiguousarray(desc["table"], dtype=np.float64)
            w = np.ascontiguousarray(desc["w"], dtype=np.float64)
            if "cell_radius" in desc:
                # one radius per processed cell; the rows of a cell read it on the device (no 6x host expansion)
                if cells_k is None or desc.get("cell_radius_is_local"):
                    cr = desc["cell_radius"]  # already the radii of the processed cells, in order
                else:
                    cr = desc["cell_radius"][cells_k]
                radius = np.ascontiguousarray(cr, dtype=np.float64)
                assert radius.shape[0] == n_cells_k
                per_row = 2
            else:
                radius = np.ascontiguousarray(desc["radius"], dtype=np.float64)
                per_row = 1 if desc.get("radius_per_row") else 0
                if per_row:
                    assert radius.shape[0] == n_cells_k * nip
        keep = (gx, gcells, cells_k, xref, row_dofs, table, w, radius)

        def addr(a):
            return None if a is None else a.ctypes.data

        d = _lib.RowsDesc(addr(gx), gx.shape[0], addr(gcells), gcells.shape[0], addr(cells_k), n_cells_k, addr(xref), nip,
                          addr(row_dofs), int(n_rows_total), kind, nq, addr(table), addr(w), addr(radius), per_row)
        return d, keep, n_cells_k * nip, (1 if kind == _lib.OP_TRACE else nq)

    @classmethod
    def from_gamma(cls, gx, gcells, cells_k, xref, row_dofs, n_rows_total, desc: dict, stream=None) -> "DeviceRows":
        d, keep, n_point_rows, nq = cls._gamma_desc(gx, gcells, cells_k, xref, row_dofs, n_rows_total, desc)
        out = C.c_void_p()
        s = _lib.current_stream() if stream is None else stream
        _lib.check(_lib.load().fxii_rows_create(
            d.gx, d.n_gnodes, d.gcells, d.n_gcells, d.cells_k, d.n_cells_k, d.xref, d.nip, d.row_dofs, d.n_rows_total,
            d.kind, d.nq, d.table, d.w, d.radius, d.radius_per_row, s, C.byref(out)))
        return cls(out, n_point_rows, nq, int(n_rows_total))

    @classmethod
    def build_to_host(cls, space: "DeviceSpace", gamma_args: tuple, indptr, indices, values, nnz_hint: int = 0, stream=None):
        """One C call (fxii_interpolation_matrix_host): upload Γ + operator, build Π, download the CSR into the
        caller's host arrays, one synchronisation.  `gamma_args` = the arguments of `from_gamma`.
        Returns (DeviceRows, nnz, PiStats)."""
        _lib.require_cuda()
        d, keep, n_point_rows, nq = cls._gamma_desc(*gamma_args)
        assert indptr.dtype == np.int64 and indices.dtype == np.int32 and values.dtype == np.float64
        assert indptr.shape[0] >= d.n_rows_total + 1 and indices.shape[0] == values.shape[0]
        out, st, nnz = C.c_void_p(), _lib.PiStats(), C.c_int64()
        s = _lib.current_stream() if stream is None else stream
        rc = _lib.load().fxii_interpolation_matrix_host(space._h, C.byref(d), int(nnz_hint), s, _lib.ptr(indptr), _lib.ptr(indices),
                                                        _lib.ptr(values), int(indices.shape[0]), C.byref(nnz), C.byref(st),
                                                        C.byref(out))
        del keep
        if rc == _lib.FXII_E_INVALID and nnz.value > indices.shape[0]:
            raise ValueError(f"host arrays hold {indices.shape[0]} entries, the matrix has {nnz.value}")
        if rc != _lib.FXII_OK and out.value:
            _lib.load().fxii_rows_destroy(out)
        _lib.check(rc)
        rows = cls(out, n_point_rows, nq, int(d.n_rows_total))
        rows.nd = space.nd
        return rows, int(nnz.value), PiStats({k: getattr(st, k) for k, _ in _lib.PiStats._fields_})

    @classmethod
    def from_points(cls, points, weights, scales, row_dofs, n_rows_total, stream=None) -> "DeviceRows":
        scales = np.ascontiguousarray(scales, dtype=np.float64).reshape(-1)
        nr = scales.shape[0]
        weights = np.ascontiguousarray(weights, dtype=np.float64).reshape(nr, -1)
        nq = weights.shape[1] if nr else 1
        pts = np.asarray(points, dtype=np.float64).reshape(nr * nq, -1)
        p3 = np.zeros((nr * nq, 3))
        p3[:, : pts.shape[1]] = pts  # pad to 3D (interpolation.py:56-62)
        row_dofs = np.ascontiguousarray(np.asarray(row_dofs, dtype=np.int32).reshape(-1))
        assert row_dofs.shape[0] == nr
        out = C.c_void_p()
        s = _lib.current_stream() if stream is None else stream
        _lib.check(_lib.load().fxii_rows_create_points(_lib.ptr(p3), _lib.ptr(weights), _lib.ptr(scales), nr, nq,
                                                       _lib.ptr(row_dofs), int(n_rows_total), s, C.byref(out)))
        return cls(out, nr, nq, int(n_rows_total))

    def keep_points(self, flag: bool = True):
        _lib.check(_lib.load().fxii_rows_keep_points(self._h, 1 if flag else 0))

    def set_nnz_hint(self, nnz: int):
        """Expected nnz (from an earlier build of the same Γ/operator): one host sync instead of two."""
        _lib.check(_lib.load().fxii_rows_set_nnz_hint(self._h, int(nnz)))

    def nnz_hint(self) -> int:
        return int(_lib.load().fxii_rows_get_nnz_hint(self._h))

    def assemble_to_host(self, space: DeviceSpace, indptr, indices, values, stream=None, allow_not_found: bool = False):
        """Run the Π assembly and deliver the CSR into caller-owned host arrays (int64 / int32 / float64,
        C-contiguous, ideally pinned): with an nnz hint the download rides behind the kernel, before the
        build's single host synchronisation.  Returns (nnz, PiStats)."""
        assert indptr.dtype == np.int64 and indices.dtype == np.int32 and values.dtype == np.float64
        assert indptr.shape[0] >= self.n_rows_total + 1 and indices.shape[0] == values.shape[0]
        st = _lib.PiStats()
        nnz = C.c_int64()
        s = _lib.current_stream() if stream is None else stream
        rc = _lib.load().fxii_pi_assemble_host(space._h, self._h, s, _lib.ptr(indptr), _lib.ptr(indices), _lib.ptr(values),
                                               int(indices.shape[0]), C.byref(nnz), C.byref(st))
        self.nd = space.nd
        stats = PiStats({k: getattr(st, k) for k, _ in _lib.PiStats._fields_})
        if rc == _lib.FXII_E_INVALID and nnz.value > indices.shape[0]:
            raise ValueError(f"host arrays hold {indices.shape[0]} entries, the matrix has {nnz.value}")
        if not (rc == _lib.FXII_E_NOT_FOUND and allow_not_found):
            _lib.check(rc)
        return int(nnz.value), stats

    def set_mode(self, mode: int):
        """0: fused row reduction when possible (default); 1: always materialise the COO;
        2: fused, as separate kernels + CUDA-graph replay instead of the single cooperative launch."""
        _lib.check(_lib.load().fxii_rows_set_mode(self._h, int(mode)))

    def assemble(self, space: DeviceSpace, stream=None, allow_not_found: bool = False):
        """Run the Π assembly; returns (DeviceCSR, PiStats)."""
        out = C.c_void_p()
        st = _lib.PiStats()
        s = _lib.current_stream() if stream is None else stream
        rc = _lib.load().fxii_pi_assemble(space._h, self._h, s, C.byref(out), C.byref(st))
        self.nd = space.nd
        stats = PiStats({k: getattr(st, k) for k, _ in _lib.PiStats._fields_})
        if rc == _lib.FXII_E_NOT_FOUND and allow_not_found and out.value:
            return DeviceCSR(out.value), stats
        if rc != _lib.FXII_OK and out.value:
            _lib.load().fxii_csr_destroy(out)
        _lib.check(rc)
        return DeviceCSR(out.value), stats

    def diagnostics(self, points: bool = False):
        """(owning cell i32[Np], colliding-set size u8[Np][, points f64[Np,3]])."""
        Np = self.n_point_rows * self.nq
        cell = np.empty(Np, dtype=np.int32)
        ncoll = np.empty(Np, dtype=np.uint8)
        pts = np.empty((Np, 3), dtype=np.float64) if points else None
        _lib.check(_lib.load().fxii_rows_download_diag(self._h, _lib.ptr(cell), _lib.ptr(ncoll), _lib.ptr(pts),
                                                       _lib.current_stream()))
        return (cell, ncoll, pts) if points else (cell, ncoll)

    def coo(self):
        """(cols i32[E], vals f64[E]) exactly as the locate kernel wrote them."""
        E = self.n_point_rows * self.nq * self.nd
        cols = np.empty(E, dtype=np.int32)
        vals = np.empty(E, dtype=np.float64)
        _lib.check(_lib.load().fxii_rows_download_coo(self._h, _lib.ptr(cols), _lib.ptr(vals), _lib.current_stream()))
        return cols, vals

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.load().fxii_rows_destroy(h)
            except Exception:
                pass


# ---- extraction of the arrays from (dolfinx-like) objects ---------------------------------------
def _element_degree(V) -> int:
    el = V.element
    if hasattr(el, "degree"):
        return int(el.degree)
    return int(el.basix_element.degree)  # dolfinx


def _comm_size(mesh) -> int:
    comm = getattr(mesh, "comm", None)
    return 1 if comm is None else int(getattr(comm, "size", 1))


def device_space(V, tol: float = 1.0e-8) -> DeviceSpace:
    """Device handles of V, cached on the Python objects (mesh: per tol; space: per V)."""
    mesh = V.mesh
    if _comm_size(mesh) > 1:
        raise NotImplementedError("distributed 3D meshes are not supported on the device path (no CPU fallback)")
    # one handle per tol, rebuilt when the coordinates were moved (fingerprint of ~32 sampled nodes, see _geometry_token)
    token = _geometry_token(mesh)
    cache = mesh.__dict__.setdefault("_fxii_meshes", {})
    dmesh, seen = cache.get(float(tol), (None, None))
    if dmesh is None or seen != token:
        x = np.asarray(mesh.geometry.x)
        dmesh = DeviceMesh(x, np.asarray(mesh.geometry.dofmap), tol)
        cache[float(tol)] = (dmesh, token)
    spaces = V.__dict__.setdefault("_fxii_spaces", {})
    dspace = spaces.get(float(tol))
    if dspace is None or dspace.mesh is not dmesh:
        n_dofs = V.dofmap.index_map.size_local + V.dofmap.index_map.num_ghosts
        dspace = spaces[float(tol)] = DeviceSpace(dmesh, np.asarray(V.dofmap.list), _element_degree(V), n_dofs)
    return dspace


def _uses_stock_quadrature(red_op) -> bool:
    from . import restriction_operators as ro

    if not hasattr(red_op, "device_descriptor"):
        return False
    t = type(red_op)
    if isinstance(red_op, ro.PointwiseTrace):
        return t.compute_quadrature is ro.PointwiseTrace.compute_quadrature and t.device_descriptor is ro.PointwiseTrace.device_descriptor
    if isinstance(red_op, ro._RadialAverage):
        stock = (ro.Circle, ro.Disk)
        base = next((b for b in stock if isinstance(red_op, b)), None)
        return base is not None and all(getattr(t, name) is getattr(base, name) for name in
                                        ("compute_quadrature", "quadrature", "_quadrature_R", "_weight_multiplier", "device_descriptor"))
    return False


def device_rows(K, red_op, cells_K=None, args_only: bool = False):
    """interpolation.py:41-64: interpolation points of K, cells_K, the operator's quadrature.
    `args_only`: return the arguments of ``DeviceRows.from_gamma`` instead of a handle (None when the operator
    has no device descriptor)."""
    mesh_to = K.mesh
    ref_points = np.asarray(K.element.interpolation_points, dtype=np.float64)
    assert K.element.interpolation_ident
    assert not K.element.needs_dof_transformations
    k_list = np.asarray(K.dofmap.list)
    n_cells = mesh_to.topology.index_map(mesh_to.topology.dim).size_local
    if cells_K is not None:
        cells_K = np.ascontiguousarray(cells_K, dtype=np.int32)
        if cells_K.shape[0] == n_cells and (n_cells == 0 or (cells_K[0] == 0 and cells_K[-1] == n_cells - 1
                                                             and bool(np.all(np.diff(cells_K) == 1)))):
            cells_K = None  # all cells in order
    all_cells = cells_K is None and k_list.shape[0] == n_cells
    n_rows_total = K.dofmap.index_map.size_local + K.dofmap.index_map.num_ghosts  # scalar rows; blocks are expanded later
    if k_list.shape[1] != ref_points.shape[0]:
        raise ValueError("K must have one dof per interpolation point")
    # cell-local numbering (Quadrature / DG spaces): row r of the point rows IS matrix row r; checked once per K
    identity = False
    if all_cells:
        identity = K.__dict__.get("_fxii_identity_dofs") if hasattr(K, "__dict__") else None
        if identity is None:
            flat = k_list.reshape(-1)
            identity = bool(flat.shape[0] <= n_rows_total and np.array_equal(flat, np.arange(flat.shape[0], dtype=flat.dtype)))
            if hasattr(K, "__dict__"):
                K.__dict__["_fxii_identity_dofs"] = identity
    if identity:
        row_dofs = None
    else:
        row_dofs = k_list.reshape(-1) if cells_K is None else k_list[cells_K].reshape(-1)

    def all_or(cells):
        return np.arange(n_cells, dtype=np.int32) if cells is None else cells

    # points are generated on the device only for the stock operators: a subclass that overrides the plugin contract
    # (compute_quadrature / quadrature) is a generic operator and goes through compute_quadrature on the host
    on_device = _uses_stock_quadrature(red_op) and red_op._mesh is mesh_to and mesh_to.topology.dim == 1 \
        and np.asarray(mesh_to.geometry.dofmap).shape[1] == 2
    if on_device:
        from .restriction_operators import get_physical_points

        desc = red_op.device_descriptor(lambda: get_physical_points(mesh_to, all_or(cells_K), ref_points))
        gamma_args = (np.asarray(mesh_to.geometry.x), np.asarray(mesh_to.geometry.dofmap), cells_K, ref_points, row_dofs,
                      n_rows_total, desc)
        if args_only:
            return gamma_args
        return DeviceRows.from_gamma(*gamma_args)
    if args_only:
        return None
    cells_K = all_or(cells_K)
    if row_dofs is None:
        row_dofs = k_list.reshape(-1)
    q = red_op.compute_quadrature(cells_K, ref_points)
    return DeviceRows.from_points(q.points, q.weights, q.scales, row_dofs, n_rows_total)


def device_rows_block(K, red_op, cells_block: np.ndarray, known_contiguous: bool = False):
    """A rows handle for a contiguous ascending block of Γ cells with LOCAL row numbering: row r of the block's Π is
    the dof ``cells_block[0]*nip + r`` of K.  Only for spaces with cell-local dofs (Quadrature / DG: the rows of a block
    of cells are a block of rows) and the stock operators; returns None otherwise.  Used to build Π block by block
    (``distributed.assemble_blocks_to_host``) — the block matrices stacked are the matrix of one call."""
    mesh_to = K.mesh
    k_list = np.asarray(K.dofmap.list)
    ref_points = np.asarray(K.element.interpolation_points, dtype=np.float64)
    cells_block = np.ascontiguousarray(cells_block, dtype=np.int32)
    if cells_block.size == 0 or k_list.shape[1] != ref_points.shape[0]:
        return None
    identity = K.__dict__.get("_fxii_identity_dofs") if hasattr(K, "__dict__") else None
    if identity is None:
        flat = k_list.reshape(-1)
        identity = bool(np.array_equal(flat, np.arange(flat.shape[0], dtype=flat.dtype)))
        if hasattr(K, "__dict__"):
            K.__dict__["_fxii_identity_dofs"] = identity
    contiguous = known_contiguous or (int(cells_block[-1]) - int(cells_block[0]) + 1 == cells_block.size
                                      and bool(np.all(np.diff(cells_block) == 1)))
    on_device = _uses_stock_quadrature(red_op) and red_op._mesh is mesh_to and mesh_to.topology.dim == 1 \
        and np.asarray(mesh_to.geometry.dofmap).shape[1] == 2
    if not (identity and contiguous and on_device):
        return None
    from .restriction_operators import get_physical_points

    desc = dict(red_op.device_descriptor(lambda: get_physical_points(mesh_to, cells_block, ref_points)))
    if "cell_radius" in desc:  # a contiguous block: its radii are a slice, not a gather
        desc["cell_radius"] = desc["cell_radius"][int(cells_block[0]):int(cells_block[-1]) + 1]
        desc["cell_radius_is_local"] = True
    nip = ref_points.shape[0]
    return DeviceRows.from_gamma(np.asarray(mesh_to.geometry.x), np.asarray(mesh_to.geometry.dofmap), cells_block, ref_points, None,
                                 cells_block.size * nip, desc)


# Π cache: the reference rebuilds Π for every integral of every block and twice per
# assemble (matrix_assembler.py:177,203,229,239,267-294).  Entries hold strong references to their key
# objects so that `id()` stays unique while an entry lives; least recently used entries are dropped.
_PI_CACHE: "dict[tuple, tuple]" = {}
_PI_CACHE_SIZE = 8


def clear_interpolation_cache():
    _PI_CACHE.clear()


def _geometry_token(mesh):
    """A cheap fingerprint of a mesh's coordinates (address, shape and a weighted sum over ~32 sampled nodes): a cached Π is
    not reused after the mesh was moved in place.  Not a hash: changes it cannot see (and changed radius
    functions) need ``clear_interpolation_cache``."""
    x = np.asarray(mesh.geometry.x)
    rows = x.reshape(x.shape[0], -1)[:: max(1, x.shape[0] // 32)]  # ~32 nodes, every coordinate of each
    w = 1.0 + 0.25 * np.arange(rows.shape[1])                      # columns weighted apart: a swap of axes is seen too
    return (x.__array_interface__["data"][0], x.shape, float((rows * w).sum()))


def _cache_key(V, K, red_op, tol, cells_K):
    ck = None if cells_K is None else np.ascontiguousarray(cells_K, dtype=np.int32).tobytes()
    return (id(V), id(K), id(red_op), float(tol), ck, _geometry_token(V.mesh), _geometry_token(K.mesh))


def create_interpolation_matrix(V, K, red_op, tol: float = 1.0e-8, use_petsc: bool = False, cells_K=None,
                                complex_dtype: bool = False, materialize_coo: bool = False, cache: bool = False,
                                host_out=None):
    """Create an interpolation matrix from `V` to `K` with a specific reduction operator applied to
    the interpolation points of `K`.

    Args:
        V: The function space to interpolate from (Lagrange P1/P2 on affine tetrahedra, scalar or blocked).
        K: The function space to interpolate to (on the 1D mesh; same block size as V).
        red_op: Reduction operator for each interpolation point on `K`.
        tol: Tolerance (box padding) for determining point ownership.
        use_petsc: Return a PETSc AIJ matrix (requires petsc4py) instead of a ``MatrixCSR``.
        cells_K: Optional subset of the cells of K to process.
        complex_dtype: Return a matrix whose host values are complex128 (the entries of Π are real: the imaginary part
            is zero; the device arithmetic is float64 either way).
        materialize_coo: (extension) force the unfused COO path, e.g. to inspect the COO.
        cache: (extension) reuse the matrix of an earlier call with the same `V`, `K`, `red_op`,
            `tol` and `cells_K` objects (see ``clear_interpolation_cache``).
        host_out: (extension) ``(indptr int64[rows+1], indices int32[cap], values float64[cap])`` host arrays,
            ideally pinned: the matrix is delivered into them (the returned ``MatrixCSR`` holds views) and the
            download overlaps the build's only host synchronisation.  ``ValueError`` if ``cap < nnz``.

    Returns:
        ``(A, imap_K, imap_V)``: the matrix and the (serial: unchanged) index maps of K and V.
    """
    # scalar type of the returned matrix (interpolation.py:175-203): complex when asked for, else the mesh's real type
    geom_dtype = np.asarray(V.mesh.geometry.x).dtype
    out_dtype = np.complex128 if complex_dtype else (np.float32 if geom_dtype == np.float32 else np.float64)
    if complex_dtype and geom_dtype == np.float32:
        out_dtype = np.complex64
    imap_K, imap_V = K.dofmap.index_map, V.dofmap.index_map
    key = _cache_key(V, K, red_op, tol, cells_K) if cache and not materialize_coo else None
    if key is not None and key in _PI_CACHE:
        A = _PI_CACHE.pop(key)[0]
        _PI_CACHE[key] = (A, V, K, red_op)  # most recently used last
        if A.dtype != np.dtype(out_dtype):  # same device matrix, another host scalar type
            A = MatrixCSR(A.device, stats=A.stats, rows=A.rows, dtype=out_dtype)
        return (A.to_petsc() if use_petsc else A), imap_K, imap_V
    bs = int(getattr(V.dofmap, "bs", 1))
    if bs != int(getattr(K.dofmap, "bs", 1)):
        raise NotImplementedError("V and K must have the same block size")
    dspace = device_space(V, tol)
    if host_out is not None and bs == 1 and not use_petsc and not materialize_coo:
        gamma_args = device_rows(K, red_op, cells_K, args_only=True)
        if gamma_args is not None:
            # host arrays in, host CSR out: ONE C call and one synchronisation
            hints = red_op.__dict__.setdefault("_fxii_nnz_hints", {}) if hasattr(red_op, "__dict__") else {}
            hkey = (id(K), id(V), None if cells_K is None or len(cells_K) == 0 else (len(cells_K), int(cells_K[0]), int(cells_K[-1])))
            indptr_h, indices_h, values_h = host_out
            rows, nnz, stats = DeviceRows.build_to_host(dspace, gamma_args, indptr_h, indices_h, values_h, hints.get(hkey, 0))
            hints[hkey] = nnz
            n_rows = rows.n_rows_total
            A = MatrixCSR(None, stats=stats, rows=rows, host=(indptr_h[: n_rows + 1], indices_h[:nnz], values_h[:nnz]),
                          shape=(n_rows, dspace.n_dofs), dtype=out_dtype)
            return A, imap_K, imap_V
    rows = device_rows(K, red_op, cells_K)
    if materialize_coo:
        rows.set_mode(1)
    # nnz of the previous build with the same (K, operator, cells_K): lets the build allocate up front
    hints = red_op.__dict__.setdefault("_fxii_nnz_hints", {}) if hasattr(red_op, "__dict__") else {}
    hkey = (id(K), id(V), None if cells_K is None or len(cells_K) == 0 else (len(cells_K), int(cells_K[0]), int(cells_K[-1])))
    if hkey in hints:
        rows.set_nnz_hint(hints[hkey])
    if host_out is not None and bs == 1 and not use_petsc:
        # straight into the caller's (pinned) host arrays: one C call, one synchronisation
        indptr_h, indices_h, values_h = host_out
        nnz, stats = rows.assemble_to_host(dspace, indptr_h, indices_h, values_h)
        hints[hkey] = nnz
        n_rows = rows.n_rows_total
        A = MatrixCSR(None, stats=stats, rows=rows, host=(indptr_h[: n_rows + 1], indices_h[:nnz], values_h[:nnz]),
                      shape=(n_rows, dspace.n_dofs), dtype=out_dtype)
        return A, imap_K, imap_V
    general = K.__dict__.get("_fxii_general_path", False) if hasattr(K, "__dict__") else False
    if general and not materialize_coo:
        rows.set_mode(1)  # this K was seen to share dofs between point rows: skip the fused attempt
    csr, stats = rows.assemble(dspace)
    if not general and not materialize_coo and stats["fused"] == 0 and hasattr(K, "__dict__"):
        K.__dict__["_fxii_general_path"] = True
    hints[hkey] = csr.nnz
    if bs > 1:
        csr = csr.expand_blocks(bs)  # device kernel: bs x bs blocks, explicit zeros off the diagonal
    A = MatrixCSR(csr, stats=stats, rows=rows, dtype=out_dtype)
    if key is not None:
        _PI_CACHE[key] = (A, V, K, red_op)
        while len(_PI_CACHE) > _PI_CACHE_SIZE:
            _PI_CACHE.pop(next(iter(_PI_CACHE)))
    if use_petsc:
        return A.to_petsc(), imap_K, imap_V
    return A, imap_K, imap_V
