"""Array-backed stand-ins for the dolfinx objects the hot path reads.

The reference takes ``dolfinx.fem.FunctionSpace`` / ``dolfinx.mesh.Mesh`` objects
(``interpolation.py:15-23``).  dolfinx is absent from this image, so the same *attribute surface*
is provided over plain numpy arrays: ``mesh.geometry.x``, ``mesh.geometry.dofmap``,
``mesh.geometry.dim``, ``mesh.topology.dim``, ``mesh.topology.index_map(d).size_local``,
``V.mesh``, ``V.dofmap.list``, ``V.dofmap.bs``, ``V.dofmap.index_map``, ``V.dofmap.index_map_bs``,
``V.element.interpolation_points`` ...  ``create_interpolation_matrix`` only uses these attributes,
so real dolfinx objects work unchanged where dolfinx is installed.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .quadrature import gauss_jacobi_interval

__all__ = ["FunctionSpace", "IndexMap", "Mesh", "functionspace"]


class IndexMap:
    """Serial index map: ``size_local`` owned indices, no ghosts (dolfinx.common.IndexMap surface)."""

    def __init__(self, size_local: int, comm=None):
        self.size_local = int(size_local)
        self.num_ghosts = 0
        self.ghosts = np.zeros(0, dtype=np.int64)
        self.owners = np.zeros(0, dtype=np.int32)
        self.comm = comm
        self.local_range = (0, self.size_local)
        self.size_global = self.size_local

    def local_to_global(self, idx):
        return np.asarray(idx, dtype=np.int64)

    def global_to_local(self, idx):
        idx = np.asarray(idx, dtype=np.int64)
        return np.where((idx >= 0) & (idx < self.size_local), idx, -1).astype(np.int32)


@dataclass
class _CMap:
    degree: int = 1


@dataclass
class Geometry:
    x: np.ndarray
    dofmap: np.ndarray
    dim: int = 3
    cmap: _CMap = field(default_factory=_CMap)


class Topology:
    def __init__(self, dim: int, n_cells: int, n_vertices: int):
        self.dim = dim
        self._maps = {dim: IndexMap(n_cells), 0: IndexMap(n_vertices)}
        self.cell_name = {1: "interval", 3: "tetrahedron"}.get(dim, "unknown")  # "hexahedron" is set by Mesh

    def index_map(self, dim: int) -> IndexMap:
        return self._maps[dim]


class Mesh:
    """Mesh given by node coordinates ``x [n,3]`` and cells: simplices ``[n_cells, tdim+1]`` (intervals, tetrahedra) or
    hexahedra ``[n_cells, 8]`` with Q1 geometry, vertices in dolfinx's tensor-product order v = i + 2j + 4k."""

    def __init__(self, x, cells, tdim: int, name: str = "mesh"):
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.shape[1] != 3:
            xx = np.zeros((x.shape[0], 3))
            xx[:, : x.shape[1]] = x
            x = xx
        cells = np.ascontiguousarray(cells, dtype=np.int32)
        assert cells.shape[1] == tdim + 1 or (tdim == 3 and cells.shape[1] == 8)
        self.geometry = Geometry(x=x, dofmap=cells, dim=3)
        self.topology = Topology(tdim, cells.shape[0], x.shape[0])
        if tdim == 3 and cells.shape[1] == 8:
            self.topology.cell_name = "hexahedron"
        self.name = name
        self.comm = None

    def __str__(self) -> str:
        return f"Mesh({self.name}, tdim={self.topology.dim}, cells={self.geometry.dofmap.shape[0]})"


@dataclass
class Element:
    family: str
    degree: int
    interpolation_points: np.ndarray  # reference points, shape (n, tdim)
    interpolation_ident: bool = True
    needs_dof_transformations: bool = False
    dtype: type = np.float64


class DofMap:
    def __init__(self, dof_list: np.ndarray, n_dofs: int, bs: int = 1):
        self.list = np.ascontiguousarray(dof_list, dtype=np.int32)
        self.bs = bs
        self.index_map_bs = bs
        self.index_map = IndexMap(n_dofs)


class FunctionSpace:
    def __init__(self, mesh: Mesh, element: Element, dofmap: DofMap):
        self.mesh = mesh
        self.element = element
        self.dofmap = dofmap

    @property
    def num_dofs(self) -> int:
        """Scalar dof count times the block size."""
        return self.dofmap.index_map.size_local * self.dofmap.index_map_bs


_TET_EDGES = ((2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1))  # basix reference tetrahedron


def _tet_p2_dofmap(cells: np.ndarray, n_nodes: int) -> tuple[np.ndarray, int]:
    """Vertex dofs followed by one dof per mesh edge, cell-local order = basix P2 order."""
    cells64 = cells.astype(np.int64)
    keys = []
    for a, b in _TET_EDGES:
        lo = np.minimum(cells64[:, a], cells64[:, b])
        hi = np.maximum(cells64[:, a], cells64[:, b])
        keys.append(lo * n_nodes + hi)
    keys = np.stack(keys, axis=1)
    uniq, inv = np.unique(keys.reshape(-1), return_inverse=True)
    edge_dofs = (n_nodes + inv).reshape(keys.shape)
    dof_list = np.concatenate([cells64, edge_dofs], axis=1).astype(np.int32)
    return dof_list, n_nodes + uniq.shape[0]


_TET_FACES = ((1, 2, 3), (0, 2, 3), (0, 1, 3), (0, 1, 2))  # basix reference tetrahedron


def _tet_p3_dofmap(cells: np.ndarray, n_nodes: int) -> tuple[np.ndarray, int]:
    """Vertex dofs, two dofs per mesh edge, one per mesh face; cell-local order = basix P3 order (vertices, edges,
    faces).  The two dofs of an edge are numbered along its REFERENCE direction, from the vertex with the lower global
    index to the higher one — the permutation dolfinx applies to the dofmap of a cell that sees the edge reversed — so
    cells sharing an edge agree on which dof sits near which vertex (the space is conforming)."""
    cells64 = cells.astype(np.int64)
    ekeys, flipped = [], []
    for a, b in _TET_EDGES:
        va, vb = cells64[:, a], cells64[:, b]
        ekeys.append(np.minimum(va, vb) * n_nodes + np.maximum(va, vb))
        flipped.append(va > vb)  # the cell's local direction a -> b runs against the reference direction
    ekeys = np.stack(ekeys, axis=1)
    flipped = np.stack(flipped, axis=1)
    uniq_e, inv_e = np.unique(ekeys.reshape(-1), return_inverse=True)
    edge_id = inv_e.reshape(ekeys.shape)
    n_edges = uniq_e.shape[0]
    first = n_nodes + 2 * edge_id + flipped   # local dof 0 (nearer local vertex a)
    second = n_nodes + 2 * edge_id + (~flipped)
    ftri = np.stack([np.sort(cells64[:, list(f)], axis=1) for f in _TET_FACES], axis=1)  # [cells, 4, 3] sorted vertices
    if n_nodes < 2_000_000:  # the packed key fits 63 bits
        fkeys = (ftri[:, :, 0] * n_nodes + ftri[:, :, 1]) * n_nodes + ftri[:, :, 2]
        uniq_f, inv_f = np.unique(fkeys.reshape(-1), return_inverse=True)
    else:
        uniq_f, inv_f = np.unique(ftri.reshape(-1, 3), axis=0, return_inverse=True)
    face_dofs = n_nodes + 2 * n_edges + np.asarray(inv_f).reshape(cells.shape[0], 4)
    edge_dofs = np.stack([first, second], axis=2).reshape(cells.shape[0], 12)
    dof_list = np.concatenate([cells64, edge_dofs, face_dofs], axis=1).astype(np.int32)
    return dof_list, n_nodes + 2 * n_edges + uniq_f.shape[0]


def functionspace(mesh: Mesh, element, dof_list: np.ndarray | None = None, n_dofs: int | None = None) -> FunctionSpace:
    """``dolfinx.fem.functionspace`` for the element families the hot path supports.

    3D (tetrahedra): ("Lagrange", 1|2|3).  1D (intervals): ("Quadrature", q) — the ``(q+2)//2``
    Gauss points per cell, cell-local dofs; ("DG", k) k=0,1,2; ("Lagrange", 1|2).
    A precomputed ``dof_list`` (e.g. the closed-form P2 numbering of a structured mesh) may be given."""
    bs = 1
    if len(element) == 3:  # ("Lagrange", 1, (3,)): blocked (vector) space, dolfinx style
        family, degree, shape = element
        bs = int(np.prod(shape))
        V = functionspace(mesh, (family, degree), dof_list, n_dofs)
        V.dofmap.bs = V.dofmap.index_map_bs = bs
        return V
    family, degree = element
    family = {"P": "Lagrange", "CG": "Lagrange", "Discontinuous Lagrange": "DG"}.get(family, family)
    cells = mesh.geometry.dofmap
    n_cells, n_nodes = cells.shape[0], mesh.geometry.x.shape[0]
    tdim = mesh.topology.dim
    if tdim == 3 and cells.shape[1] == 8:
        if family != "Lagrange" or degree != 1:
            raise NotImplementedError("hexahedra: Lagrange degree 1 (Q1)")
        return FunctionSpace(mesh, Element("Lagrange", 1, np.zeros((0, 3))), DofMap(cells.copy(), n_nodes))
    if tdim == 3:
        if family != "Lagrange" or degree not in (1, 2, 3):
            raise NotImplementedError("3D spaces: Lagrange degree 1, 2 or 3 on affine tetrahedra")
        if dof_list is None:
            if degree == 1:
                dof_list, n_dofs = cells.copy(), n_nodes
            elif degree == 2:
                dof_list, n_dofs = _tet_p2_dofmap(cells, n_nodes)
            else:
                dof_list, n_dofs = _tet_p3_dofmap(cells, n_nodes)
        pts = np.zeros((0, 3))
        return FunctionSpace(mesh, Element("Lagrange", degree, pts), DofMap(dof_list, int(n_dofs)))
    if tdim != 1:
        raise NotImplementedError("only interval and tetrahedron meshes")
    if family == "Quadrature":
        pts, _ = gauss_jacobi_interval(degree)
        nip = pts.shape[0]
        dof_list = np.arange(n_cells * nip, dtype=np.int32).reshape(n_cells, nip)
        return FunctionSpace(mesh, Element("Quadrature", degree, pts), DofMap(dof_list, n_cells * nip))
    nodal = {0: [0.5], 1: [0.0, 1.0], 2: [0.0, 1.0, 0.5]}
    if degree not in nodal or (family == "Lagrange" and degree == 0):
        raise NotImplementedError(f"{family} degree {degree} on intervals")
    pts = np.array(nodal[degree]).reshape(-1, 1)
    if family == "DG":
        nip = pts.shape[0]
        dof_list = np.arange(n_cells * nip, dtype=np.int32).reshape(n_cells, nip)
        return FunctionSpace(mesh, Element("DG", degree, pts), DofMap(dof_list, n_cells * nip))
    if family == "Lagrange":
        if degree == 1:
            return FunctionSpace(mesh, Element("Lagrange", 1, pts), DofMap(cells.copy(), n_nodes))
        mid = n_nodes + np.arange(n_cells, dtype=np.int32)
        dof_list = np.concatenate([cells, mid[:, None]], axis=1)
        return FunctionSpace(mesh, Element("Lagrange", 2, pts), DofMap(dof_list, n_nodes + n_cells))
    raise NotImplementedError(f"element family {family}")
