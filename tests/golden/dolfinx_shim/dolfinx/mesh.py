import numpy as np

from mpi4py import MPI

from .common import IndexMap


class _CMap:
    """Affine coordinate element: pull_back of CoordinateElement for simplices."""

    degree = 1

    def pull_back(self, x, cell_geometry):
        # X = K (x - v0), K = J^{-1} by the cofactor formula (oracle.cell_jacobian_inverse)
        from oracle import pi_oracle as orc

        cg = np.asarray(cell_geometry, dtype=np.float64)
        x = np.asarray(x, dtype=np.float64)
        if cg.shape[0] == 8:  # hexahedron: non-affine pull-back by Newton's method
            X, _ = orc.pull_back_hex(np.repeat(cg[None, :, :], x.shape[0], axis=0), x)
            return X
        v0, K = orc.cell_jacobian_inverse(cg, np.arange(4).reshape(1, 4))
        return orc.pull_back(v0, K, np.zeros(x.shape[0], dtype=np.int64), x)


    def push_forward(self, X, cell_geometry):
        # x = sum_k phi_k(X) v_k for the affine interval (MappedRestriction, restriction_operators.py:339-349)
        X = np.asarray(X, dtype=np.float64).reshape(-1)
        cg = np.asarray(cell_geometry, dtype=np.float64)
        assert cg.shape[0] == 2, "shim: push_forward for interval cells only"
        return cg[0][None, :] * (1.0 - X)[:, None] + cg[1][None, :] * X[:, None]


class _Geometry:
    def __init__(self, x, dofmap):
        self.x = np.asarray(x, dtype=np.float64)
        self.dofmap = np.asarray(dofmap, dtype=np.int32)
        self.dim = 3
        self.cmap = _CMap()


class _Topology:
    def __init__(self, dim, n_cells):
        self.dim = dim
        self._n = n_cells

    def index_map(self, dim):
        return IndexMap(MPI.COMM_WORLD, self._n)


class Mesh:
    def __init__(self, x, cells, tdim):
        self.geometry = _Geometry(x, cells)
        self.topology = _Topology(tdim, len(cells))
        self.comm = MPI.COMM_WORLD
        self.name = "mesh"

    def __str__(self):
        return f"ShimMesh(tdim={self.topology.dim})"
