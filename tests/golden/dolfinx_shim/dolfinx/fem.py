"""fem.Expression for the three expressions the reference evaluates, and a minimal function space."""
from types import SimpleNamespace

import numpy as np
import ufl

from mpi4py import MPI

from .common import IndexMap


class Expression:
    def __init__(self, e, X, comm=None, dtype=np.float64):
        self.e = e
        self.X = np.asarray(X, dtype=np.float64)

    def eval(self, mesh, cells):
        from oracle import pi_oracle as orc

        cells = np.asarray(cells)
        x, dm = mesh.geometry.x, mesh.geometry.dofmap
        if isinstance(self.e, ufl.SpatialCoordinate):  # utils.py:18-20
            pts = orc.physical_points(x, dm, cells, self.X.reshape(-1))
            return pts.reshape(len(cells), -1, 3)
        if isinstance(self.e, ufl.geometry.CellVertices):  # utils.py:29-30 -> [cells, points, vertices, gdim]
            v = x[dm[cells]]
            return v[:, None, :, :]
        if isinstance(self.e, ufl.TestFunction):  # interpolation_utils.py:63-74 -> [cells, points, dofs]
            V = self.e.V
            if V.mesh.geometry.dofmap.shape[1] == 8:  # hexahedron: Q1
                assert V.element.degree == 1
                phi = orc.q1_tabulate(self.X)
            else:
                phi = orc.tabulate(V.element.degree, self.X)
            bs = V.dofmap.bs
            if bs == 1:
                return np.broadcast_to(phi[None, :, :], (len(cells),) + phi.shape).copy()
            # blocked space: [cells, points, component, unrolled dof d*bs+c] with phi_d on the matching component
            # (the layout interpolation_utils.py:66-69 takes its diagonal of)
            nd = phi.shape[1]
            vals = np.zeros((len(cells), phi.shape[0], bs, nd * bs))
            for c in range(bs):
                vals[:, :, c, c::bs] = phi[None, :, :]
            return vals
        raise NotImplementedError(type(self.e))


class FunctionSpace:
    def __init__(self, mesh, degree, dof_list, n_dofs, interpolation_points=None, family="Lagrange", bs=1):
        """`n_dofs`: number of BLOCK dofs (the index map counts blocks, as in dolfinx)."""
        self.mesh = mesh
        imap = IndexMap(MPI.COMM_WORLD, n_dofs)
        self.dofmap = SimpleNamespace(list=np.asarray(dof_list, dtype=np.int32), bs=bs, index_map=imap, index_map_bs=bs,
                                      dof_layout=SimpleNamespace(num_dofs=np.asarray(dof_list).shape[1]))
        self.element = SimpleNamespace(
            degree=degree, interpolation_points=interpolation_points, interpolation_ident=True,
            needs_dof_transformations=False, dtype=np.float64, basix_element=SimpleNamespace(value_shape=()),
            family=family)


class CoordinateElement:  # annotation target only (compat.py:6)
    pass
