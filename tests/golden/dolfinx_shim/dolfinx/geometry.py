"""determine_point_ownership, serial: oracle GridLocator (lowest colliding cell index wins)."""
from types import SimpleNamespace

import numpy as np


def determine_point_ownership(mesh, points, padding=0.0):
    from oracle import pi_oracle as orc

    loc = getattr(mesh, "_shim_locator", None)
    if loc is None or loc.padding != padding:
        loc = mesh._shim_locator = orc.GridLocator(mesh.geometry.x, mesh.geometry.dofmap, padding=padding)
    cell, ncoll = loc.locate(np.asarray(points, dtype=np.float64))
    n = points.shape[0]
    owner = np.where(cell >= 0, 0, -1).astype(np.int32)
    mesh._shim_last = SimpleNamespace(cell=cell, n_colliding=ncoll)
    return SimpleNamespace(dest_cells=cell.astype(np.int32), dest_points=np.asarray(points, dtype=np.float64).copy(),
                           src_owner=owner, dest_owner=np.zeros(n, dtype=np.int32))
