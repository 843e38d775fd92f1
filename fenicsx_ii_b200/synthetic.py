"""Deterministic synthetic workloads of BASELINE.json's shapes.

Ω: Kuhn 6-tet split of an nx x ny x nz box (the split dolfinx's ``create_box`` uses: all six
tetrahedra of a hexahedron share the diagonal v0-v7).  Γ: straight line (config 1), polyline
network (config 2), random vessel tree with per-segment radius (configs 3-5).
"""

from __future__ import annotations

import numpy as np

from .mesh import FunctionSpace, Mesh, functionspace

SEED = 20261019

# local corner offsets v0..v7 of a hexahedron (v1=+x, v2=+y, v3=+x+y, v4=+z, ...)
_CORNER = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0], [0, 0, 1], [1, 0, 1], [0, 1, 1], [1, 1, 1]])
_KUHN = np.array([[0, 1, 3, 7], [0, 1, 7, 5], [0, 5, 7, 4], [0, 3, 2, 7], [0, 6, 4, 7], [0, 2, 6, 7]])
_TET_EDGES = ((2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1))


def kuhn_box(nx: int, ny: int, nz: int, lo=(0.0, 0.0, 0.0), hi=(1.0, 1.0, 1.0), name="volume") -> Mesh:
    """6*nx*ny*nz tetrahedra; vertex (i,j,k) has index (k*(ny+1)+j)*(nx+1)+i (x fastest)."""
    lo = np.asarray(lo, dtype=np.float64)
    hi = np.asarray(hi, dtype=np.float64)
    xs = lo[0] + (hi[0] - lo[0]) * (np.arange(nx + 1) / nx)
    ys = lo[1] + (hi[1] - lo[1]) * (np.arange(ny + 1) / ny)
    zs = lo[2] + (hi[2] - lo[2]) * (np.arange(nz + 1) / nz)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    x = np.stack([X.reshape(-1), Y.reshape(-1), Z.reshape(-1)], axis=1)
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    base = ((k * (ny + 1) + j) * (nx + 1) + i).reshape(-1).astype(np.int64)
    corner_off = (_CORNER[:, 2] * (ny + 1) + _CORNER[:, 1]) * (nx + 1) + _CORNER[:, 0]
    hexv = base[:, None] + corner_off[None, :]  # [nhex, 8]
    cells = hexv[:, _KUHN].reshape(-1, 4).astype(np.int32)
    mesh = Mesh(x, cells, tdim=3, name=name)
    mesh.grid_shape = (nx, ny, nz)
    return mesh


def hex_box(nx: int, ny: int, nz: int, lo=(0.0, 0.0, 0.0), hi=(1.0, 1.0, 1.0), taper: float = 0.0, name="hex_volume") -> Mesh:
    """nx*ny*nz hexahedra (what dolfinx ``create_box(..., CellType.hexahedron)`` produces, before its reordering);
    cell vertices in tensor-product order v = i + 2j + 4k.  ``taper`` != 0 maps (x, y, z) -> (x (1 + taper z),
    y (1 + taper z), z): every cell becomes a frustum with PLANAR faces whose trilinear map is not affine — the
    case the Newton pull-back exists for."""
    lo = np.asarray(lo, dtype=np.float64)
    hi = np.asarray(hi, dtype=np.float64)
    xs = lo[0] + (hi[0] - lo[0]) * (np.arange(nx + 1) / nx)
    ys = lo[1] + (hi[1] - lo[1]) * (np.arange(ny + 1) / ny)
    zs = lo[2] + (hi[2] - lo[2]) * (np.arange(nz + 1) / nz)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    x = np.stack([X.reshape(-1), Y.reshape(-1), Z.reshape(-1)], axis=1)
    if taper != 0.0:
        s = 1.0 + taper * x[:, 2]
        x = np.stack([x[:, 0] * s, x[:, 1] * s, x[:, 2]], axis=1)
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    base = ((k * (ny + 1) + j) * (nx + 1) + i).reshape(-1).astype(np.int64)
    corner_off = (_CORNER[:, 2] * (ny + 1) + _CORNER[:, 1]) * (nx + 1) + _CORNER[:, 0]
    cells = (base[:, None] + corner_off[None, :]).astype(np.int32)  # _CORNER is already v = i + 2j + 4k
    mesh = Mesh(x, cells, tdim=3, name=name)
    mesh.grid_shape = (nx, ny, nz)
    return mesh


def kuhn_p2_space(mesh: Mesh) -> FunctionSpace:
    """P2 space on a ``kuhn_box`` mesh with a closed-form edge numbering (no global sort).

    Every mesh edge joins a vertex to the vertex offset by a nonzero vector in {0,1}^3, so an edge is
    (direction code 1..7, lower vertex); edge dofs are numbered by the rank of that pair."""
    nx, ny, nz = mesh.grid_shape
    cells = mesh.geometry.dofmap.astype(np.int64)
    n_nodes = mesh.geometry.x.shape[0]
    sx, sy = 1, nx + 1
    sz = (nx + 1) * (ny + 1)
    vid = np.arange(n_nodes, dtype=np.int64)
    vi, vj, vk = vid % (nx + 1), (vid // (nx + 1)) % (ny + 1), vid // sz
    rank = np.full(8 * n_nodes, -1, dtype=np.int64)
    count = 0
    for code in range(1, 8):
        dx, dy, dz = code & 1, (code >> 1) & 1, (code >> 2) & 1
        valid = (vi + dx <= nx) & (vj + dy <= ny) & (vk + dz <= nz)
        n_valid = int(valid.sum())
        rank[code * n_nodes + vid[valid]] = count + np.arange(n_valid)
        count += n_valid
    dof_list = np.empty((cells.shape[0], 10), dtype=np.int32)
    dof_list[:, :4] = cells
    for e, (a, b) in enumerate(_TET_EDGES):
        lo = np.minimum(cells[:, a], cells[:, b])
        diff = np.maximum(cells[:, a], cells[:, b]) - lo
        dz = diff // sz
        dy = (diff - dz * sz) // sy
        dx = diff - dz * sz - dy * sy
        # offsets like (-1,+1,0) cannot occur in a Kuhn mesh; guard anyway
        assert ((dx >= 0) & (dx <= 1) & (dy >= 0) & (dy <= 1)).all()
        code = dx + 2 * dy + 4 * dz
        r = rank[code * n_nodes + lo]
        assert (r >= 0).all()
        dof_list[:, 4 + e] = (n_nodes + r).astype(np.int32)
    return functionspace(mesh, ("Lagrange", 2), dof_list=dof_list, n_dofs=n_nodes + count)


def line_mesh(nodes: np.ndarray, name="line") -> Mesh:
    n = nodes.shape[0]
    cells = np.stack([np.arange(n - 1), np.arange(1, n)], axis=1).astype(np.int32)
    return Mesh(nodes, cells, tdim=1, name=name)


def straight_line(n_nodes: int = 64, x=0.5, y=0.5, z0=1.0, z1=0.0) -> Mesh:
    """Config 1 (demos/coupled_poisson.py:24-31): nodes on x=y=0.5, z from 1 down to 0."""
    nodes = np.zeros((n_nodes, 3))
    nodes[:, 0] = x
    nodes[:, 1] = y
    nodes[:, 2] = np.linspace(min(z0, z1), max(z0, z1), n_nodes)[::-1] if z0 > z1 else np.linspace(z0, z1, n_nodes)
    return line_mesh(nodes)


def _fold(p: np.ndarray, lo: float, hi: float) -> np.ndarray:
    """Reflect coordinates into [lo, hi] (triangle wave): keeps a random walk inside the box."""
    span = hi - lo
    q = np.mod(p - lo, 2 * span)
    return lo + np.where(q > span, 2 * span - q, q)


def polyline_network(n_segments: int, h: float, lo: float = 0.1, hi: float = 0.9, n_lines: int = 16,
                     seed: int = SEED) -> Mesh:
    """Config 2: ``n_lines`` independent polylines, ``n_segments`` segments in total, segment length
    about ``h``, inside [lo,hi]^3, generic position (irrational offsets)."""
    rng = np.random.default_rng(seed)
    per = [n_segments // n_lines + (1 if i < n_segments % n_lines else 0) for i in range(n_lines)]
    nodes, cells = [], []
    off = 0
    for n_seg in per:
        if n_seg == 0:
            continue
        start = rng.uniform(lo, hi, size=3)
        path = np.vstack([start[None, :], _walk_vec(rng, start, n_seg, h * 0.97)])
        path = _fold(path, lo, hi)
        nodes.append(path)
        idx = off + np.arange(n_seg + 1)
        cells.append(np.stack([idx[:-1], idx[1:]], axis=1))
        off += n_seg + 1
    x = np.vstack(nodes) + np.sqrt(2.0) * 1e-7  # keep nodes off the mesh planes
    return Mesh(x, np.vstack(cells).astype(np.int32), tdim=1, name="network")


def _walk_vec(rng, start: np.ndarray, n_seg: int, step: float, wobble: float = 0.35) -> np.ndarray:
    """Vectorised smooth walk: directions are a normalised AR(1) process (no Python loop per step)."""
    kicks = rng.normal(size=(n_seg, 3))
    # exponential smoothing of the kicks gives slowly turning directions
    from scipy.signal import lfilter

    d = lfilter([wobble], [1.0, -(1.0 - wobble)], kicks, axis=0) + rng.normal(size=3) * (1.0 - wobble) ** np.arange(1, n_seg + 1)[:, None]
    nrm = np.linalg.norm(d, axis=1)
    nrm[nrm == 0] = 1.0
    d = d / nrm[:, None]
    return start[None, :] + np.cumsum(d * step, axis=0)


def vessel_tree(n_segments: int, h: float, r_min: float | None = None, r_max: float | None = None,
                branch_len: int = 200, seed: int = SEED, box=(0.0, 1.0)):
    """Configs 3-5: a random tree of ``n_segments`` segments of length about ``h``.  Branches of
    ``branch_len`` segments start at random nodes of earlier branches (bifurcations); the radius
    follows Murray's law (child = parent * 2**(-1/3)) clipped below at ``r_min``; the root has
    ``r_max``.  Everything stays at least ``r_max + 2h`` away from the box boundary.

    Returns (mesh, seg_radius[n_segments])."""
    import bisect

    rng = np.random.default_rng(seed)
    r_min = 0.5 * h if r_min is None else r_min
    r_max = 4.0 * h if r_max is None else r_max
    margin = r_max + 2.0 * h
    lo, hi = box[0] + margin, box[1] - margin
    assert hi > lo, "box too small for the requested radii"
    blocks, cells, radii = [], [], []
    block_start, block_radius = [], []  # first node index / radius of every branch
    n_nodes = 0
    remaining = n_segments
    while remaining > 0:
        n_seg = min(branch_len, remaining)
        if n_nodes == 0:
            start = rng.uniform(lo, hi, size=3)
            radius = r_max
            path = np.vstack([start[None, :], _fold(_walk_vec(rng, start, n_seg, h * 0.97), lo, hi)])
            ids = np.arange(n_seg + 1)
            new_cells = np.stack([ids[:-1], ids[1:]], axis=1)
        else:
            parent = int(rng.integers(0, n_nodes))
            b = bisect.bisect_right(block_start, parent) - 1
            start = blocks[b][parent - block_start[b]]
            radius = max(r_min, block_radius[b] * 2.0 ** (-1.0 / 3.0))
            path = _fold(_walk_vec(rng, start, n_seg, h * 0.97), lo, hi)
            ids = n_nodes + np.arange(n_seg)
            new_cells = np.vstack([np.array([[parent, n_nodes]]), np.stack([ids[:-1], ids[1:]], axis=1)])
        block_start.append(n_nodes)
        block_radius.append(radius)
        blocks.append(path)
        cells.append(new_cells)
        radii.append(np.full(n_seg, radius))
        n_nodes += path.shape[0]
        remaining -= n_seg
    x = np.vstack(blocks)
    mesh = Mesh(x, np.vstack(cells).astype(np.int32), tdim=1, name="vessel_tree")
    return mesh, np.concatenate(radii)
