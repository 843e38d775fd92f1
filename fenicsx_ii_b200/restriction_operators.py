"""Reduction operators from 3D to 1D — same classes, constructor arguments and validation as the
reference's ``restriction_operators.py`` (``ReductionOperator`` :17-47, ``PointwiseTrace`` :50-75,
``Circle`` :78-205, ``Disk`` :208-315, ``MappedRestriction`` :318-362).

``compute_quadrature`` keeps working on the host (it is the plugin contract, and generic operators
go through it); the three built-ins additionally expose ``device_descriptor()`` so that
``create_interpolation_matrix`` generates their averaging points on the GPU from Γ's tangents and
radius instead of uploading ``nq`` points per row.
"""

from __future__ import annotations

import typing

import numpy as np
import numpy.typing as npt

from ._lib import OP_CIRCLE, OP_DISK, OP_TRACE
from .quadrature import Quadrature, compute_disk_quadrature, gauss_jacobi_interval, rotation_matrix

__all__ = ["Circle", "Disk", "MappedRestriction", "PointwiseTrace", "ReductionOperator"]

RadiusT = typing.Union[float, typing.Callable[[npt.NDArray[np.floating]], npt.NDArray[np.floating]], np.ndarray]


def get_cmap(mesh):
    """compat.py:6-11: cmap is a method on some dolfinx versions and a property on others."""
    cmap = mesh.geometry.cmap
    return cmap() if callable(cmap) else cmap


def get_physical_points(mesh, cells, reference_points) -> np.ndarray:
    """Push-forward of reference points on affine interval cells (utils.py:11-20)."""
    x = np.asarray(mesh.geometry.x)
    dm = np.asarray(mesh.geometry.dofmap)
    X = np.asarray(reference_points, dtype=np.float64).reshape(-1)
    a = x[dm[cells, 0]][:, None, :]
    b = x[dm[cells, 1]][:, None, :]
    pts = a * (1.0 - X)[None, :, None] + b * X[None, :, None]
    return pts.reshape(-1, x.shape[1])[:, : mesh.geometry.dim]


def get_cell_normals(mesh, cells) -> np.ndarray:
    """Unit tangent of every Γ cell, vertex 0 minus vertex 1 (utils.py:23-33)."""
    assert mesh.topology.dim == 1, "Can only compute cell normals for 1D meshes."
    x = np.asarray(mesh.geometry.x)
    dm = np.asarray(mesh.geometry.dofmap)
    d = x[dm[cells, 0]] - x[dm[cells, 1]]
    return d / np.linalg.norm(d, axis=1)[:, None]


class ReductionOperator(typing.Protocol):
    """Protocol for reduction operators from 3D to 1D."""

    _mesh: typing.Any

    def compute_quadrature(
        self, cells: npt.NDArray[np.int32], reference_points: npt.NDArray[np.floating]
    ) -> Quadrature: ...

    @property
    def num_points(self) -> int: ...

    def __str__(self) -> str: ...


class PointwiseTrace:
    """Trace of the 3D function on the line segment."""

    def __init__(self, mesh):
        self._mesh = mesh

    def compute_quadrature(self, cells, reference_points) -> Quadrature:
        points = get_physical_points(self._mesh, cells, reference_points)
        return Quadrature(
            name="CircleAvg",  # sic: the reference names the trace rule "CircleAvg" (:66-68)
            points=points,
            weights=np.ones((points.shape[0], 1), dtype=points.dtype),
            scales=np.ones(points.shape[0], dtype=points.dtype),
        )

    @property
    def num_points(self) -> int:
        return 1

    def __str__(self) -> str:
        return f"PointwiseTrace({self._mesh})"

    def device_descriptor(self, x0_provider=None) -> dict:
        return dict(kind=OP_TRACE, nq=1, table=None, w=None, radius=None)


class _RadialAverage:
    """Shared body of Circle and Disk: a table of points on the unit circle/disk, rotated into the
    plane normal to Γ's tangent and scaled by the radius."""

    _name = ""
    _kind = -1

    def _init_common(self, mesh, radius: RadiusT):
        if mesh.topology.dim != 1:
            raise ValueError("Circle quadrature can only be used on 1D meshes")
        if get_cmap(mesh).degree > 1:
            raise NotImplementedError("Not implemented for curved meshes")
        self._mesh = mesh
        self._cell_radius = None
        if isinstance(radius, np.ndarray):
            # extension: one radius per Γ cell (vessel trees); the reference takes float | callable
            assert radius.shape == (mesh.geometry.dofmap.shape[0],)
            assert (radius >= 0).all(), "Radius cannot be zero or negative"
            self._cell_radius = np.ascontiguousarray(radius, dtype=np.float64)
            self._radius = None
        elif not callable(radius):
            assert radius >= 0, "Radius cannot be zero or negative"
            self._scalar_radius = float(radius)
            self._radius = lambda x: np.full(x.shape[1], radius)
        else:
            self._radius = radius

    @property
    def num_points(self) -> int:
        assert self._xp.shape[0] == len(self._w)
        return len(self._w)

    @property
    def interpolation_matrix(self):
        return np.eye(self.num_points)

    def _row_radius(self, cells, nip: int, x0: np.ndarray) -> np.ndarray:
        if self._cell_radius is not None:
            return np.repeat(self._cell_radius[cells], nip)
        return np.asarray(self._radius(x0.T), dtype=np.float64)

    def compute_quadrature(self, cells, reference_points) -> Quadrature:
        if get_cmap(self._mesh).degree > 1:
            raise NotImplementedError("Non-affine lines not supported yet")
        x0 = get_physical_points(self._mesh, cells, reference_points)
        nip = np.asarray(reference_points).shape[0]
        normals = np.repeat(get_cell_normals(self._mesh, cells), nip, axis=0)
        assert normals.shape == x0.shape
        if self._cell_radius is not None:
            R = np.repeat(self._cell_radius[cells], nip)
            points, weights, scales = self._quadrature_R(x0, normals, R)
        else:
            points, weights, scales = self.quadrature(x0, normals)
        return Quadrature(name=self._name, points=points, weights=weights, scales=scales)

    def quadrature(self, x0, normals):
        """Points [n,nq,3], weights [n,nq] and scales [n] for centres ``x0`` and normals."""
        assert x0.shape[0] == normals.shape[0], "Number of centers and normals must match"
        assert x0.shape[1] == 3, "Center points must be 3D"
        assert normals.shape[1] == 3, "Normal vectors must be 3D"
        if x0.shape[0] == 0:
            return np.zeros((0, x0.shape[1])), np.zeros((0, self._w.shape[0])), np.zeros((0,))
        if self._radius is None:
            raise ValueError("this operator holds one radius per Γ cell: centres alone do not determine the radius; "
                             "use compute_quadrature(cells, reference_points)")
        return self._quadrature_R(x0, normals, np.asarray(self._radius(x0.T), dtype=np.float64))

    def _quadrature_R(self, x0, normals, R):
        Rot = rotation_matrix(normals)
        offsets = np.einsum("ijk,lk->ilj", Rot, self._xp)
        pts = x0[:, None, :] + offsets * R[:, None, None]
        mult, scale = self._weight_multiplier(R)
        return pts, self._w[None, :] * mult[:, None], scale

    def device_descriptor(self, x0_provider=None) -> dict:
        """Tables and radii for on-device point generation.  ``x0_provider()`` returns the row
        centres [Nr,3]; it is only called for a callable radius (evaluated on the host at the Nr
        centres)."""
        d = dict(kind=self._kind, nq=self.num_points, table=np.ascontiguousarray(self._xp), w=np.ascontiguousarray(self._w))
        if self._cell_radius is not None:
            d["cell_radius"] = self._cell_radius
        elif hasattr(self, "_scalar_radius"):
            d["radius"] = np.array([self._scalar_radius])
        else:
            x0 = x0_provider()
            d["radius"] = np.ascontiguousarray(self._radius(x0.T), dtype=np.float64)
            d["radius_per_row"] = True
        return d


class Circle(_RadialAverage):
    """The average on the perimeter of a circle with a given `radius`.

    Args:
        mesh: The 1D mesh on which the circles are centered.
        radius: float, or a function of points (shape (3, n)) returning radii (shape (n,)).
            Extension: an array with one radius per Γ cell.
        degree: degree of the quadrature rule on the circle (at least 3).
    """

    _name = "CircleAvg"
    _kind = OP_CIRCLE

    def __init__(self, mesh, radius: RadiusT, degree: int = 15):
        if degree < 3:
            raise ValueError("Degree must be bigger than 3 for a sensible approximation")
        self._init_common(mesh, radius)
        s, w = gauss_jacobi_interval(degree)
        self._w = np.asarray(w)
        ang = 2 * np.pi * s[:, 0]
        self._xp = np.stack([np.cos(ang), np.sin(ang), np.zeros_like(ang)], axis=1)

    def __str__(self) -> str:
        return f"Circle({self._mesh}, radius={self._radius}, num_points={self._w.shape[0]})"

    @staticmethod
    def _weight_multiplier(R):
        circumference = 2 * R * np.pi
        return circumference, circumference


class Disk(_RadialAverage):
    """The average on a disk with a given `radius` (Bojanov-Petrova rule).

    Args:
        mesh: The 1D mesh on which the disks are centered.
        radius: float, or a function of points (shape (3, n)) returning radii (shape (n,)).
        degree: number of chords ``n`` of the rule (``n*n`` points, accuracy ``2n-1``).
    """

    _name = "DiskAvg"
    _kind = OP_DISK

    def __init__(self, mesh, radius: RadiusT, degree: int = 7):
        self._init_common(mesh, radius)
        xp, self._w = compute_disk_quadrature(degree)
        self._xp = np.zeros((xp.shape[0], 3), dtype=xp.dtype)
        self._xp[:, :2] = xp

    def __str__(self) -> str:
        return f"Disk({self._mesh}, radius={self._radius}, num_points={self._w.shape[0]})"

    @staticmethod
    def _weight_multiplier(R):
        return R**2, np.pi * R**2


class MappedRestriction:
    """A restriction mapping a point in space to another through ``x_new = operator(x)``.
    Generic operator: its points are computed on the host and uploaded (no device descriptor)."""

    def __init__(self, mesh, operator):
        self._mesh = mesh
        self._operator = operator

    def compute_quadrature(self, cells, reference_points) -> Quadrature:
        if self._mesh.topology.dim != 1:
            raise NotImplementedError("MappedRestriction: only interval meshes on this path")
        phys = get_physical_points(self._mesh, cells, reference_points)
        moved = np.asarray(self._operator(phys.T)).T
        return Quadrature(
            name="Translation", points=moved, weights=np.ones((phys.shape[0], 1)), scales=np.ones(phys.shape[0])
        )

    @property
    def num_points(self) -> int:
        return 1

    def __str__(self) -> str:
        return f"MappedRestriction({self._mesh})"
