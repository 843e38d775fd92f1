"""Quadrature record and the tables the reduction operators need (host side, tiny, once per
operator).  Mirrors the names of the reference's ``quadrature.py`` (``Quadrature`` :11-20,
``rotation_matrix`` :35-69, ``compute_disk_quadrature`` :72-103); basix is not required:
``basix.make_quadrature(interval, degree)`` is the Gauss-Legendre rule with ``(degree+2)//2``
points mapped to [0,1].
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import numpy.typing as npt

__all__ = ["Quadrature", "compute_disk_quadrature", "gauss_jacobi_interval", "rotation_matrix"]


@dataclass
class Quadrature:
    name: str
    """Name of the quadrature rule"""
    points: npt.NDArray[np.floating]
    """Points in the rule (physical space)"""
    weights: npt.NDArray[np.floating]
    """Weights of the quadrature points"""
    scales: npt.NDArray[np.floating]
    """Scaling factors for the quadrature points"""


def gauss_jacobi_interval(degree: int) -> tuple[npt.NDArray[np.float64], npt.NDArray[np.float64]]:
    """Rule exact for polynomials of ``degree`` on [0,1]: points ``(m,1)`` ascending, weights ``(m,)``
    summing to one, ``m = (degree+2)//2`` (what basix returns for the interval)."""
    m = (int(degree) + 2) // 2
    nodes, weights = np.polynomial.legendre.leggauss(m)
    return ((nodes + 1.0) / 2.0).reshape(m, 1), weights / 2.0


def compute_disk_quadrature(n: int) -> tuple[npt.NDArray[np.float64], npt.NDArray[np.float64]]:
    """``n*n``-point rule of accuracy ``2n-1`` on the unit disk: ``n`` chords at
    ``x = cos(k pi/(n+1))``, each carrying an ``n``-point Gauss rule, weighted by
    ``pi/(n+1) sin(k pi/(n+1))`` (Bojanov & Petrova 1998, DOI 10.1007/s002110050358)."""
    s, ws = gauss_jacobi_interval(2 * n - 1)
    s = s[:, 0]
    k = np.arange(1, n + 1)
    eta = np.cos(k * np.pi / (n + 1))
    chord_weight = np.pi / (n + 1) * np.sin(k * np.pi / (n + 1))
    half = np.sqrt(1 - eta**2)
    # chord k runs from -half to +half; the reference parametrises it from the top down
    y = s[None, :] * (-half)[:, None] + (1 - s)[None, :] * half[:, None]
    w = ((half - (-half))[:, None] * ws[None, :]) * chord_weight[:, None]
    pts = np.stack([np.repeat(eta, s.shape[0]), y.reshape(-1)], axis=1)
    return pts, w.reshape(-1)


def rotation_matrix(n: npt.NDArray[np.floating]) -> npt.NDArray[np.floating]:
    """Rotation(s) taking ``ez`` to the unit normal(s) ``n`` by Rodrigues' formula; the rotation
    axis falls back to ``ez`` when ``ez x n`` vanishes (n parallel to ez)."""
    n = np.asarray(n, dtype=np.float64).reshape(-1, 3)
    n = n / np.linalg.norm(n, axis=1)[:, None]
    ez = np.array([0.0, 0.0, 1.0])
    axis = np.cross(ez, n)
    small = np.isclose(np.linalg.norm(axis, axis=1), 0)
    axis[small] = ez
    axis /= np.linalg.norm(axis, axis=1)[:, None]
    c = n @ ez
    s = np.sqrt(1 - c**2)
    cross = np.zeros((n.shape[0], 3, 3))
    cross[:, 0, 1], cross[:, 0, 2] = -axis[:, 2], axis[:, 1]
    cross[:, 1, 0], cross[:, 1, 2] = axis[:, 2], -axis[:, 0]
    cross[:, 2, 0], cross[:, 2, 1] = -axis[:, 1], axis[:, 0]
    eye = np.broadcast_to(np.eye(3), cross.shape)
    outer = axis[:, :, None] * axis[:, None, :]
    return eye * c[:, None, None] + cross * s[:, None, None] + outer * (1 - c)[:, None, None]
