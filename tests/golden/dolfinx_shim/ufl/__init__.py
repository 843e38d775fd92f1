"""ufl stand-in: three symbolic markers consumed by the shim's Expression."""
from . import geometry  # noqa: F401


class SpatialCoordinate:
    def __init__(self, mesh):
        self.mesh = mesh


class TestFunction:
    def __init__(self, V):
        self.V = V


class Form:  # only referenced in annotations
    pass
