"""A 60-line stand-in for petsc4py used ONLY by the tests of the PETSc hand-off (petsc4py is not installable in this
image): it records the arguments of ``Mat.createAIJ(size=, csr=)``, ``Mat.setLGMap`` and ``LGMap.create`` and implements
``getValuesCSR`` / ``axpy`` / ``Vec`` with scipy / numpy so that the callers' control flow can run."""
from . import PETSc  # noqa: F401
