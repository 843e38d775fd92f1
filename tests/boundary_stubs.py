"""Duck-typed stand-ins for the UFL / dolfinx objects the reference's entry points receive (no UFL in this image):
averaged arguments and coefficients with exactly the attributes ``matrix_assembler.py`` / ``assembly.py`` /
``vector_assembler.py`` read."""

from __future__ import annotations

import numpy as np


class Vector:
    def __init__(self, n):
        self.array = np.zeros(n)
        self.scatters = 0

    def scatter_forward(self):
        self.scatters += 1


class Function:
    def __init__(self, space, values=None):
        self.function_space = space
        self.x = Vector(space.num_dofs)
        if values is not None:
            self.x.array[:] = values


class AveragedCoefficient(Function):
    """ufl_operations.py:75-115"""

    def __init__(self, space, restriction_operator, parent_coefficient):
        super().__init__(space)
        self._restriction_operator = restriction_operator
        self._parent_coefficient = parent_coefficient

    @property
    def parent_coefficient(self):
        return self._parent_coefficient

    @property
    def restriction_operator(self):
        return self._restriction_operator


class Argument:
    def __init__(self, space, number):
        self._space, self._number = space, number

    def ufl_function_space(self):
        return self._space

    def number(self):
        return self._number


class AveragedArgument(Argument):
    def __init__(self, space, number, parent_space, restriction_operator):
        super().__init__(space, number)
        self.parent_space, self.restriction_operator = parent_space, restriction_operator


class Form:
    def __init__(self, arguments=(), coefficients=(), tag=None):
        self._arguments, self._coefficients, self.tag = tuple(arguments), tuple(coefficients), tag

    def arguments(self):
        return self._arguments

    def coefficients(self):
        return self._coefficients


class SplitForm(Form):
    """A form that ``apply_replacer`` has already split into parts (passed as ``apply_replacer=lambda a: a.parts``)."""

    def __init__(self, parts):
        super().__init__(parts[0].arguments())
        self.parts = list(parts)
