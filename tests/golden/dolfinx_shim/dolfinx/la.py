"""la.matrix_csr over the shim SparsityPattern; `add` accumulates entry by entry like MatrixCSR::add."""
import enum

import numpy as np


class BlockMode(enum.Enum):
    compact = 0
    expanded = 1


class MatrixCSR:
    def __init__(self, sp, dtype):
        self.indptr, self.indices = sp.graph()
        self.data = np.zeros(self.indices.shape[0], dtype=dtype)
        self.shape = sp.shape

    def add(self, values, rows, cols):
        values = np.asarray(values).reshape(len(rows), len(cols))
        for i, r in enumerate(rows):
            lo, hi = self.indptr[r], self.indptr[r + 1]
            for j, c in enumerate(cols):
                k = lo + np.searchsorted(self.indices[lo:hi], c)
                assert k < hi and self.indices[k] == c, "entry not in the sparsity pattern"
                self.data[k] += values[i, j]

    def scatter_reverse(self):
        pass


def matrix_csr(sp, block_mode=BlockMode.compact, dtype=np.float64):
    return MatrixCSR(sp, dtype)
