"""cpp.la.SparsityPattern: per-row column sets, sorted and made unique by finalize().

Block sizes: rows and columns are inserted as BLOCK indices (interpolation.py:140-163); finalize() expands every
block to its bs_row x bs_col scalar entries, which is what both back ends of the reference end up with (PETSc AIJ
preallocated from the blocked pattern and filled through scalar local indices; dolfinx MatrixCSR.add with
unblocked indices into a blocked matrix)."""
import numpy as np


class SparsityPattern:
    def __init__(self, comm, maps, bs):
        self.bs = (int(bs[0]), int(bs[1]))
        self.n_block_rows = maps[0].size_local + maps[0].num_ghosts
        self.n_block_cols = maps[1].size_local + maps[1].num_ghosts
        self.n_rows = self.n_block_rows * self.bs[0]
        self.n_cols = self.n_block_cols * self.bs[1]
        self.shape = (self.n_rows, self.n_cols)
        self._rows = [[] for _ in range(self.n_block_rows)]
        self._final = None

    def insert(self, rows, cols):
        for r in np.asarray(rows).reshape(-1):
            self._rows[int(r)].extend(int(c) for c in np.asarray(cols).reshape(-1))

    def finalize(self):
        b0, b1 = self.bs
        uniq = []
        for r in self._rows:
            u = np.unique(np.asarray(r, dtype=np.int64))
            scalar_cols = (u[:, None] * b1 + np.arange(b1)[None, :]).reshape(-1)
            uniq.extend([scalar_cols] * b0)
        indptr = np.zeros(self.n_rows + 1, dtype=np.int64)
        indptr[1:] = np.cumsum([u.shape[0] for u in uniq])
        indices = np.concatenate(uniq).astype(np.int32) if self.n_rows else np.zeros(0, dtype=np.int32)
        self._final = (indptr, indices)

    def graph(self):
        assert self._final is not None
        return self._final
