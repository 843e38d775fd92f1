"""cpp.la.SparsityPattern: per-row column sets, sorted and made unique by finalize()."""
import numpy as np


class SparsityPattern:
    def __init__(self, comm, maps, bs):
        self.n_rows = maps[0].size_local + maps[0].num_ghosts
        self.n_cols = maps[1].size_local + maps[1].num_ghosts
        self.shape = (self.n_rows, self.n_cols)
        self._rows = [[] for _ in range(self.n_rows)]
        self._final = None

    def insert(self, rows, cols):
        for r in np.asarray(rows).reshape(-1):
            self._rows[int(r)].extend(int(c) for c in np.asarray(cols).reshape(-1))

    def finalize(self):
        uniq = [np.unique(np.asarray(r, dtype=np.int64)) for r in self._rows]
        indptr = np.zeros(self.n_rows + 1, dtype=np.int64)
        indptr[1:] = np.cumsum([u.shape[0] for u in uniq])
        indices = np.concatenate(uniq).astype(np.int32) if self.n_rows else np.zeros(0, dtype=np.int32)
        self._final = (indptr, indices)

    def graph(self):
        assert self._final is not None
        return self._final
