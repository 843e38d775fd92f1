import numpy as np


class IndexMap:
    def __init__(self, comm, size_local, ghosts=None, owners=None, tag=0):
        self.comm = comm
        self.size_local = int(size_local)
        self.ghosts = np.zeros(0, dtype=np.int64) if ghosts is None else np.asarray(ghosts, dtype=np.int64)
        self.owners = np.zeros(0, dtype=np.int32) if owners is None else np.asarray(owners, dtype=np.int32)
        self.num_ghosts = self.ghosts.shape[0]
        assert self.num_ghosts == 0, "serial shim: no ghosts expected"

    def local_to_global(self, idx):
        return np.asarray(idx, dtype=np.int64)

    def global_to_local(self, idx):
        idx = np.asarray(idx, dtype=np.int64)
        return np.where((idx >= 0) & (idx < self.size_local), idx, -1).astype(np.int32)
