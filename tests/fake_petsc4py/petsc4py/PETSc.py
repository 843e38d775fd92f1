import numpy as np
import scipy.sparse as sp

IntType = np.int32
ScalarType = np.float64
CALLS = []  # (name, kwargs) in call order, inspected by the tests


class LGMap:
    def create(self, indices, bsize=1, comm=None):
        self.indices, self.bsize, self.comm = np.asarray(indices), bsize, comm
        CALLS.append(("LGMap.create", dict(n=len(self.indices), bsize=bsize, dtype=self.indices.dtype)))
        return self


class Mat:
    def __init__(self):
        self.m = None
        self.lgmap = None
        self.assembled = False

    def createAIJ(self, size=None, csr=None, comm=None):
        indptr, indices, data = csr
        CALLS.append(("Mat.createAIJ", dict(size=tuple(size), indptr_dtype=np.asarray(indptr).dtype,
                                            indices_dtype=np.asarray(indices).dtype, nnz=len(indices))))
        self.m = sp.csr_matrix((np.array(data), np.array(indices), np.array(indptr)), shape=tuple(size))
        return self

    def assemble(self):
        self.assembled = True

    def setLGMap(self, rmap, cmap):
        self.lgmap = (rmap, cmap)
        CALLS.append(("Mat.setLGMap", dict(rows=len(rmap.indices), cols=len(cmap.indices), rbs=rmap.bsize, cbs=cmap.bsize)))

    def getValuesCSR(self):
        return self.m.indptr.astype(IntType), self.m.indices.astype(IntType), self.m.data

    def getSize(self):
        return self.m.shape

    def axpy(self, alpha, other):
        self.m = (self.m + alpha * other.m).tocsr()
        self.m.sort_indices()
        CALLS.append(("Mat.axpy", dict(alpha=alpha)))


class Vec:
    def createWithArray(self, array, comm=None):
        self.array = np.array(array, dtype=ScalarType)
        return self

    def getArray(self):
        return self.array

    def axpy(self, alpha, other):
        self.array = self.array + alpha * other.array
