class CellVertices:
    def __init__(self, mesh):
        self.mesh = mesh
