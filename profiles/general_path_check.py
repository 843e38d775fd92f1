#!/usr/bin/env python
"""Times the general (COO) path — continuous K on Γ, first-visit-wins — next to the fused path on the same Γ.
Run on the GPU box: python profiles/general_path_check.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fenicsx_ii_b200 import Circle, PointwiseTrace  # noqa: E402
from fenicsx_ii_b200 import synthetic as sy  # noqa: E402
from fenicsx_ii_b200.interpolation import device_rows, device_space  # noqa: E402
from fenicsx_ii_b200.mesh import functionspace  # noqa: E402


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return 1e3 * float(np.median(ts))


def main():
    mesh = sy.kuhn_box(64, 64, 64)
    V = functionspace(mesh, ("Lagrange", 1))
    ds = device_space(V)
    for nseg in (10_000, 200_000):
        gamma = sy.polyline_network(nseg, 1 / 64)
        for opname, op in (("trace", PointwiseTrace(gamma)), ("circle8", Circle(gamma, 0.01, degree=15))):
            for kname, K in (("Quadrature10", functionspace(gamma, ("Quadrature", 10))), ("DG1", functionspace(gamma, ("DG", 1))),
                             ("P1 (continuous)", functionspace(gamma, ("Lagrange", 1)))):
                rows = device_rows(K, op)
                csr, st = rows.assemble(ds)
                ms = timeit(lambda: rows.assemble(ds))
                csr, st = rows.assemble(ds)
                print(f"{nseg:7d} segments  {opname:8s} K={kname:16s} rows {K.num_dofs:8d} points {st['n_points']:9d} nnz {csr.nnz:9d} "
                      f"fused={st['fused']}  {ms:8.3f} ms  ({csr.nnz / ms / 1e6:7.2f} G nnz/s)  "
                      f"stages locate {st['ms_locate']:.3f} csr {st['ms_csr']:.3f} (count {st['ms_csr_count']:.3f} fill {st['ms_csr_fill']:.3f})")


if __name__ == "__main__":
    main()
