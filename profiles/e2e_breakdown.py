#!/usr/bin/env python
"""Where the end-to-end time of one create_interpolation_matrix call goes (host wall clock between
device synchronisations; run on the GPU box):  python profiles/e2e_breakdown.py [--config 2]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from fenicsx_ii_b200.interpolation import create_interpolation_matrix, device_rows, device_space  # noqa: E402


def timeit(fn, reps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return 1e3 * float(np.median(ts)), 1e3 * float(np.min(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    args = ap.parse_args()
    cfg = bench.CONFIGS[args.config]
    wl = bench.build_workload(cfg, pinned=True)
    mesh, V, gamma, K, op, seg_r = (wl[k] for k in ("mesh", "V", "gamma", "K", "op", "seg_r"))
    ds = device_space(V)
    rows = device_rows(K, op)
    csr, st = rows.assemble(ds)
    nnz = csr.nnz
    indptr_h = torch.empty(K.num_dofs + 1, dtype=torch.int64).pin_memory().numpy()
    indices_h = torch.empty(max(nnz, 1), dtype=torch.int32).pin_memory().numpy()
    values_h = torch.empty(max(nnz, 1), dtype=torch.float64).pin_memory().numpy()
    hold = {}

    def make_rows():
        hold["rows"] = device_rows(K, op)

    def assemble_reused():
        hold["csr"], _ = rows.assemble(ds)

    def assemble_fresh():
        r = device_rows(K, op)
        r.set_nnz_hint(nnz)
        hold["csr"], _ = r.assemble(ds)

    def download():
        csr.to_host_into(indptr_h, indices_h, values_h)

    def assemble_host_reused():
        rows.assemble_to_host(ds, indptr_h, indices_h, values_h)

    def full_host_out():
        create_interpolation_matrix(V, K, op, host_out=(indptr_h, indices_h, values_h))

    def full():
        A, _, _ = create_interpolation_matrix(V, K, op)
        A.device.to_host_into(indptr_h, indices_h, values_h)

    print(f"config {args.config}: rows {K.num_dofs}, nnz {nnz}, CSR bytes {8 * (K.num_dofs + 1) + 12 * nnz}")
    for name, fn in [("device_rows (H2D of Γ + operator)", make_rows), ("assemble, reused handle", assemble_reused),
                     ("device_rows + assemble (fresh handle, hint)", assemble_fresh), ("CSR download (D2H)", download),
                     ("create_interpolation_matrix + download", full),
                     ("assemble_to_host, reused handle", assemble_host_reused),
                     ("create_interpolation_matrix(host_out=...): one C call", full_host_out)]:
        med, best = timeit(fn)
        print(f"  {name:48s} median {med:8.4f} ms   min {best:8.4f} ms")


if __name__ == "__main__":
    main()
