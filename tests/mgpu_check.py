#!/usr/bin/env python
"""Multi-GPU correctness check of the sharded block products (run under torchrun on N GPUs; not a pytest file):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py

Every rank builds Π for its block of Γ cells, the ranks exchange the shards (NCCL all-gather-v) and compute their windows
of C = Πᵀ(M_ΓΠ) and MΠ; the windows are gathered again and rank 0 compares them, BIT FOR BIT, with the products of the
unsharded Π computed on one GPU.  Prints 'mgpu_check ok' on success."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from bench import CONFIGS, build_workload, scaled
    from fenicsx_ii_b200.distributed import (ShardedProducts, allgatherv_csr, p1_quadrature_coupling, quadrature_mass_diagonal,
                                             shard_range)
    from fenicsx_ii_b200.interpolation import DeviceRows, device_space
    from fenicsx_ii_b200.sparse import DeviceCSR

    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    scale = float(os.environ.get("MGPU_SCALE", "0.25"))
    cfg = scaled(CONFIGS[5], scale)
    wl = build_workload(cfg, pinned=False)
    V, gamma, K, Q, op = wl["V"], wl["gamma"], wl["K"], wl["Q"], wl["op"]
    nip = K.element.interpolation_points.shape[0]
    n_cells = gamma.geometry.dofmap.shape[0]
    dspace = device_space(V)
    desc = op.device_descriptor(None)
    diag = torch.from_numpy(quadrature_mass_diagonal(gamma, K)).to("cuda")
    M = DeviceCSR.from_scipy(p1_quadrature_coupling(gamma, Q, K))

    def build(c0, c1):
        cells = None if (c0, c1) == (0, n_cells) else np.arange(c0, c1, dtype=np.int32)
        rows = DeviceRows.from_gamma(np.asarray(gamma.geometry.x), np.asarray(gamma.geometry.dofmap), cells,
                                     np.asarray(K.element.interpolation_points), None, (c1 - c0) * nip, desc)
        return rows.assemble(dspace)[0]

    c0, c1 = shard_range(n_cells, world, rank)
    pi_local = build(c0, c1)
    prod = ShardedProducts(V.num_dofs, diag, M, world, rank)
    c_w, mp_w, ip, nnz_a, ip_mp = prod(pi_local)
    c_w2 = prod(pi_local)[0]  # second call: cached plan, no size exchange
    assert c_w2.to_host()[2].tobytes() == c_w.to_host()[2].tobytes()
    # the same blocks through the host-delivery path (Π in blocks of Γ cells, C in row pieces, overlapped downloads):
    # every rank's host arrays must hold exactly its device results
    from fenicsx_ii_b200.distributed import assemble_blocks_to_host

    def pinned(n, dtype):
        return torch.zeros(max(int(n), 1), dtype=dtype).pin_memory().numpy()

    def triple(rows, nnz):
        return pinned(rows + 1, torch.int64), pinned(nnz, torch.int32), pinned(nnz, torch.float64)

    h_pi, h_c, h_mp = triple(pi_local.shape[0], pi_local.nnz), triple(c_w.shape[0], c_w.nnz), triple(mp_w.shape[0], mp_w.nnz)
    r = assemble_blocks_to_host(V, K, op, prod, h_pi, h_c, h_mp, cells_K=np.arange(c0, c1, dtype=np.int32), pi_pieces=2, c_pieces=3)
    r["copy_stream"].synchronize()
    host_ok = True
    for name, got, mat in (("Pi", h_pi, pi_local), ("C", h_c, c_w), ("MPi", h_mp, mp_w)):
        for g, e, what in zip(got, mat.to_host(), ("indptr", "indices", "values")):
            if g.shape != e.shape or g.tobytes() != e.tobytes():
                host_ok = False
                print(f"rank {rank}: host delivery MISMATCH {name}.{what}", flush=True)
    del r
    # gather the windows (they are row shards of C and of MΠ)
    gc = allgatherv_csr(*c_w.device_tensors(), packed=False)
    gm = allgatherv_csr(*mp_w.device_tensors(), packed=False)
    gp = allgatherv_csr(*pi_local.device_tensors())  # packed or not, by size
    ok = True
    if rank == 0:
        full = build(0, n_cells)
        ref_c, ref_mp = ShardedProducts(V.num_dofs, diag, M)(full)[:2]
        for name, got, ref in (("Pi", gp, full.device_tensors()), ("C", gc, ref_c.device_tensors()), ("MPi", gm, ref_mp.device_tensors())):
            for a, b, what in zip(got, ref, ("indptr", "indices", "values")):
                same = a.shape == b.shape and bool(torch.equal(a, b))
                if not same:
                    ok = False
                    print(f"MISMATCH {name}.{what}: shapes {tuple(a.shape)} vs {tuple(b.shape)}", flush=True)
        print(f"world={world} scale={scale} nnz(Pi)={full.nnz} nnz(C)={ref_c.nnz} nnz(MPi)={ref_mp.nnz} windows={prod.plan.windows}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    hflag = torch.tensor([1 if host_ok else 0], device="cuda")
    dist.all_reduce(hflag, op=dist.ReduceOp.MIN)
    flag = torch.minimum(flag, hflag)
    ok = ok and int(hflag.item()) == 1
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("mgpu_check ok" if ok else "mgpu_check FAILED", flush=True)
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
