"""CPU, world size 2 (gloo): the all-gather-v of row-sharded CSR matrices and the row-shard
arithmetic used by the multi-GPU block product."""

import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fenicsx_ii_b200.distributed import allgatherv_csr, plan_allgatherv, shard_range


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, shards, out_dir, packed):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = shards[rank]
        arrays = (torch.from_numpy(m.indptr.astype(np.int64)), torch.from_numpy(m.indices.astype(np.int32)),
                  torch.from_numpy(m.data.astype(np.float64)))
        ip, ij, iv, plan = allgatherv_csr(*arrays, packed=packed, return_plan=True)
        # the cached plan (no size exchange, no host read) gives the same matrix; a stale plan is rebuilt
        ip2, ij2, iv2 = allgatherv_csr(*arrays, packed=packed, plan=plan)
        assert torch.equal(ip, ip2) and torch.equal(ij, ij2) and torch.equal(iv, iv2)
        stale = plan_allgatherv(m.shape[0] + 1, m.nnz, arrays[0].device)
        ip3, _, _ = allgatherv_csr(*arrays, packed=packed, plan=stale)
        assert torch.equal(ip, ip3)
        assert plan.n_rows == sum(s.shape[0] for s in shards) and plan.nnz == sum(s.nnz for s in shards)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), indptr=ip.numpy(), indices=ij.numpy(), data=iv.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("packed", [None, True, False])  # automatic choice, one padded all-gather, per-shard broadcasts
@pytest.mark.parametrize("rows", [(7, 12), (5, 0), (1, 1), (4, 3, 3)])  # the last: 10 rows on 3 ranks (uneven shards)
def test_allgatherv_csr_gloo(tmp_path, rows, packed):
    world = len(rows)
    rng = np.random.default_rng(5)
    shards = []
    for r, m in enumerate(rows):
        M = sp.random(m, 9, density=0.3 + 0.3 * r, random_state=r + 1, format="csr")
        M.sort_indices()
        shards.append(M)
    port = _free_port()
    mp.spawn(_worker, args=(world, port, shards, str(tmp_path), packed), nprocs=world, join=True)
    full = sp.vstack(shards).tocsr()
    full.sort_indices()
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        np.testing.assert_array_equal(z["indptr"], full.indptr)
        np.testing.assert_array_equal(z["indices"], full.indices)
        np.testing.assert_array_equal(z["data"], full.data)


def test_shard_range_partitions():
    for n in (0, 1, 7, 16, 17_000_003):
        for world in (1, 2, 3, 8):
            edges = [shard_range(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for a, b in zip(edges[:-1], edges[1:]):
                assert a[1] == b[0]
            sizes = [e[1] - e[0] for e in edges]
            assert max(sizes) - min(sizes) <= 1


def test_balanced_dof_windows_cover_and_balance():
    from fenicsx_ii_b200.distributed import balanced_dof_windows

    rng = np.random.default_rng(1)
    n_cols = 5000
    # entries concentrated in a narrow band of columns (a vessel in a big mesh)
    idx = torch.from_numpy(np.concatenate([rng.integers(1000, 1200, 90_000), rng.integers(0, n_cols, 10_000)]).astype(np.int32))
    for world in (1, 2, 3, 8):
        w = balanced_dof_windows(idx, n_cols, world)
        assert w[0][0] == 0 and w[-1][1] == n_cols and all(a[1] == b[0] for a, b in zip(w[:-1], w[1:]))
        loads = [int(((idx >= a) & (idx < b)).sum()) for a, b in w]
        assert sum(loads) == idx.numel()
        if world > 1:
            assert max(loads) <= 1.3 * idx.numel() / world + 600


def test_balanced_dof_windows_by_intermediate_products():
    """With the row pointers given the windows balance the SpGEMM work (sum over a column's entries of the length of
    their rows), not the entry counts."""
    from fenicsx_ii_b200.distributed import balanced_dof_windows

    rng = np.random.default_rng(3)
    n_rows, n_cols = 4000, 3000
    # rows touching the low columns are 10x longer than the rows touching the high columns
    lens = np.where(np.arange(n_rows) < n_rows // 2, 40, 4)
    cols = [np.sort(rng.choice(n_cols // 2, size=k, replace=False)) + (0 if i < n_rows // 2 else n_cols // 2)
            for i, k in enumerate(lens)]
    indptr = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64))
    idx = torch.from_numpy(np.concatenate(cols).astype(np.int32))
    work = np.zeros(n_cols)
    np.add.at(work, idx.numpy(), np.repeat(lens, lens))
    for world in (2, 3, 8):
        w = balanced_dof_windows(idx, n_cols, world, indptr_t=indptr)
        assert w[0][0] == 0 and w[-1][1] == n_cols and all(a[1] == b[0] for a, b in zip(w[:-1], w[1:]))
        loads = np.array([work[a:b].sum() for a, b in w])
        assert loads.max() <= 1.15 * work.sum() / world + work.max()
        by_count = balanced_dof_windows(idx, n_cols, world)
        loads_c = np.array([work[a:b].sum() for a, b in by_count])
        assert loads.max() <= loads_c.max() + 1e-9


@pytest.mark.parametrize("world,n_sub", [(1, 4), (2, 2), (4, 3), (8, 2)])
def test_row_pieces_refine_the_rank_windows(world, n_sub):
    """ShardedProducts.to_host cuts a rank's window of 3D dofs into `n_sub` pieces of equal work by asking for
    world*n_sub windows: every n_sub-th cut of the fine partition must be a cut of the coarse one (same formula), so the
    pieces of rank r tile exactly the window of rank r."""
    from fenicsx_ii_b200.distributed import balanced_dof_windows

    rng = np.random.default_rng(world * 10 + n_sub)
    n_rows, n_cols = 4000, 900
    lens = rng.integers(1, 12, size=n_rows)
    indptr = torch.from_numpy(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64))
    cols = np.concatenate([np.sort(rng.choice(n_cols, size=k, replace=False)) for k in lens]).astype(np.int32)
    # an uneven column distribution: the work is concentrated in a third of the dofs
    cols = (cols.astype(np.int64) ** 2 // n_cols).astype(np.int32)
    idx = torch.from_numpy(cols)
    coarse = balanced_dof_windows(idx, n_cols, world, indptr_t=indptr)
    fine = balanced_dof_windows(idx, n_cols, world * n_sub, indptr_t=indptr)
    assert fine[0][0] == 0 and fine[-1][1] == n_cols
    assert all(fine[i][1] == fine[i + 1][0] for i in range(len(fine) - 1))
    for r in range(world):
        mine = fine[r * n_sub:(r + 1) * n_sub]
        assert (mine[0][0], mine[-1][1]) == coarse[r]
