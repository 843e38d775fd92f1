"""Multi-GPU plumbing: Γ-row sharding, the all-gather of row-sharded CSR matrices and the
V-dof-row-sharded block product Πᵀ(M_ΓΠ).

Π assembly needs no communication (every Γ row is independent; Ω is replicated).  The only
collective of the path is the all-gather-v of the CSR arrays of the row shards, done with
``torch.distributed`` (NCCL on GPUs; gloo on CPU in the tests) on max-padded buffers.
"""

from __future__ import annotations

import numpy as np


def shard_range(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous balanced block [begin, end) of n items for `rank` of `world`."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def allgatherv_csr(indptr, indices, values, group=None, packed: "bool | None" = None):
    """All-gather row-sharded CSR arrays (torch tensors on the collective's device).

    Every rank passes its shard (indptr int64[m_r+1] starting at 0, indices int32[nnz_r],
    values float64[nnz_r]) and receives the row-concatenated matrix (indptr, indices, values).
    NCCL has no all-gather-v.  After one small all-gather of the shard sizes:

    * packed (default while the padded exchange stays below 32 MiB, i.e. latency bound): row counts, values and indices of
      a shard travel as ONE byte record padded to the largest shard — a single all-gather — and are sliced
      out of the received records.  24 small broadcasts cost 3.4 ms on 8 GPUs for 3 MB shards; this is one
      collective.
    * otherwise every shard is broadcast straight into its slice of the preallocated result (no padding)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = indptr.device
    sizes = torch.tensor([indptr.numel() - 1, indices.numel()], dtype=torch.int64, device=dev)
    all_sizes = torch.empty(2 * world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(all_sizes, sizes, group=group)
    all_sizes = all_sizes.view(world, 2).cpu().numpy()
    row_off = np.concatenate([[0], np.cumsum(all_sizes[:, 0])])
    nnz_off = np.concatenate([[0], np.cumsum(all_sizes[:, 1])])
    max_rows, max_nnz = int(all_sizes[:, 0].max()), int(all_sizes[:, 1].max())
    if packed is None:  # small exchanges only (latency bound); large shards go straight into place, unpadded
        packed = 12 * max_nnz * world <= (32 << 20)
    if packed:
        def pad16(n):
            return (n + 15) // 16 * 16

        o_val = pad16(8 * max_rows)
        o_idx = o_val + pad16(8 * max_nnz)
        rec = o_idx + pad16(4 * max_nnz)
        send = torch.zeros(rec, dtype=torch.uint8, device=dev)
        m, n = int(all_sizes[rank, 0]), int(all_sizes[rank, 1])
        send[: 8 * m].view(torch.int64).copy_(indptr[1:] - indptr[:-1])
        send[o_val:o_val + 8 * n].view(torch.float64).copy_(values)
        send[o_idx:o_idx + 4 * n].view(torch.int32).copy_(indices)
        recv = torch.empty(world * rec, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(recv, send, group=group)
        recv = recv.view(world, rec)
        counts = torch.cat([recv[r, : 8 * int(all_sizes[r, 0])].view(torch.int64) for r in range(world)])
        g_values = torch.cat([recv[r, o_val:o_val + 8 * int(all_sizes[r, 1])].view(torch.float64) for r in range(world)])
        g_indices = torch.cat([recv[r, o_idx:o_idx + 4 * int(all_sizes[r, 1])].view(torch.int32) for r in range(world)])
    else:
        counts = torch.empty(int(row_off[-1]), dtype=torch.int64, device=dev)
        g_indices = torch.empty(int(nnz_off[-1]), dtype=torch.int32, device=dev)
        g_values = torch.empty(int(nnz_off[-1]), dtype=torch.float64, device=dev)
        counts[row_off[rank]:row_off[rank + 1]] = indptr[1:] - indptr[:-1]
        g_indices[nnz_off[rank]:nnz_off[rank + 1]] = indices
        g_values[nnz_off[rank]:nnz_off[rank + 1]] = values
        ranks = dist.get_process_group_ranks(group) if group is not None else list(range(world))
        work = []
        for r in range(world):
            src = ranks[r]
            for buf, off in ((counts, row_off), (g_indices, nnz_off), (g_values, nnz_off)):
                if off[r + 1] > off[r]:
                    work.append(dist.broadcast(buf[off[r]:off[r + 1]], src=src, group=group, async_op=True))
        for w in work:
            w.wait()
    g_indptr = torch.zeros(counts.numel() + 1, dtype=torch.int64, device=dev)
    torch.cumsum(counts, 0, out=g_indptr[1:])
    return g_indptr, g_indices, g_values


def allgatherv_vector(v, group=None):
    """All-gather of per-rank float64 vectors of different lengths (same broadcast scheme)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = torch.tensor([v.numel()], dtype=torch.int64, device=v.device)
    all_n = torch.empty(world, dtype=torch.int64, device=v.device)
    dist.all_gather_into_tensor(all_n, n, group=group)
    off = np.concatenate([[0], np.cumsum(all_n.cpu().numpy())])
    out = torch.empty(int(off[-1]), dtype=v.dtype, device=v.device)
    out[off[rank]:off[rank + 1]] = v
    ranks = dist.get_process_group_ranks(group) if group is not None else list(range(world))
    work = [dist.broadcast(out[off[r]:off[r + 1]], src=ranks[r], group=group, async_op=True) for r in range(world)
            if off[r + 1] > off[r]]
    for w in work:
        w.wait()
    return out


def balanced_dof_windows(indices_t, n_cols: int, world: int):
    """Contiguous ranges of 3D dofs (columns of the gathered Π) with equal numbers of Π entries — a
    proxy for equal SpGEMM work (intermediate products of row j of Πᵀ ~ entries in column j of Π).
    Every rank computes the same boundaries from the same gathered matrix."""
    import torch

    if world == 1:
        return [(0, n_cols)]
    counts = torch.bincount(indices_t, minlength=n_cols)  # int32 input: no 64-bit copy of the gathered indices
    cum = torch.cumsum(counts, 0)
    total = int(cum[-1].item())
    targets = torch.tensor([total * (r + 1) // world for r in range(world - 1)], device=cum.device, dtype=cum.dtype)
    cuts = torch.searchsorted(cum, targets, right=False).add_(1).clamp_(max=n_cols).cpu().tolist()
    edges = [0] + cuts + [n_cols]
    for i in range(1, len(edges)):  # keep the ranges monotone even for degenerate distributions
        edges[i] = max(edges[i], edges[i - 1])
    return [(edges[r], edges[r + 1]) for r in range(world)]


def quadrature_mass_diagonal(gamma, K) -> np.ndarray:
    """Diagonal of the quadrature-element mass matrix M_Γ on Γ (what dolfinx assembles for
    ``inner(z, w) * dx`` with a quadrature element, demos/coupled_poisson.py:164-166): Gauss weight
    times segment length, one entry per K dof."""
    from .quadrature import gauss_jacobi_interval

    x, dm = gamma.geometry.x, gamma.geometry.dofmap
    length = np.linalg.norm(x[dm[:, 1]] - x[dm[:, 0]], axis=1)
    _, w = gauss_jacobi_interval(K.element.degree)
    diag = np.zeros(K.num_dofs)
    diag[K.dofmap.list.reshape(-1)] = (length[:, None] * w[None, :]).reshape(-1)
    return diag


def product_pt_m_p(pi_local, gamma, K, world: int, rank: int, steps: int = 5, flush=None) -> dict:
    """Times the block product C = Πᵀ (M_Γ Π) for the row-sharded Π of this rank.

    N = 1: transpose + row scaling + SpGEMM on the local matrix.  N > 1: all-gather of Π's and
    M_ΓΠ's shards (NCCL), then every rank computes the rows of C of its contiguous range of 3D
    dofs.  Returns timings (max over ranks) and roofline numbers for the SpGEMM."""
    import torch
    import torch.distributed as dist

    from .sparse import DeviceCSR

    diag = quadrature_mass_diagonal(gamma, K)
    diag_dev = torch.from_numpy(diag).to("cuda")  # resident like every other input of the timed product
    n_dofs = pi_local.shape[1]

    def once():
        import torch as _t

        e = [_t.cuda.Event(enable_timing=True) for _ in range(5)]
        if world > 1:
            dist.barrier()  # start together: otherwise the first collective absorbs the ranks' skew
        e[0].record()
        if world > 1:
            # the one exchange of the path: Π's row shards (and the mass diagonal) to every rank
            g_ptr, g_idx, g_val = allgatherv_csr(*pi_local.device_tensors())
            gp = DeviceCSR.from_device_arrays(g_ptr, g_idx, g_val, (pi_local.shape[0] * world, n_dofs))
            gdiag = allgatherv_vector(diag_dev)  # stays on the device: scale_rows takes it in place
            r0, r1 = balanced_dof_windows(g_idx, n_dofs, world)[rank]
            del g_ptr, g_idx, g_val
        else:
            gp, gdiag = pi_local, diag_dev
            r0, r1 = 0, n_dofs
        e[1].record()
        gmp = gp.copy().scale_rows(gdiag)  # M_Γ Π (diagonal quadrature mass)
        e[2].record()
        pt = gp.transpose(r0, r1)  # only the rows of Πᵀ this rank owns (3D dofs r0..r1)
        e[3].record()
        c, ip = pt.matmult(gmp, return_ip=True)
        e[4].record()
        _t.cuda.synchronize()
        t = dict(allgather_ms=e[0].elapsed_time(e[1]), scale_ms=e[1].elapsed_time(e[2]), transpose_ms=e[2].elapsed_time(e[3]),
                 spgemm_ms=e[3].elapsed_time(e[4]), total_ms=e[0].elapsed_time(e[4]))
        return t, c, ip, pt.nnz, (r1 - r0)

    once()
    acc = None
    for _ in range(steps):
        if flush is not None:
            flush.zero_()
        t, c, ip, nnz_a, m = once()
        acc = t if acc is None else {k: acc[k] + t[k] for k in t}
    avg = {k: v / steps for k, v in acc.items()}
    tt = torch.tensor([avg[k] for k in sorted(avg)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    avg = {k: float(v) for k, v in zip(sorted(avg), tt.tolist())}
    alg_bytes = 12 * nnz_a + 12 * ip + 12 * c.nnz + 16 * (m + 1)
    avg.update(ip=int(ip), nnz_c_local=int(c.nnz), nnz_a=int(nnz_a),
               spgemm_alg_GBps=alg_bytes / (avg["spgemm_ms"] * 1e-3) / 1e9 if avg["spgemm_ms"] > 0 else 0.0,
               spgemm_s=avg["spgemm_ms"] * 1e-3, pit_m_pi_s=avg["total_ms"] * 1e-3,
               association="Pi^T (M Pi), rows sharded by 3D dof (windows balanced by Pi column counts)")
    return avg
