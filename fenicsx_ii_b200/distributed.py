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


def allgatherv_csr(indptr, indices, values, group=None):
    """All-gather row-sharded CSR arrays (torch tensors on the collective's device).

    Every rank passes its shard (indptr int64[m_r+1] starting at 0, indices int32[nnz_r],
    values float64[nnz_r]) and receives the row-concatenated matrix (indptr, indices, values).
    NCCL has no v-variant: shards are padded to the largest shard and trimmed after the gather."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    dev = indptr.device
    sizes = torch.tensor([indptr.numel() - 1, indices.numel()], dtype=torch.int64, device=dev)
    all_sizes = [torch.empty_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    all_sizes = torch.stack(all_sizes).cpu().numpy()
    m_max, nnz_max = int(all_sizes[:, 0].max()), int(all_sizes[:, 1].max())

    def gather(t, n_max, dtype):
        pad = torch.zeros(n_max, dtype=dtype, device=dev)
        pad[: t.numel()] = t
        out = torch.empty(world * n_max, dtype=dtype, device=dev)
        dist.all_gather_into_tensor(out, pad, group=group)
        return out.view(world, n_max)

    counts = gather(indptr[1:] - indptr[:-1], m_max, torch.int64)
    idx = gather(indices, nnz_max, torch.int32)
    val = gather(values, nnz_max, torch.float64)
    rows = torch.cat([counts[r, : all_sizes[r, 0]] for r in range(world)])
    g_indptr = torch.zeros(rows.numel() + 1, dtype=torch.int64, device=dev)
    torch.cumsum(rows, 0, out=g_indptr[1:])
    g_indices = torch.cat([idx[r, : all_sizes[r, 1]] for r in range(world)])
    g_values = torch.cat([val[r, : all_sizes[r, 1]] for r in range(world)])
    return g_indptr, g_indices, g_values


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
    n_dofs = pi_local.shape[1]

    def once():
        t = {}
        e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        e[0].record()
        mp_local = pi_local.copy().scale_rows(diag)  # M_Γ Π (diagonal mass)
        e[1].record()
        if world > 1:
            gp = DeviceCSR.from_device_arrays(*allgatherv_csr(*pi_local.device_tensors()), (pi_local.shape[0] * world, n_dofs))
            gmp = DeviceCSR.from_device_arrays(*allgatherv_csr(*mp_local.device_tensors()), (pi_local.shape[0] * world, n_dofs))
        else:
            gp, gmp = pi_local, mp_local
        e[2].record()
        pt = gp.transpose()
        e[3].record()
        r0, r1 = shard_range(n_dofs, world, rank)
        c, ip = pt.matmult(gmp, r0, r1, return_ip=True)
        e[4].record()
        torch.cuda.synchronize()
        t = dict(scale_ms=e[0].elapsed_time(e[1]), allgather_ms=e[1].elapsed_time(e[2]), transpose_ms=e[2].elapsed_time(e[3]),
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
               association="Pi^T (M Pi), rows sharded by 3D dof")
    return avg
