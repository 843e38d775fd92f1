"""Multi-GPU plumbing: Γ-row sharding, the all-gather-v of row-sharded CSR matrices and the
V-dof-row-sharded block products Πᵀ(M_ΓΠ) and MΠ.

Π assembly needs no communication (every Γ row is independent; Ω is replicated).  The only
collective of the path is the all-gather-v of the CSR arrays of Π's row shards
(``torch.distributed``: NCCL on GPUs, gloo on CPU in the tests).  Shard sizes are exchanged ONCE per
Γ (``GatherPlan``); with a plan the exchange itself runs without any host synchronisation and writes
every shard straight into its slice of the gathered arrays.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


def shard_range(n: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous balanced block [begin, end) of n items for `rank` of `world`."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


@dataclass
class GatherPlan:
    """Sizes of the row shards of one matrix on every rank (one small all-gather + one host read, done once)."""

    world: int
    rank: int
    rows: np.ndarray  # int64[world]
    nnzs: np.ndarray  # int64[world]
    row_off: np.ndarray = field(init=False)  # int64[world+1]
    nnz_off: np.ndarray = field(init=False)

    def __post_init__(self):
        self.row_off = np.concatenate([[0], np.cumsum(self.rows)]).astype(np.int64)
        self.nnz_off = np.concatenate([[0], np.cumsum(self.nnzs)]).astype(np.int64)

    @property
    def n_rows(self) -> int:
        return int(self.row_off[-1])

    @property
    def nnz(self) -> int:
        return int(self.nnz_off[-1])

    def matches(self, n_rows_local: int, nnz_local: int) -> bool:
        return int(self.rows[self.rank]) == n_rows_local and int(self.nnzs[self.rank]) == nnz_local


def plan_allgatherv(n_rows_local: int, nnz_local: int, device, group=None) -> GatherPlan:
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = torch.tensor([n_rows_local, nnz_local], dtype=torch.int64, device=device)
    all_sizes = torch.empty(2 * world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(all_sizes, sizes, group=group)
    all_sizes = all_sizes.view(world, 2).cpu().numpy()  # the one host read of the exchange; cached in the plan
    return GatherPlan(world, rank, all_sizes[:, 0].copy(), all_sizes[:, 1].copy())


def _gatherv_into(out, off, local, plan: GatherPlan, group):
    """out[off[r]:off[r+1]] <- rank r's `local`, for every r, in place.  NCCL: ONE coalesced group of
    broadcasts per array (``all_gather`` with a list of differently sized views); otherwise (gloo, or a torch
    build without that path) one asynchronous broadcast per shard."""
    import torch.distributed as dist

    world, rank = plan.world, plan.rank
    views = [out[int(off[r]):int(off[r + 1])] for r in range(world)]
    if dist.get_backend(group) == "nccl" and all(v.numel() > 0 for v in views):
        try:
            dist.all_gather(views, local, group=group)
            return
        except (RuntimeError, ValueError, TypeError):
            pass
    views[rank].copy_(local)
    ranks = dist.get_process_group_ranks(group) if group is not None else list(range(world))
    work = [dist.broadcast(views[r], src=ranks[r], group=group, async_op=True) for r in range(world) if views[r].numel() > 0]
    for w in work:
        w.wait()


# diagnostics: set to a list to receive CUDA events [start, packed, gathered, unpacked] of every packed exchange
GATHER_TRACE = None


def allgatherv_csr(indptr, indices, values, group=None, packed: "bool | None" = None, plan: "GatherPlan | None" = None,
                   return_plan: bool = False):
    """All-gather row-sharded CSR arrays (torch tensors on the collective's device).

    Every rank passes its shard (indptr int64[m_r+1] starting at 0, indices int32[nnz_r],
    values float64[nnz_r]) and receives the row-concatenated matrix (indptr, indices, values).
    NCCL has no all-gather-v.  The shard sizes come from `plan` (``plan_allgatherv``; built here — one small
    all-gather and one host read — when absent or stale); then

    * packed (default unless a shard is more than 25 % smaller than the largest): row pointers, values and indices of
      a shard travel as ONE byte record padded to the largest shard — a single ``all_gather_into_tensor`` — and are
      copied out of the received records.  Measured on 8 B200 for the 84.3 M entries of config 5 (0.885 GB received
      per rank, profiles/r02_gather_variants_n8.txt): 2.36 ms including the unpack copies, against 3.61 ms for
      per-shard broadcasts written in place (``all_gather`` with uneven views) and 3.19 ms for grouped send/recv;
    * otherwise every shard goes straight into its slice of the preallocated result (no padding, no copies).

    Row pointers are shifted by the shard's entry offset BEFORE they travel, so the gathered pointer array needs
    no scan.  With a valid plan nothing here synchronises with the host."""
    import torch
    import torch.distributed as dist

    dev = indptr.device
    m, n = indptr.numel() - 1, indices.numel()
    if plan is None or not plan.matches(m, n):
        plan = plan_allgatherv(m, n, dev, group)
    world, rank = plan.world, plan.rank
    max_rows, max_nnz = int(plan.rows.max()), int(plan.nnzs.max())
    if packed is None:  # padding waste bounded: balanced shards (the normal case) always take the single collective
        packed = 4 * int(plan.nnzs.min()) >= 3 * max_nnz and 4 * int(plan.rows.min()) >= 3 * max_rows
    g_indptr = torch.empty(plan.n_rows + 1, dtype=torch.int64, device=dev)
    g_indptr[:1] = 0
    g_indices = torch.empty(plan.nnz, dtype=torch.int32, device=dev)
    g_values = torch.empty(plan.nnz, dtype=torch.float64, device=dev)
    trace = GATHER_TRACE if dev.type == "cuda" else None

    def mark():
        if trace is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            trace.append(e)

    mark()
    if packed:
        def pad16(k):
            return (k + 15) // 16 * 16

        o_ptr = pad16(8 * max_nnz)
        o_idx = o_ptr + pad16(8 * max_rows)
        rec = o_idx + pad16(4 * max_nnz)
        send = torch.empty(rec, dtype=torch.uint8, device=dev)  # padding bytes are never read
        send[: 8 * n].view(torch.float64).copy_(values)
        torch.add(indptr[1:], int(plan.nnz_off[rank]), out=send[o_ptr:o_ptr + 8 * m].view(torch.int64))  # global row ends
        send[o_idx:o_idx + 4 * n].view(torch.int32).copy_(indices)
        recv = torch.empty(world * rec, dtype=torch.uint8, device=dev)
        mark()
        dist.all_gather_into_tensor(recv, send, group=group)
        mark()
        recv = recv.view(world, rec)
        for r in range(world):
            mr, nr = int(plan.rows[r]), int(plan.nnzs[r])
            r0, e0 = int(plan.row_off[r]), int(plan.nnz_off[r])
            g_values[e0:e0 + nr].copy_(recv[r, : 8 * nr].view(torch.float64))
            g_indptr[1 + r0:1 + r0 + mr].copy_(recv[r, o_ptr:o_ptr + 8 * mr].view(torch.int64))
            g_indices[e0:e0 + nr].copy_(recv[r, o_idx:o_idx + 4 * nr].view(torch.int32))
        mark()
    else:
        shifted = indptr[1:] + int(plan.nnz_off[rank])  # global row ends of this shard
        _gatherv_into(g_indptr[1:], plan.row_off, shifted, plan, group)
        _gatherv_into(g_indices, plan.nnz_off, indices, plan, group)
        _gatherv_into(g_values, plan.nnz_off, values, plan, group)
    if return_plan:
        return g_indptr, g_indices, g_values, plan
    return g_indptr, g_indices, g_values


class PeerGather:
    """All-gather-v of the row shards of a ``DeviceCSR`` as direct peer-to-peer copies over NVLink (one node).

    Every rank exports CUDA IPC handles of its three CSR arrays; the peers map them once (re-mapping only when a
    pointer changes — the library's allocator hands the same blocks back to equally sized builds) and then PULL every
    shard straight into its slice of the gathered arrays with ``cudaMemcpyAsync``: no padding, no staging copies, the
    copy engines at NVLink rate (770 GB/s peer copy measured on this pool against ~430-475 GB/s for the NCCL all-gather
    plus its pack/unpack copies).  Ordering: a rank enters after its own build has synchronised, so once the (tiny)
    pointer all-gather has been read on the host every peer's shard is complete; a closing all-reduce on the stream keeps
    any rank from releasing its shard before all peers have copied it.  Any failure (IPC not permitted, no peer access)
    disables the path for the process and the caller falls back to the NCCL exchange."""

    def __init__(self, group=None):
        self.group = group
        self.enabled = True
        self.seen = set()       # (rank, block base) whose handle has been exchanged (the same set on every rank)
        self.mapped = {}        # (peer rank, block base on the peer) -> local mapping

    def _fail(self, why):
        import warnings

        self.enabled = False
        warnings.warn(f"peer-to-peer gather disabled, using NCCL: {why}")

    def gather(self, pi_local, plan: GatherPlan):
        """Returns (g_indptr, g_indices, g_values) torch tensors, or None when the path is unavailable (disabled, or
        some rank's arrays do not live in blocks of the library's allocator, e.g. torch tensors: this call then goes
        through NCCL)."""
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _lib

        if not self.enabled:
            return None
        lib = _lib.load()
        world, rank = plan.world, plan.rank
        dev = torch.device("cuda", torch.cuda.current_device())
        p = [C.c_void_p(), C.c_void_p(), C.c_void_p()]
        _lib.check(lib.fxii_csr_device_ptrs(pi_local._h, C.byref(p[0]), C.byref(p[1]), C.byref(p[2])))
        row = []
        for a, v in enumerate(p):  # (block base, offset) per array; base 0: not exportable
            ptr = int(v.value or 0)
            base = C.c_void_p()
            if ptr:
                _lib.check(lib.fxii_alloc_base(C.c_void_p(ptr), C.byref(base)))
            b = int(base.value or 0)
            empty = (plan.nnzs[rank] == 0 and a > 0)
            row += [b if not empty else -1, ptr - b if b else 0]
        mine = torch.tensor(row, dtype=torch.int64, device=dev)
        table = torch.empty(6 * world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(table, mine, group=self.group)
        table = table.view(world, 6).cpu().numpy()  # host read: every rank's shard is complete from here on
        if (table[:, 0::2] == 0).any():
            return None  # the same table on every rank: a collective decision
        bases = table[:, 0::2]
        current = {(r, int(bases[r, a])) for r in range(world) for a in range(3) if bases[r, a] > 0}
        if not current <= self.seen:
            # a block not exchanged before (the allocator hands out a small, recurring set of blocks for equally sized
            # builds, so this happens in the first steps only): every rank takes this branch together and maps the new
            # blocks of its peers; mappings are kept (cudaIpcOpenMemHandle costs about a millisecond)
            ok, why = 1, ""
            try:
                if len(self.mapped) > 96:  # deterministic on every rank: start over
                    for v in self.mapped.values():
                        lib.fxii_ipc_close(C.c_void_p(v))
                    self.mapped.clear()
                    self.seen.clear()
                hbuf = (C.c_ubyte * 192)()
                for a in range(3):
                    if bases[rank, a] > 0:
                        _lib.check(lib.fxii_ipc_export(C.c_void_p(int(bases[rank, a])), C.cast(C.byref(hbuf, 64 * a), C.c_void_p)))
                hs = torch.frombuffer(bytearray(hbuf), dtype=torch.uint8).to(dev)
            except Exception as e:  # noqa: BLE001
                ok, why = 0, f"{type(e).__name__}: {e}"
                hs = torch.zeros(192, dtype=torch.uint8, device=dev)
            allh = torch.empty(192 * world, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(allh, hs, group=self.group)
            allh = allh.view(world, 192).cpu().numpy()
            if ok:
                try:
                    for r in range(world):
                        for a in range(3):
                            b = int(bases[r, a])
                            if r == rank or b <= 0 or (r, b) in self.mapped:
                                continue
                            raw = (C.c_ubyte * 64).from_buffer_copy(allh[r, 64 * a:64 * a + 64].tobytes())
                            out = C.c_void_p()
                            _lib.check(lib.fxii_ipc_open(C.cast(raw, C.c_void_p), C.byref(out)))
                            self.mapped[(r, b)] = int(out.value)
                except Exception as e:  # noqa: BLE001
                    ok, why = 0, f"{type(e).__name__}: {e}"
            flag = torch.tensor([ok], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            if int(flag.item()) == 0:  # some rank cannot export or map: every rank falls back, for good
                self._fail(why if not ok else "a peer rank could not export or map its arrays")
                return None
            self.seen |= current
        g_indptr = torch.empty(plan.n_rows + 1, dtype=torch.int64, device=dev)
        g_indptr[:1] = 0
        g_indices = torch.empty(plan.nnz, dtype=torch.int32, device=dev)
        g_values = torch.empty(plan.nnz, dtype=torch.float64, device=dev)
        stream = _lib.current_stream()
        order = [(rank + d) % world for d in range(world)]  # the own shard first, then ring order: spreads the sources
        for r in order:
            mr, nr = int(plan.rows[r]), int(plan.nnzs[r])
            r0, e0 = int(plan.row_off[r]), int(plan.nnz_off[r])
            src = []
            for a in range(3):
                b = int(bases[r, a])
                src.append((b if r == rank else self.mapped.get((r, b), 0)) + int(table[r, 2 * a + 1]))
            if mr:
                _lib.check(lib.fxii_copy_dev_async(C.c_void_p(g_indptr.data_ptr() + 8 * (1 + r0)), C.c_void_p(src[0] + 8), 8 * mr, stream))
                if e0:
                    g_indptr[1 + r0:1 + r0 + mr] += e0  # shard-local row ends -> global
            if nr:
                _lib.check(lib.fxii_copy_dev_async(C.c_void_p(g_indices.data_ptr() + 4 * e0), C.c_void_p(src[1]), 4 * nr, stream))
                _lib.check(lib.fxii_copy_dev_async(C.c_void_p(g_values.data_ptr() + 8 * e0), C.c_void_p(src[2]), 8 * nr, stream))
        fence = torch.zeros(1, dtype=torch.int32, device=dev)
        dist.all_reduce(fence, group=self.group)  # no rank runs ahead (and frees its shard) before every peer's copies ran
        return g_indptr, g_indices, g_values


def allgatherv_vector(v, group=None):
    """All-gather of per-rank float64 vectors of different lengths."""
    import torch

    plan = plan_allgatherv(v.numel(), v.numel(), v.device, group)
    out = torch.empty(plan.n_rows, dtype=v.dtype, device=v.device)
    _gatherv_into(out, plan.row_off, v, plan, group)
    return out


def balanced_dof_windows(indices_t, n_cols: int, world: int, indptr_t=None):
    """Contiguous ranges of 3D dofs (columns of the gathered Π) of equal SpGEMM work.  The work of row j of
    C = Πᵀ(MΠ) is its number of intermediate products, the sum of the lengths of the Π rows with an entry in column j
    (`indptr_t` given), or, as a proxy, the number of entries in column j.  Every rank computes the same boundaries
    from the same gathered matrix.  (One host read: done once per Γ and kept in the ``ProductPlan``.)"""
    import torch

    if world == 1:
        return [(0, n_cols)]
    if indptr_t is None:
        counts = torch.bincount(indices_t, minlength=n_cols)  # int32 input: no 64-bit copy of the gathered indices
        cum = torch.cumsum(counts, 0)
    else:
        row_len = (indptr_t[1:] - indptr_t[:-1]).to(torch.int32)
        w = torch.repeat_interleave(row_len, row_len.to(torch.int64)).to(torch.float64)  # length of the row of every entry
        cum = torch.cumsum(torch.bincount(indices_t, weights=w, minlength=n_cols), 0).to(torch.int64)
        del w
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


def p1_quadrature_coupling(gamma, Q, K):
    """The coupling block M (Q x K) of the 3D-1D problem, ``inner(z, q) * dx`` with z in the quadrature space K and
    q in P1(Γ) (demos/coupled_poisson.py:92-93): M[q, w] = φ_q(x_w) · weight_w · |segment|, at most 2·nip entries per
    row.  Returns a scipy CSR matrix (what dolfinx would assemble)."""
    import scipy.sparse as sp

    from .quadrature import gauss_jacobi_interval

    x, dm = gamma.geometry.x, gamma.geometry.dofmap
    length = np.linalg.norm(x[dm[:, 1]] - x[dm[:, 0]], axis=1)
    xref, w = gauss_jacobi_interval(K.element.degree)
    xref, w = np.asarray(xref).reshape(-1), np.asarray(w).reshape(-1)
    nip = xref.shape[0]
    qd = np.asarray(Q.dofmap.list)
    kd = np.asarray(K.dofmap.list)
    n_cells = dm.shape[0]
    phi = np.stack([1.0 - xref, xref], axis=0)  # [2, nip]
    rows = np.repeat(qd[:, :, None], nip, axis=2).reshape(-1)  # (cell, a, j)
    cols = np.repeat(kd[:, None, :], 2, axis=1).reshape(-1)
    vals = (phi[None, :, :] * (w[None, None, :] * length[:, None, None])).reshape(-1)
    M = sp.csr_matrix((vals, (rows, cols)), shape=(Q.num_dofs, K.num_dofs))
    M.sum_duplicates()
    assert M.nnz == 2 * nip * n_cells
    return M


@dataclass
class ProductPlan:
    """What the sharded products of one Γ reuse between calls: shard sizes and the 3D-dof windows."""

    gather: "GatherPlan | None"
    windows: list
    setup_ms: float = 0.0


class ShardedProducts:
    """The block products of the coupled 3D-1D problem for a row-sharded Π (matrix_assembler.py:201-258):

        C  = Πᵀ (M_Γ Π)   rows sharded by 3D dof (contiguous windows balanced by Π's column counts)
        MP = M Π          rows sharded by P1(Γ) dof (optional)

    `diag_dev`: the diagonal of the quadrature-element mass M_Γ for ALL Γ rows (float64 CUDA tensor; every rank
    assembles it from the replicated Γ mesh).  `m_dev`: the coupling block M as a DeviceCSR (replicated), or None.
    N = 1: no collective.  N > 1: one all-gather-v of Π's shards, then every rank transposes ITS window of the
    gathered Π and multiplies; M_Γ is applied while B's rows are read (``fxii_spgemm_scaled``), M_ΓΠ is never
    materialised."""

    def __init__(self, n_dofs: int, diag_dev, m_dev=None, world: int = 1, rank: int = 0, group=None):
        self.n_dofs, self.diag, self.m = int(n_dofs), diag_dev, m_dev
        self.world, self.rank, self.group = world, rank, group
        self.plan: "ProductPlan | None" = None
        import os

        self.peer = PeerGather(group) if (world > 1 and os.environ.get("FXII_GATHER", "p2p") == "p2p") else None

    def _make_plan(self, pi_local):
        import torch

        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        if self.world == 1:
            plan = ProductPlan(None, [(0, self.n_dofs)])
        else:
            ptr, idx, val = pi_local.device_tensors()
            g_ptr, g_idx, g_val, gplan = allgatherv_csr(ptr, idx, val, group=self.group, return_plan=True)
            plan = ProductPlan(gplan, balanced_dof_windows(g_idx, self.n_dofs, self.world, indptr_t=g_ptr))
            del g_ptr, g_idx, g_val
        t1.record()
        torch.cuda.synchronize()
        plan.setup_ms = t0.elapsed_time(t1)
        return plan

    def __call__(self, pi_local, events=None):
        """Returns (C_window, MP_window or None, ip_c, nnz_a, ip_mp).  `events`: list receiving CUDA events after
        [start, gather, transpose, spgemm C, spgemm MP]."""
        import torch

        from .sparse import DeviceCSR

        def mark():
            if events is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                events.append(e)

        if self.plan is None or (self.plan.gather is not None and not self.plan.gather.matches(pi_local.shape[0], pi_local.nnz)):
            self.plan = self._make_plan(pi_local)
        mark()
        if self.world > 1:
            got = self.peer.gather(pi_local, self.plan.gather) if self.peer is not None else None
            if got is None:  # NCCL: one padded collective + unpack copies
                ptr, idx, val = pi_local.device_tensors()
                got = allgatherv_csr(ptr, idx, val, group=self.group, plan=self.plan.gather)
            g_ptr, g_idx, g_val = got
            gp = DeviceCSR.wrap_device_arrays(g_ptr, g_idx, g_val, (self.plan.gather.n_rows, self.n_dofs))
        else:
            gp = pi_local
        assert gp.shape[0] == self.diag.numel(), "M_Γ diagonal must cover all Γ rows"
        r0, r1 = self.plan.windows[self.rank]
        mark()
        pt = gp.transpose(r0, r1)  # rows r0..r1 of Πᵀ: this rank's 3D dofs
        mark()
        c, ip = pt.matmult(gp, return_ip=True, row_scale=self.diag)
        mark()
        mp, ip_mp = None, 0
        if self.m is not None:
            q0, q1 = shard_range(self.m.shape[0], self.world, self.rank)
            mp, ip_mp = self.m.matmult(gp, q0, q1, return_ip=True)
        mark()
        return c, mp, ip, pt.nnz, ip_mp


def product_pt_m_p(pi_local, gamma, K, world: int, rank: int, steps: int = 5, flush=None, m_dev=None) -> dict:
    """Times the block products C = Πᵀ (M_Γ Π) (and MΠ when `m_dev` is given) for the row-sharded Π of this rank.
    Returns timings (max over ranks) and the SpGEMM roofline numbers."""
    import torch
    import torch.distributed as dist

    diag_dev = torch.from_numpy(quadrature_mass_diagonal(gamma, K)).to("cuda")  # resident like every other input
    prod = ShardedProducts(pi_local.shape[1], diag_dev, m_dev, world, rank)
    names = ("allgather_ms", "transpose_ms", "spgemm_ms", "spgemm_mp_ms")

    def once():
        ev = []
        if world > 1:
            dist.barrier()  # start together: otherwise the first collective absorbs the ranks' skew
        c, mp, ip, nnz_a, _ = prod(pi_local, events=ev)
        torch.cuda.synchronize()
        t = {n: ev[i].elapsed_time(ev[i + 1]) for i, n in enumerate(names)}
        t["total_ms"] = ev[0].elapsed_time(ev[-1])
        return t, c, mp, ip, nnz_a

    once()
    acc = None
    for _ in range(steps):
        if flush is not None:
            flush.zero_()
        t, c, mp, ip, nnz_a = once()
        acc = t if acc is None else {k: acc[k] + t[k] for k in t}
    avg = {k: v / steps for k, v in acc.items()}
    tt = torch.tensor([avg[k] for k in sorted(avg)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    avg = {k: float(v) for k, v in zip(sorted(avg), tt.tolist())}
    m = prod.plan.windows[rank][1] - prod.plan.windows[rank][0]
    alg_bytes = 12 * nnz_a + 12 * ip + 12 * c.nnz + 16 * (m + 1)
    avg.update(ip=int(ip), nnz_c_local=int(c.nnz), nnz_a=int(nnz_a), nnz_mp_local=int(mp.nnz) if mp is not None else 0,
               plan_setup_ms=prod.plan.setup_ms,
               spgemm_alg_GBps=alg_bytes / (avg["spgemm_ms"] * 1e-3) / 1e9 if avg["spgemm_ms"] > 0 else 0.0,
               spgemm_s=avg["spgemm_ms"] * 1e-3, pit_m_pi_s=avg["total_ms"] * 1e-3,
               association="Pi^T (M Pi), rows sharded by 3D dof (windows balanced by Pi column counts); M_Gamma fused")
    return avg
