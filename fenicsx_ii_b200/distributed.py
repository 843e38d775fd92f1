"""Multi-GPU plumbing: Γ-row sharding, the all-gather-v of row-sharded CSR matrices and the
V-dof-row-sharded block products Πᵀ(M_ΓΠ) and MΠ.

Π assembly needs no communication (every Γ row is independent; Ω is replicated).  The only
collective of the path is the all-gather-v of the CSR arrays of Π's row shards
(``torch.distributed``: NCCL on GPUs, gloo on CPU in the tests).  Shard sizes are exchanged ONCE per
Γ (``GatherPlan``); with a plan the exchange itself runs without any host synchronisation and writes
every shard straight into its slice of the gathered arrays.
"""

from __future__ import annotations

import time
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

    Set-up, once per ``GatherPlan`` (collective): every rank allocates persistent SEND buffers for its shard in the
    library's allocator, exports their CUDA IPC handles, maps the peers' buffers, and allocates the persistent gathered
    arrays.  Every exchange then runs WITHOUT host synchronisation, allocation or pointer exchange, in four launches:

      1. one kernel copies the local shard into the send buffers (row pointers shifted to their global values);
      2. a one-element NCCL all-reduce on the stream: when it completes every rank's send buffers are complete;
      3. one kernel PULLS all shards — peers' over NVLink, the own one locally — straight into their slices of the
         gathered CSR (no padding, no staging: the SMs' loads keep the links full, 570-650 GB/s per peer measured);
      4. a second one-element all-reduce: no rank refills its send buffers before every peer has read them.

    Why not NCCL's all-gather: it needs equal sizes, i.e. a padded record plus unpack copies (2.35 ms for the 0.885 GB a
    rank receives on 8 GPUs, profiles/r02_gather_variants_n8.txt), and its uneven variants are slower still (3.2-3.6 ms).
    A first peer-to-peer version that exchanged the current pointers through a host-read all-gather every step spent
    1.7-3.5 ms per exchange in that handshake on an 8-rank box.  The gathered arrays belong to this object and are
    overwritten by the next exchange.  Any failure of the set-up (IPC not permitted, no peer access) disables the path
    on every rank and the caller uses the NCCL exchange."""

    def __init__(self, group=None):
        self.group = group
        self.enabled = True
        self.plan = None
        self.send = None      # local send buffers (device pointers: indptr, indices, values)
        self.peer = None      # [world][3] mapped pointers of every rank's send buffers (own rank: the local pointers)
        self.out = None       # persistent gathered tensors
        self.fence = None

    def _fail(self, why):
        import warnings

        self.enabled = False
        warnings.warn(f"peer-to-peer gather disabled, using NCCL: {why}")

    def _release(self):
        import ctypes as C

        from . import _lib

        lib = _lib.load()
        if self.peer is not None:
            for r, ptrs in enumerate(self.peer):
                if r != self.plan.rank:
                    for v in ptrs:
                        if v:
                            lib.fxii_ipc_close(C.c_void_p(v))
        if self.send is not None:
            for v in self.send:
                if v:
                    lib.fxii_device_free(C.c_void_p(v), _lib.current_stream())
        self.send = self.peer = self.out = self.plan = None

    def _setup(self, plan: GatherPlan) -> bool:
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _lib

        lib = _lib.load()
        world, rank = plan.world, plan.rank
        dev = torch.device("cuda", torch.cuda.current_device())
        torch.cuda.synchronize()
        if self.plan is not None:
            dist.barrier(group=self.group)  # nobody still reads the old send buffers
            self._release()
        self.plan = plan
        m, n = int(plan.rows[rank]), int(plan.nnzs[rank])
        sizes = [8 * max(m, 1), 4 * max(n, 1), 8 * max(n, 1)]
        ok, why = 1, ""
        hbuf = (C.c_ubyte * 192)()
        try:
            self.send = []
            for a, nb in enumerate(sizes):
                p = C.c_void_p()
                _lib.check(lib.fxii_device_alloc(nb, _lib.current_stream(), C.byref(p)))
                self.send.append(int(p.value))
                _lib.check(lib.fxii_ipc_export(p, C.cast(C.byref(hbuf, 64 * a), C.c_void_p)))
        except Exception as e:  # noqa: BLE001
            ok, why = 0, f"{type(e).__name__}: {e}"
        hs = torch.frombuffer(bytearray(hbuf), dtype=torch.uint8).to(dev)
        allh = torch.empty(192 * world, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allh, hs, group=self.group)
        allh = allh.view(world, 192).cpu().numpy()
        self.peer = [[0, 0, 0] for _ in range(world)]
        if ok:
            try:
                for r in range(world):
                    for a in range(3):
                        if r == rank:
                            self.peer[r][a] = self.send[a]
                        else:
                            raw = (C.c_ubyte * 64).from_buffer_copy(allh[r, 64 * a:64 * a + 64].tobytes())
                            out = C.c_void_p()
                            _lib.check(lib.fxii_ipc_open(C.cast(raw, C.c_void_p), C.byref(out)))
                            self.peer[r][a] = int(out.value)
            except Exception as e:  # noqa: BLE001
                ok, why = 0, f"{type(e).__name__}: {e}"
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) == 0:  # some rank cannot export or map: every rank falls back, for good
            self._fail(why if not ok else "a peer rank could not export or map its send buffers")
            return False
        g_indptr = torch.empty(plan.n_rows + 1, dtype=torch.int64, device=dev)
        g_indptr[:1] = 0
        self.out = (g_indptr, torch.empty(plan.nnz, dtype=torch.int32, device=dev), torch.empty(plan.nnz, dtype=torch.float64, device=dev))
        self.fence = torch.zeros(1, dtype=torch.int32, device=dev)
        # the pull: every shard from its owner's send buffers into its slice (own shard first, then ring order)
        dst, src, nbytes = [], [], []
        for d in range(world):
            r = (rank + d) % world
            mr, nr = int(plan.rows[r]), int(plan.nnzs[r])
            r0, e0 = int(plan.row_off[r]), int(plan.nnz_off[r])
            if mr:
                dst.append(g_indptr.data_ptr() + 8 * (1 + r0)), src.append(self.peer[r][0]), nbytes.append(8 * mr)
            if nr:
                dst.append(self.out[1].data_ptr() + 4 * e0), src.append(self.peer[r][1]), nbytes.append(4 * nr)
                dst.append(self.out[2].data_ptr() + 8 * e0), src.append(self.peer[r][2]), nbytes.append(8 * nr)
        k = len(dst)
        self.pull = (k, (C.c_void_p * k)(*dst), (C.c_void_p * k)(*src), (C.c_int64 * k)(*nbytes))
        return True

    def gather(self, pi_local, plan: GatherPlan):
        """Returns (g_indptr, g_indices, g_values) — persistent tensors owned by this object, overwritten by the next
        call — or None when the path is unavailable."""
        import ctypes as C

        import torch.distributed as dist

        from . import _lib

        if not self.enabled:
            return None
        if plan is not self.plan and not self._setup(plan):
            return None
        lib = _lib.load()
        rank = plan.rank
        m, n = int(plan.rows[rank]), int(plan.nnzs[rank])
        p = [C.c_void_p(), C.c_void_p(), C.c_void_p()]
        _lib.check(lib.fxii_csr_device_ptrs(pi_local._h, C.byref(p[0]), C.byref(p[1]), C.byref(p[2])))
        stream = _lib.current_stream()
        # 1. local shard -> send buffers; row pointers become global row ends (shard-local ends + entry offset)
        e0 = int(plan.nnz_off[rank])
        segs = [(self.send[0], int(p[0].value or 0) + 8, 8 * m, e0), (self.send[1], int(p[1].value or 0), 4 * n, 0),
                (self.send[2], int(p[2].value or 0), 8 * n, 0)]
        segs = [s for s in segs if s[2] > 0]
        k = len(segs)
        if k:
            _lib.check(lib.fxii_copy_segments_async(k, (C.c_void_p * k)(*[s[0] for s in segs]), (C.c_void_p * k)(*[s[1] for s in segs]),
                                                    (C.c_int64 * k)(*[s[2] for s in segs]), (C.c_int64 * k)(*[s[3] for s in segs]), stream))
        dist.all_reduce(self.fence, group=self.group)  # 2. every rank's send buffers are complete
        kp, dst, src, nbytes = self.pull
        if kp:
            _lib.check(lib.fxii_copy_segments_async(kp, dst, src, nbytes, None, stream))  # 3. pull all shards into place
        dist.all_reduce(self.fence, group=self.group)  # 4. nobody refills its send buffers before everyone has read them
        return self.out


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


    # ---- host delivery: downloads overlapped with the remaining products -----------------------------------------
    def _sub_windows(self, gp, n_sub: int):
        """This rank's window of 3D dofs cut into `n_sub` contiguous pieces of equal SpGEMM work (cached in the plan)."""
        r0, r1 = self.plan.windows[self.rank]
        key = ("sub", n_sub)
        cache = self.plan.__dict__.setdefault("_sub", {})
        if key not in cache:
            if n_sub <= 1:
                cache[key] = [(r0, r1)]
            else:
                ptr, idx, _ = gp.device_tensors()
                fine = balanced_dof_windows(idx, self.n_dofs, self.world * n_sub, indptr_t=ptr)
                mine = fine[self.rank * n_sub:(self.rank + 1) * n_sub]
                # the fine cuts at multiples of n_sub are the coarse cuts (same formula); pin the ends anyway
                mine[0] = (r0, mine[0][1])
                mine[-1] = (mine[-1][0], r1)
                cache[key] = [(a, max(a, b)) for a, b in mine]
        return cache[key]

    def to_host(self, pi_local, host_c, host_mp=None, n_sub: int = 4, copy_stream=None, marks=None):
        """The products of ``__call__`` delivered into caller-owned host CSR arrays ``(indptr int64, indices int32,
        values float64)``, ideally pinned, with the device→host copies overlapped with the products still to come
        (what bounds the end-to-end time of a block assembly is the PCIe link: 3 GB for config 5):

          * MΠ is computed first and its download queued on `copy_stream`;
          * this rank's window of C = Πᵀ(M_ΓΠ) is computed in `n_sub` row pieces of equal work; a piece's row
            pointers are shifted to the window's entry numbering on the device and its download is queued as soon as
            it is finished, while the next piece is transposed and multiplied on the main stream.

        The rows, their order and their arithmetic are those of ``__call__`` (the result is bit-identical).
        Returns ``dict(nnz_c, nnz_mp, rows_c, keep)``; synchronise `copy_stream` before reading the arrays and keep
        ``keep`` alive until then.  ``ValueError`` if a host array is too small."""
        import torch

        from .sparse import DeviceCSR

        if copy_stream is None:
            copy_stream = torch.cuda.Stream()
        if self.plan is None or (self.plan.gather is not None and not self.plan.gather.matches(pi_local.shape[0], pi_local.nnz)):
            self.plan = self._make_plan(pi_local)
        if self.world > 1:
            got = self.peer.gather(pi_local, self.plan.gather) if self.peer is not None else None
            if got is None:
                ptr, idx, val = pi_local.device_tensors()
                got = allgatherv_csr(ptr, idx, val, group=self.group, plan=self.plan.gather)
            gp = DeviceCSR.wrap_device_arrays(*got, (self.plan.gather.n_rows, self.n_dofs))
        else:
            gp = pi_local
        keep = [gp]

        def queue_download(mat, bufs, name):
            done = torch.cuda.Event(enable_timing=marks is not None)
            done.record()
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done)
                if marks is not None:
                    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    b0.record()
                mat.to_host_into(*bufs, sync=False)
                if marks is not None:
                    b1.record()
                    marks.extend([(name + " ready (main stream)", done), (name + " D2H begin", b0), (name + " D2H end", b1)])
            keep.append(mat)

        nnz_mp = 0
        if self.m is not None and host_mp is not None:
            q0, q1 = shard_range(self.m.shape[0], self.world, self.rank)
            mp = self.m.matmult(gp, q0, q1)
            nnz_mp = mp.nnz
            if host_mp[0].shape[0] < mp.shape[0] + 1 or host_mp[1].shape[0] < nnz_mp or host_mp[2].shape[0] < nnz_mp:
                raise ValueError(f"host arrays for MΠ too small: {mp.shape[0]} rows, {nnz_mp} entries")
            queue_download(mp, (host_mp[0][: mp.shape[0] + 1], host_mp[1][:nnz_mp], host_mp[2][:nnz_mp]), "MPi")
        r0, r1 = self.plan.windows[self.rank]
        if host_c[0].shape[0] < r1 - r0 + 1:
            raise ValueError(f"host row pointers for C too small: {r1 - r0} rows")
        off = 0
        for w, (a, b) in enumerate(self._sub_windows(gp, n_sub)):
            if b <= a:
                continue
            pt = gp.transpose(a, b)
            c = pt.matmult(gp, row_scale=self.diag)
            del pt
            if host_c[1].shape[0] < off + c.nnz or host_c[2].shape[0] < off + c.nnz:
                raise ValueError(f"host arrays for C too small: at least {off + c.nnz} entries")
            if off:
                c.device_tensors()[0].add_(off)  # row pointers in the numbering of the whole window
            # the piece's first pointer equals the previous piece's last one: the one-element overlap is consistent
            queue_download(c, (host_c[0][a - r0:b - r0 + 1], host_c[1][off:off + c.nnz], host_c[2][off:off + c.nnz]), f"C[{w}]")
            off += c.nnz
        return dict(nnz_c=off, nnz_mp=nnz_mp, rows_c=r1 - r0, keep=keep, copy_stream=copy_stream)


def assemble_blocks_to_host(V, K, red_op, products: ShardedProducts, host_pi, host_c, host_mp=None, cells_K=None,
                            pi_pieces: int = 2, c_pieces: int = 4, copy_stream=None, marks=None, tol: float = 1.0e-8) -> dict:
    """One block assembly of the coupled 3D-1D problem delivered into caller-owned host CSR arrays (ideally pinned):
    Π for `cells_K` (this rank's block of Γ cells; None: all), this rank's rows of C = Πᵀ(M_ΓΠ) and of MΠ.

    What bounds the end-to-end time is the device→host link (config 5: 3.2 GB at ~55 GB/s ≈ 60 ms against 35 ms of
    kernels), so the copy engine is started as early as possible and never left idle: Π is built in `pi_pieces`
    contiguous blocks of Γ cells (``create_interpolation_matrix(..., cells_K=block)``) and every block is downloaded
    while the next one is built; the blocks are concatenated on the device for the products, which
    ``ShardedProducts.to_host`` delivers piece by piece.  Blocks need K's cell-local dof numbering (Quadrature / DG
    spaces: the rows of a block of cells are a block of rows) and a contiguous ascending `cells_K`; otherwise Π is
    built in one piece.  The matrices are bit-identical to the one-shot calls.

    `host_pi` / `host_c` / `host_mp`: ``(indptr int64, indices int32, values float64)`` with room for the rows and
    entries (``ValueError`` otherwise).  Returns ``dict(nnz_pi, nnz_c, nnz_mp, rows_pi, rows_c, keep, copy_stream)``:
    synchronise `copy_stream` before reading the arrays and keep the dict alive until then."""
    import torch

    from .interpolation import _uses_stock_quadrature, create_interpolation_matrix, device_rows_block, device_space
    from .sparse import DeviceCSR

    if copy_stream is None:
        copy_stream = torch.cuda.Stream()
    k_list = np.asarray(K.dofmap.list)
    nip = k_list.shape[1]
    n_cells = k_list.shape[0]
    cells = np.arange(n_cells, dtype=np.int32) if cells_K is None else np.ascontiguousarray(cells_K, dtype=np.int32)
    contiguous = cells.size > 0 and int(cells[-1]) - int(cells[0]) + 1 == cells.size and bool(np.all(np.diff(cells) == 1))
    identity = hasattr(K, "__dict__") and K.__dict__.get("_fxii_identity_dofs")
    if identity is None:
        flat = k_list.reshape(-1)
        identity = bool(np.array_equal(flat, np.arange(flat.shape[0], dtype=flat.dtype)))
    # the block path (local row numbering, no explicit row-dof list, no full-height row pointers) also serves a single
    # block of cells — a rank's shard; only a K / operator that cannot be built block by block takes the general call
    pi_pieces = max(1, min(int(pi_pieces), int(cells.size)))
    blocks, first_rows = [cells], None
    use_blocks = bool(contiguous and identity and _uses_stock_quadrature(red_op))
    if use_blocks:
        blocks = np.array_split(cells, pi_pieces)
        first_rows = device_rows_block(K, red_op, blocks[0], known_contiguous=True)
        use_blocks = first_rows is not None
    if not use_blocks:
        pi_pieces, blocks = 1, [cells]
    keep = []

    def mark3(name, done, b0, b1):
        if marks is not None:
            marks.extend([(name + " ready (main stream)", done), (name + " D2H begin", b0), (name + " D2H end", b1)])

    def ev():
        return torch.cuda.Event(enable_timing=marks is not None)

    if not use_blocks:
        A, _, _ = create_interpolation_matrix(V, K, red_op, tol=tol, cells_K=cells_K)
        pi_local = A.device
        if cells_K is not None and contiguous and identity:
            pi_local = pi_local.row_slice(int(cells[0]) * nip, (int(cells[-1]) + 1) * nip)
        n_rows, nnz = pi_local.shape[0], pi_local.nnz
        if host_pi[0].shape[0] < n_rows + 1 or host_pi[1].shape[0] < nnz or host_pi[2].shape[0] < nnz:
            raise ValueError(f"host arrays for Π too small: {n_rows} rows, {nnz} entries")
        done, b0, b1 = ev(), ev(), ev()
        done.record()
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done)
            b0.record()
            pi_local.to_host_into(host_pi[0][: n_rows + 1], host_pi[1][:nnz], host_pi[2][:nnz], sync=False)
            b1.record()
        mark3("Pi", done, b0, b1)
        keep += [A, pi_local]
    else:
        dspace = device_space(V, tol)
        n_rows = cells.size * nip
        if host_pi[0].shape[0] < n_rows + 1:
            raise ValueError(f"host row pointers for Π too small: {n_rows} rows")
        host_pi[0][0] = 0
        h_ptr, h_idx, h_val = (torch.from_numpy(a) for a in host_pi)
        hints = red_op.__dict__.setdefault("_fxii_nnz_hints", {}) if hasattr(red_op, "__dict__") else {}
        subs, off, row0 = [], 0, 0
        for i, blk in enumerate(blocks):
            # local numbering: the block's Π has blk.size * nip rows
            t_h0 = time.perf_counter()
            rows = first_rows if i == 0 else device_rows_block(K, red_op, blk, known_contiguous=True)
            t_h1 = time.perf_counter()
            hkey = ("block", id(K), id(V), blk.size, int(blk[0]))
            if hkey in hints:
                rows.set_nnz_hint(hints[hkey])
            sub, st_blk = rows.assemble(dspace)
            if marks is not None:
                marks.append((f"host: Pi[{i}] rows handle [host ms]", t_h1 - t_h0))
                marks.append((f"host: Pi[{i}] assemble call [host ms]", time.perf_counter() - t_h1))
                marks.append((f"device: Pi[{i}] frames+rows+csr kernels [host ms]", 1e-3 * (st_blk["ms_frames"] + st_blk["ms_locate"] + st_blk["ms_csr"])))
            hints[hkey] = n = sub.nnz
            m = sub.shape[0]
            if h_idx.numel() < off + n or h_val.numel() < off + n:
                raise ValueError(f"host arrays for Π too small: at least {off + n} entries")
            ptr, idx, val = sub.device_tensors()
            if off:
                ptr[1:].add_(off)  # row ends in the numbering of the concatenated matrix (vstack copies them as they are)
            done, b0, b1 = ev(), ev(), ev()
            done.record()
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done)
                b0.record()
                h_ptr[row0 + 1:row0 + m + 1].copy_(ptr[1:], non_blocking=True)
                h_idx[off:off + n].copy_(idx, non_blocking=True)
                h_val[off:off + n].copy_(val, non_blocking=True)
                b1.record()
            mark3(f"Pi[{i}]", done, b0, b1)
            keep += [rows, sub]
            subs.append(sub)
            off += n
            row0 += m
        # (pointers are already shifted; a single block is used as it is)
        pi_local = subs[0] if len(subs) == 1 else DeviceCSR.vstack(subs, entry_offsets=[0] * len(subs))
        nnz = off
    out = products.to_host(pi_local, host_c, host_mp, n_sub=c_pieces, copy_stream=copy_stream, marks=marks)
    out["keep"] = keep + [pi_local] + out["keep"]
    out.update(nnz_pi=int(nnz), rows_pi=int(n_rows))
    return out


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
