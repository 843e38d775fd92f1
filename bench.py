#!/usr/bin/env python
"""bench.py — Π assembly throughput (nnz/s) and ΠᵀM_ΓΠ / MΠ time on B200, beside the host-CPU oracle.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C] [--impl ours|reference]

Default workload: BASELINE.json configs[4] ("cfg5": full coupled 3D-1D block assembly, Circle P1 on the
256^3 mesh, 1M-segment tree) — the largest configuration, and it fits one GPU.  One "step" of cfg5 = the
whole block assembly: Π build (row frames -> locate + tabulate + row reduction -> CSR), then C = Πᵀ(M_ΓΠ)
(all-gather of Π's row shards at N > 1, windowed transpose, SpGEMM with the mass diagonal fused) and MΠ.
Configs 1-4 name a Π build only; their step is the Π build.  `--gpus N` is STRONG scaling: the same Γ is
cut into N contiguous blocks of cells, Ω is replicated, the products are sharded by 3D dof.
Prints ONE JSON line (rank 0), last.  See DESIGN.md §6 for every key.  Nothing here reads /root/reference.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "pi_assembly_nnz_per_s"
UNIT = "nnz/s"

# BASELINE.json configs.  `scale` shrinks the mesh edge count and Γ for quick runs.
CONFIGS = {
    1: dict(name="cfg1 demos/coupled_poisson.py: P1 32x32x64, 63-seg line, Circle R=0.2 deg 5", n=(32, 32, 64), p=1,
            gamma="line", nseg=63, op="circle", qdeg=5, radius=0.2, products=False),
    2: dict(name="cfg2 PointwiseTrace, P1, 64^3 tets (1.57M), 10k-segment line network", n=(64, 64, 64), p=1,
            gamma="network", nseg=10_000, op="trace", products=False),
    3: dict(name="cfg3 Circle 64-pt, P2, 128^3 tets (12.6M), 100k-segment vessel tree", n=(128, 128, 128), p=2,
            gamma="tree", nseg=100_000, op="circle", qdeg=126, products=False),
    4: dict(name="cfg4 Disk 49-pt, P1, 256^3 tets (100.7M), 1M-segment vessel tree", n=(256, 256, 256), p=1,
            gamma="tree", nseg=1_000_000, op="disk", qdeg=7, products=False),
    5: dict(name="cfg5 full coupled 3D-1D block assembly (Pi, Pi^T M_Gamma Pi, M Pi): Circle 8-pt, P1, 256^3 tets (100.7M), "
                 "1M-segment vessel tree", n=(256, 256, 256), p=1, gamma="tree", nseg=1_000_000, op="circle", qdeg=15,
            products=True),
}
K_QUAD_DEGREE = 10  # quadrature element on Γ (demos/coupled_poisson.py:68-72): 6 Gauss points / segment
NVLINK_GBS = 770.0  # measured peer-copy bandwidth per direction (B200_PROFILING.md)


def build_workload(cfg: dict, pinned: bool):
    """Synthetic Ω, V, Γ, K, Q and the operator — the same problem on every rank (strong scaling)."""
    from fenicsx_ii_b200 import Circle, Disk, PointwiseTrace
    from fenicsx_ii_b200 import synthetic as sy
    from fenicsx_ii_b200.mesh import Mesh, functionspace

    nx, ny, nz = cfg["n"]
    mesh = sy.kuhn_box(nx, ny, nz)
    V = functionspace(mesh, ("Lagrange", 1)) if cfg["p"] == 1 else sy.kuhn_p2_space(mesh)
    h = 1.0 / nx
    seg_r = None
    if cfg["gamma"] == "line":
        gamma = sy.straight_line(cfg["nseg"] + 1)
    elif cfg["gamma"] == "network":
        gamma = sy.polyline_network(cfg["nseg"], h, seed=sy.SEED)
    else:
        gamma, seg_r = sy.vessel_tree(cfg["nseg"], h, seed=sy.SEED)
    if pinned:
        import torch

        def pin(a):
            return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()

        gamma = Mesh(pin(gamma.geometry.x), pin(gamma.geometry.dofmap), tdim=1, name=gamma.name)
        if seg_r is not None:
            seg_r = pin(seg_r)
    K = functionspace(gamma, ("Quadrature", K_QUAD_DEGREE))
    Q = functionspace(gamma, ("Lagrange", 1))
    if cfg["op"] == "trace":
        op = PointwiseTrace(gamma)
    elif cfg["op"] == "circle":
        op = Circle(gamma, seg_r if seg_r is not None else cfg["radius"], degree=cfg["qdeg"])
    else:
        op = Disk(gamma, seg_r if seg_r is not None else cfg["radius"], degree=cfg["qdeg"])
    return dict(mesh=mesh, V=V, gamma=gamma, K=K, Q=Q, op=op, seg_r=seg_r)


def scaled(cfg: dict, scale: float) -> dict:
    if scale == 1.0:
        return cfg
    c = dict(cfg)
    c["n"] = tuple(max(4, int(round(v * scale))) for v in cfg["n"])
    c["nseg"] = max(8, int(round(cfg["nseg"] * scale**3))) if cfg["gamma"] != "line" else cfg["nseg"]
    c["name"] = cfg["name"] + f" [scaled x{scale}]"
    return c


def config_dict(cfg: dict, wl: dict) -> dict:
    """The `config` object of the JSON line — built the same way by both arms."""
    K, V, op = wl["K"], wl["V"], wl["op"]
    nq = op.num_points
    nd = V.dofmap.list.shape[1]
    return {
        "workload": cfg["name"],
        "step": "Pi build + Pi^T (M_Gamma Pi) + M Pi" if cfg["products"] else "Pi build",
        "omega_tets": int(wl["mesh"].geometry.dofmap.shape[0]), "v_dofs": int(V.num_dofs),
        "gamma_segments": int(wl["gamma"].geometry.dofmap.shape[0]), "rows": int(K.num_dofs),
        "points": int(K.num_dofs * nq), "coo_entries": int(K.num_dofs * nq * nd),
        "l2": "256 MiB flush between timed steps (GPU arm)",
    }


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int, period: float = 0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def stop(self) -> dict:
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "nvml unavailable"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def measured_peak_gbs() -> tuple[float, str]:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


# algorithmic bytes per unit
def locate_bytes_per_point(nd: int) -> int:
    return 24 + 16 + 96 + 4 * nd + 12 * nd  # point + cell nodes + 4 vertices + dofmap row ; COO (col i32 + val f64)


def run_ours(args):
    import torch
    import torch.distributed as dist

    from fenicsx_ii_b200 import _lib, create_interpolation_matrix
    from fenicsx_ii_b200.distributed import ShardedProducts, p1_quadrature_coupling, quadrature_mass_diagonal, shard_range
    from fenicsx_ii_b200.interpolation import DeviceRows, device_space
    from fenicsx_ii_b200.sparse import DeviceCSR

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))  # NCCL_DEBUG is left as the caller set it
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    cfg = scaled(CONFIGS[args.config], args.scale)
    wl = build_workload(cfg, pinned=True)
    mesh, V, gamma, K, Q, op, seg_r = (wl[k] for k in ("mesh", "V", "gamma", "K", "Q", "op", "seg_r"))
    lib = _lib.load()
    products_on = bool(cfg["products"])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- this rank's block of Γ cells (strong scaling) ---------------------------------------------------
    n_cells = int(gamma.geometry.dofmap.shape[0])
    nip = int(K.element.interpolation_points.shape[0])
    c0, c1 = shard_range(n_cells, world, rank)
    my_cells = np.arange(c0, c1, dtype=np.int32)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device="cuda")  # 256 MiB > 126 MB L2

    # ---- cold e2e: the FIRST public call on a fresh mesh (Ω upload + LBVH build + Γ upload + Π build + download) -----
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dspace = device_space(V)
    torch.cuda.synchronize()
    mesh_build_s = time.perf_counter() - t0
    A_cold, _, _ = create_interpolation_matrix(V, K, op, cells_K=my_cells)
    _ = A_cold.indptr  # host CSR, as the reference's return value
    cold_s = time.perf_counter() - t0
    nnz_cold = int(A_cold.device.nnz)
    del A_cold
    mesh_info = dspace.mesh.info()

    # ---- device-resident inputs of the timed step --------------------------------------------------------
    desc = op.device_descriptor(None)
    rows = DeviceRows.from_gamma(np.asarray(gamma.geometry.x), np.asarray(gamma.geometry.dofmap), None if world == 1 else my_cells,
                                 np.asarray(K.element.interpolation_points), None, (c1 - c0) * nip, desc)
    prod = None
    m_dev = None
    if products_on:
        diag_dev = torch.from_numpy(quadrature_mass_diagonal(gamma, K)).to("cuda")
        m_dev = DeviceCSR.from_scipy(p1_quadrature_coupling(gamma, Q, K))
        prod = ShardedProducts(V.num_dofs, diag_dev, m_dev, world, rank)

    def step(events=None):
        csr, st = rows.assemble(dspace)
        out = prod(csr, events=events) if prod is not None else None
        return csr, st, out

    # W untimed warm-up steps (short steps get extra ones: module load, memory-pool growth, plans)
    for i in range(max(args.warmup, 3)):
        flush.zero_()
        csr, st, out = step()
        torch.cuda.synchronize()
    barrier()
    # one NVML poller per box, at 50 ms: eight pollers at 20 ms contended for the driver and cost every rank
    # milliseconds per step on an 8-GPU run; it keeps sampling through the end-to-end loop below
    sampler = ClockSampler(local_rank, period=float(os.environ.get("BENCH_CLOCK_PERIOD", "0.05")))
    if rank == 0 or os.environ.get("BENCH_SAMPLE_ALL_RANKS"):  # one NVML poller per box: eight of them contend for the driver
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    pi_keys = ("ms_frames", "ms_locate", "ms_csr", "ms_csr_count", "ms_csr_fill")
    prod_keys = ("allgather_ms", "transpose_ms", "spgemm_ms", "spgemm_mp_ms")
    stage = {k: 0.0 for k in pi_keys + prod_keys}
    stage["pi_ms"] = 0.0
    launches = 0
    from fenicsx_ii_b200 import distributed as fdist

    gather_parts = np.zeros(3)
    for a, b in ev:
        flush.zero_()  # L2 flush between timed iterations, outside the event pair
        if world > 1:
            dist.barrier()
            fdist.GATHER_TRACE = []
        l0 = lib.fxii_launch_count()
        pe = []
        a.record()
        csr, st, out = step(events=pe)
        b.record()
        launches += lib.fxii_launch_count() - l0
        torch.cuda.synchronize()
        if world > 1 and fdist.GATHER_TRACE and len(fdist.GATHER_TRACE) >= 4:
            g = fdist.GATHER_TRACE
            gather_parts += np.array([g[0].elapsed_time(g[1]), g[1].elapsed_time(g[2]), g[2].elapsed_time(g[3])])
        fdist.GATHER_TRACE = None
        for k in pi_keys:
            stage[k] += st[k]
        if pe:
            stage["pi_ms"] += a.elapsed_time(pe[0])
            for i, k in enumerate(prod_keys):
                stage[k] += pe[i].elapsed_time(pe[i + 1])
        else:
            stage["pi_ms"] += a.elapsed_time(b)
    barrier()
    total_ms = float(sum(a.elapsed_time(b) for a, b in ev))
    nnz_local = int(csr.nnz)
    Np, E = int(st["n_points"]), int(st["n_entries"])
    c_nnz = mp_nnz = ip = nnz_a = ip_mp = 0
    if out is not None:
        c_mat, mp_mat, ip, nnz_a, ip_mp = out
        c_nnz, mp_nnz = int(c_mat.nnz), int(mp_mat.nnz) if mp_mat is not None else 0
    # per-rank view of the stages that wait on other ranks (diagnostics of the scaling curve)
    per_rank = None
    if world > 1:
        mine = torch.tensor([stage["pi_ms"] / args.steps, stage["allgather_ms"] / args.steps, stage["spgemm_ms"] / args.steps,
                             stage["transpose_ms"] / args.steps, float(nnz_local), float(c_nnz), float(ip)] + list(gather_parts / args.steps),
                            device="cuda", dtype=torch.float64)
        allr = torch.empty(world * mine.numel(), device="cuda", dtype=torch.float64)
        dist.all_gather_into_tensor(allr, mine)
        allr = allr.view(world, -1).cpu().numpy()
        names = ("pi_ms", "allgather_ms", "spgemm_ms", "transpose_ms", "nnz_pi", "nnz_c", "ip_c", "gather_pack_or_handshake_ms",
                 "gather_nccl_or_copies_ms", "gather_unpack_or_fence_ms")
        per_rank = {n: [round(float(v), 3) for v in allr[:, i]] for i, n in enumerate(names)}
    # max over ranks of the times; sum over ranks of the work
    keys = sorted(stage)
    t_all = torch.tensor([total_ms] + [stage[k] for k in keys], device="cuda", dtype=torch.float64)
    w_all = torch.tensor([float(nnz_local), float(Np), float(c_nnz), float(mp_nnz), float(ip), float(nnz_a), float(E)],
                         device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
        dist.all_reduce(w_all, op=dist.ReduceOp.SUM)
    ms_per_step = float(t_all[0].item()) / args.steps
    stage_ms = {k: float(v) / args.steps for k, v in zip(keys, t_all[1:].tolist())}
    nnz_total, np_total, c_total, mp_total, ip_total, nnz_a_total, e_total = (float(v) for v in w_all.tolist())
    value = nnz_total / (ms_per_step * 1e-3)

    # ---- e2e: the public API with HOST (pinned) Γ arrays in and HOST (pinned) CSR blocks out, every step -----------
    def pinned(n, dtype):
        return torch.empty(max(int(n), 1), dtype=dtype).pin_memory().numpy()

    host = {"pi": (pinned((c1 - c0) * nip + 1, torch.int64), pinned(nnz_local, torch.int32), pinned(nnz_local, torch.float64))}
    if out is not None:
        host["c"] = (pinned(c_mat.shape[0] + 1, torch.int64), pinned(c_nnz, torch.int32), pinned(c_nnz, torch.float64))
        if mp_mat is not None:
            host["mp"] = (pinned(mp_mat.shape[0] + 1, torch.int64), pinned(mp_nnz, torch.int32), pinned(mp_nnz, torch.float64))
    del out, csr
    r0, r1 = c0 * nip, c1 * nip

    copy_stream = torch.cuda.Stream()

    from fenicsx_ii_b200.distributed import assemble_blocks_to_host

    def e2e_step(marks=None):
        """One end-to-end block assembly through the public API: pinned host Γ arrays in, pinned host CSR blocks out.
        Π is built in `--e2e-pi-pieces` blocks of Γ cells and C in `--e2e-pieces` row pieces; every finished block is
        downloaded on the copy stream while the next one is computed (the D2H link bounds the step)."""
        if marks is not None:
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record()
            marks.append(("start", e0))
        if prod is not None:
            keep = assemble_blocks_to_host(V, K, op, prod, host["pi"], host["c"], host.get("mp"), cells_K=None if world == 1 else my_cells,
                                           pi_pieces=args.e2e_pi_pieces, c_pieces=args.e2e_pieces, copy_stream=copy_stream, marks=marks)
        elif world == 1:
            # Π only: host arrays in, host CSR out in ONE C call (the download rides behind the kernel, one synchronisation)
            keep = create_interpolation_matrix(V, K, op, host_out=host["pi"])
        else:
            A, _, _ = create_interpolation_matrix(V, K, op, cells_K=my_cells)
            pi_local = A.device.row_slice(r0, r1)
            done = torch.cuda.Event()
            done.record()
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done)
                pi_local.to_host_into(*host["pi"], sync=False)
            keep = [A, pi_local]
        copy_stream.synchronize()  # all blocks are in the host arrays
        torch.cuda.current_stream().wait_stream(copy_stream)
        del keep

    e2e_steps = max(2, min(args.steps, 5))
    e2e_step()
    barrier()
    e_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(e2e_steps)]
    for a, b in e_ev:
        flush.zero_()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a.record()
        e2e_step()
        b.record()
    barrier()
    clocks = sampler.stop()
    # where one end-to-end step spends its time: device timestamps (ms after the start event) of the main-stream and
    # copy-stream milestones of ONE extra, untimed step
    marks = []
    flush.zero_()
    torch.cuda.synchronize()
    e2e_step(marks)
    torch.cuda.synchronize()
    e_start = marks[0][1]
    e2e_timeline = {}
    for name, m in marks[1:]:
        e2e_timeline[name] = round(1e3 * m, 3) if isinstance(m, float) else round(e_start.elapsed_time(m), 3)
    e2e_ms = float(sum(a.elapsed_time(b) for a, b in e_ev)) / e2e_steps
    e_all = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e_all, op=dist.ReduceOp.MAX)
    e2e_ms = float(e_all.item())
    gx, gc = gamma.geometry.x, gamma.geometry.dofmap
    # every create_interpolation_matrix call (one per block of Γ cells) uploads Γ's geometry and the operator tables;
    # cells_K and the per-cell radii of the blocks add up to this rank's cells
    n_calls = args.e2e_pi_pieces if products_on else 1
    h2d = n_calls * (gx.nbytes + gc.nbytes + 8 * nip) + (0 if (world == 1 and n_calls == 1) else 4 * (c1 - c0))
    if cfg["op"] != "trace":
        h2d += n_calls * 32 * op.num_points + (8 * (c1 - c0) if seg_r is not None else 8 * n_calls)
    d2h = sum(a.nbytes + 4 * b.shape[0] + 8 * c.shape[0] for a, b, c in host.values())

    # ---- roofline: every stage of the step with its algorithmic bytes, the dominant one named ---------
    peak, peak_src = measured_peak_gbs()
    nd = dspace.nd
    n_rows_local = (c1 - c0) * nip
    fused = bool(st.get("fused", 0))
    k_loc = stage_ms["ms_locate"]
    k_csr = stage_ms["ms_csr"]
    stages = []

    def add(name, kernel, alg_bytes, ms, note=None, peak_gbs=peak):
        gbs = alg_bytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        s = {"stage": name, "kernel": kernel, "algorithmic_bytes": int(alg_bytes), "ms": ms, "GB/s": gbs, "frac": gbs / peak_gbs}
        if note:
            s["note"] = note
        stages.append(s)

    # per-rank figures (rank 0's shard; shards are balanced)
    add("locate+tabulate(+row reduction)", "pi_fused_kernel" if fused else "pi_locate_tabulate_kernel",
        locate_bytes_per_point(nd) * Np, k_loc, "216 B (P1) / 336 B (P2) per point incl. the COO the fused kernel no longer writes")
    if fused:
        add("COO->CSR (staged rows -> CSR)", "cub scan + compact_rows_kernel", 24 * nnz_local + 12 * n_rows_local, k_csr,
            "sort + segmented reduce run inside the row kernel; this is the remaining scan + compaction, bytes actually moved")
    else:
        add("COO->CSR", "rows_count_kernel + rows_fill_kernel", 16 * E + 12 * nnz_local + 8 * (n_rows_local + 1), k_csr)
    if products_on:
        g_nnz = nnz_total  # the gathered Π every rank transposes a window of
        if world > 1:
            add("all-gather-v of Pi shards", "NCCL", 12 * (g_nnz - nnz_local), stage_ms["allgather_ms"],
                "bytes received per rank; peak = measured NVLink peer copy 770 GB/s per direction", peak_gbs=NVLINK_GBS)
        m_win = V.num_dofs / world
        add("transpose (window)", "cub select + radix sort + gather_transposed_kernel",
            (4 * g_nnz if world > 1 else 0) + 24 * nnz_a + 8 * (m_win + n_rows_local * world + 2), stage_ms["transpose_ms"],
            "24 B per transposed entry + row pointers" + (" + one pass over the gathered column indices (window select)" if world > 1 else ""))
        add("SpGEMM C = Pi^T (M_Gamma Pi)", "spgemm_warp_kernel (symbolic + numeric)",
            12 * nnz_a + 12 * ip + 12 * c_nnz + 16 * (m_win + 1), stage_ms["spgemm_ms"])
        add("SpGEMM M Pi", "spgemm_warp_kernel (symbolic + numeric)",
            12 * (m_dev.nnz / world) + 12 * ip_mp + 12 * mp_nnz + 16 * (m_dev.shape[0] / world + 1), stage_ms["spgemm_mp_ms"])
    hbm_stages = [s for s in stages if s["kernel"] != "NCCL"]
    dom = max(hbm_stages, key=lambda s: s["ms"])
    roofline = {"kernel": dom["kernel"], "stage": dom["stage"], "bound": "hbm", "achieved": dom["GB/s"], "peak": peak, "unit": "GB/s",
                "frac": dom["frac"], "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": dom["algorithmic_bytes"],
                "kernel_ms": dom["ms"], "stages": stages,
                "note": "per-rank stage times are CUDA-event means over the timed steps (max over ranks); traffic: see profiles/ (ncu)"}

    out_json = None
    if rank == 0:
        pi_ms = stage_ms["pi_ms"]
        out_json = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": config_dict(cfg, wl),
            "run": {"sharding": f"gamma cells in {world} contiguous blocks, mesh replicated, products sharded by 3D dof",
                    "rows_per_gpu": int(n_rows_local), "nnz_pi_global": int(nnz_total), "nnz_c_global": int(c_total),
                    "nnz_mp_global": int(mp_total), "ip_c_global": int(ip_total)},
            "stage_ms": stage_ms,
            "per_rank": per_rank,
            "pi": {"ms": pi_ms, "nnz_per_s": nnz_total / (pi_ms * 1e-3) if pi_ms > 0 else None,
                   "points_per_s": np_total / (pi_ms * 1e-3) if pi_ms > 0 else None},
            "products": ({"pit_m_pi_s": (stage_ms["allgather_ms"] + stage_ms["transpose_ms"] + stage_ms["spgemm_ms"]) * 1e-3,
                          "spgemm_s": stage_ms["spgemm_ms"] * 1e-3, "m_pi_s": stage_ms["spgemm_mp_ms"] * 1e-3,
                          "plan_setup_ms": prod.plan.setup_ms if prod.plan else None} if products_on else None),
            "e2e": {"value": nnz_total / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms, "timeline_ms": e2e_timeline,
                    "note": ("assemble_blocks_to_host: Π built from pinned host Γ arrays in "
                             f"{args.e2e_pi_pieces} block(s) of Γ cells + ShardedProducts.to_host (MΠ first, C in {args.e2e_pieces} row piece(s)); "
                             "Π, C and MΠ land in pinned host CSR arrays every step, every block downloaded while the next is "
                             "computed; Ω handle (upload + LBVH) cached") if products_on else
                            "create_interpolation_matrix(V, K, op, host_out=pinned CSR arrays) from pinned host Γ arrays, every step; "
                            "Ω handle (upload + LBVH) cached"},
            "e2e_cold": {"value": nnz_cold / cold_s if world == 1 else None, "unit": UNIT, "s": cold_s, "mesh_build_s": mesh_build_s,
                         "note": "FIRST call on a fresh mesh: Ω upload + LBVH build + Γ upload + Π build + CSR download (Π only); "
                                 "the reference rebuilds its tree on every call"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "mesh": {"build_s": mesh_build_s, "device_bytes": int(mesh_info["device_bytes"]),
                     "lbvh_depth": int(mesh_info["max_depth"]), "cells_per_s": mesh_info["n_cells"] / mesh_build_s},
        }
        if args.cpu_baseline and world == 1:
            out_json["cpu_baseline"] = cpu_baseline(cfg, args.cpu_seconds, threads=os.cpu_count() or 1, wl=wl)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        sys.stdout.flush()
        print(json.dumps(out_json), flush=True)


# ---- CPU side: the oracle port timed on the host cores ------------------------------------------
_CPU = {}  # workload arrays + locator, inherited copy-on-write by forked workers


def _cpu_points(wl, cfg, cells):
    """The 3D averaging points of a block of Γ cells (oracle restatement)."""
    from oracle import pi_oracle as orc

    gamma, K = wl["gamma"], wl["K"]
    xref = np.asarray(K.element.interpolation_points, dtype=np.float64).reshape(-1)
    radius = cfg.get("radius")
    if wl["seg_r"] is not None:
        per_row = np.repeat(np.asarray(wl["seg_r"])[cells], xref.shape[0])
        radius = lambda x: per_row  # noqa: E731
    table = w = None
    if cfg["op"] == "circle":
        table, w = orc.circle_table(cfg["qdeg"])
    elif cfg["op"] == "disk":
        table, w = orc.disk_table(cfg["qdeg"])
    pts, _, _ = orc.compute_quadrature(cfg["op"], gamma.geometry.x, gamma.geometry.dofmap, cells, xref, radius, table, w)
    return np.asarray(pts).reshape(-1, 3)


def _cpu_prepare(cfg, wl, n_sample):
    """Locator over the neighbourhood of the sample only (the full-mesh grid of 100 M tets takes 140 s to build)."""
    from oracle import pi_oracle as orc

    mesh, V = wl["mesh"], wl["V"]
    cells = np.arange(n_sample, dtype=np.int32)
    t0 = time.perf_counter()
    pts = _cpu_points(wl, cfg, cells)
    if "h_max" not in _CPU:
        _CPU["h_max"] = orc.cell_extent_max(mesh.geometry.x, mesh.geometry.dofmap)
    sel = orc.cells_near_points(mesh.geometry.x, mesh.geometry.dofmap, pts, reach=1e-8, h_max=_CPU["h_max"])
    sub_cells = np.ascontiguousarray(mesh.geometry.dofmap[sel])
    locator = orc.GridLocator(mesh.geometry.x, sub_cells)
    _CPU.update(cfg=cfg, wl=wl, locator=locator, sub_cells=sub_cells, sub_dofmap=np.ascontiguousarray(V.dofmap.list[sel]),
                locator_build_s=time.perf_counter() - t0, sub_n=int(sel.shape[0]), n_sample=n_sample)


def _cpu_chunk(cells):
    """Oracle Π (and, for cfg5, the block products of that Π) for a block of Γ cells; returns (nnz, points, seconds)."""
    from oracle import pi_oracle as orc

    c = _CPU
    cfg, wl = c["cfg"], c["wl"]
    cells_k = np.asarray(cells, dtype=np.int32)
    mesh, V, gamma, K = wl["mesh"], wl["V"], wl["gamma"], wl["K"]
    nip = K.element.interpolation_points.shape[0]
    radius = cfg.get("radius")
    if wl["seg_r"] is not None:
        per_row = np.repeat(np.asarray(wl["seg_r"])[cells_k], nip)
        radius = lambda x: per_row  # noqa: E731
    t0 = time.perf_counter()
    r = orc.assemble_pi(mesh.geometry.x, c["sub_cells"], c["sub_dofmap"], V.element.degree, gamma.geometry.x,
                        gamma.geometry.dofmap, K.dofmap.list, K.element.interpolation_points, cfg["op"], radius=radius,
                        quad_degree=cfg.get("qdeg"), cells_K=cells_k, n_rows=K.num_dofs, locator=c["locator"])
    if cfg["products"]:
        # the same blocks on this block of Γ rows: Πᵀ (M_Γ Π) and M Π, PETSc semantics (oracle restatement)
        lo, hi = int(cells_k[0]) * nip, (int(cells_k[-1]) + 1) * nip
        ip_, ij_, iv_ = r["indptr"][lo:hi + 1] - r["indptr"][lo], r["indices"], r["data"]
        diag = c["diag"][lo:hi]
        mp_ = (ip_, ij_, iv_ * np.repeat(diag, np.diff(ip_)))
        ti, tj, tv = orc.csr_transpose(ip_, ij_, iv_, V.num_dofs)
        # only the touched rows of Πᵀ matter: compress the empty ones away before the product
        cnt = np.diff(ti)
        tsub = np.concatenate([[0], np.cumsum(cnt[cnt > 0])]).astype(np.int64)
        orc.spgemm(tsub, tj, tv, *mp_, V.num_dofs)
        M = c["M"][:, lo:hi].tocsr()
        M.sort_indices()
        Ms = M[np.nonzero(np.diff(M.indptr))[0]]
        orc.spgemm(Ms.indptr.astype(np.int64), Ms.indices.astype(np.int32), Ms.data, ip_, ij_, iv_, V.num_dofs)
    return int(r["indices"].shape[0]), int(r["cell"].shape[0]), time.perf_counter() - t0


def cpu_setup(cfg, seconds: float, threads: int, wl=None):
    """Sizes the bounded sample (the first cells of Γ, about `seconds` of work on `threads` workers, from a 48-cell probe)
    and builds the oracle's locator over the sample's neighbourhood."""
    from fenicsx_ii_b200.distributed import p1_quadrature_coupling, quadrature_mass_diagonal

    if wl is None:
        wl = build_workload(cfg, pinned=False)
    _CPU.clear()
    n_cells = int(wl["gamma"].geometry.dofmap.shape[0])
    probe_n = min(n_cells, 48)
    _cpu_prepare(cfg, wl, probe_n)
    if cfg["products"]:
        _CPU["diag"] = quadrature_mass_diagonal(wl["gamma"], wl["K"])
        _CPU["M"] = p1_quadrature_coupling(wl["gamma"], wl["Q"], wl["K"]).tocsc()
    _, _, dt = _cpu_chunk(np.arange(probe_n))
    rate = probe_n / max(dt, 1e-6)  # Γ cells per second on one thread
    n_sample = int(min(n_cells, max(probe_n, rate * seconds * threads)))
    if n_sample != probe_n:
        _cpu_prepare(cfg, wl, n_sample)
    _CPU["threads"] = threads
    _CPU["total_cells"] = n_cells


def cpu_step() -> dict:
    """One pass of the numpy oracle over the prepared sample with forked workers (the locator build — the analogue of
    dolfinx's tree build, here over the sample's neighbourhood only — is reported separately)."""
    n_sample, threads = _CPU["n_sample"], _CPU["threads"]
    chunks = [c for c in np.array_split(np.arange(n_sample, dtype=np.int32), threads) if c.size]
    t0 = time.perf_counter()
    if len(chunks) == 1:
        res = [_cpu_chunk(chunks[0])]
    else:
        import multiprocessing as mp

        with mp.get_context("fork").Pool(len(chunks)) as pool:
            res = pool.map(_cpu_chunk, chunks)
    wall = time.perf_counter() - t0
    return dict(nnz=sum(r[0] for r in res), points=sum(r[1] for r in res), wall_s=wall, locator_build_s=_CPU["locator_build_s"],
                sub_cells=_CPU["sub_n"], sample_cells=n_sample, total_cells=_CPU["total_cells"], workers=len(chunks))


def _cpu_sample_text(cfg, r) -> str:
    what = "Π + Πᵀ(M_ΓΠ) + MΠ" if cfg["products"] else "Π"
    return (f"first {r['sample_cells']} of {r['total_cells']} Γ cells of the same workload ({r['points']} points), {what}, in "
            f"{r['wall_s']:.1f} s on {r['workers']} forked workers; numpy oracle (CPU restatement, not the dolfinx reference); "
            f"locator over the {r['sub_cells']} tets near the sample built in {r['locator_build_s']:.1f} s, not included")


def cpu_baseline(cfg, seconds: float, threads: int, wl=None) -> dict:
    cpu_setup(cfg, seconds, threads, wl=wl)
    r = cpu_step()
    return {"value": r["nnz"] / r["wall_s"], "unit": UNIT, "cores": r["workers"], "kind": "port", "sample": _cpu_sample_text(cfg, r),
            "points_per_s": r["points"] / r["wall_s"]}


def run_reference(args):
    """--impl reference: the reference's algorithm on the host cores.  The dolfinx/PETSc stack is not
    installable in this image, so this is the oracle port with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = scaled(CONFIGS[args.config], args.scale)
    wl = build_workload(cfg, pinned=False)
    threads = os.cpu_count() or 1
    per_step = max(1.0, min(args.cpu_seconds, 90.0 / max(1, args.steps + args.warmup)))
    cpu_setup(cfg, per_step, threads, wl=wl)  # the sample and its locator are prepared once; every step runs the same sample
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        last = cpu_step()
        if i >= args.warmup:
            vals.append((last["nnz"], last["wall_s"]))
    nnz = sum(v[0] for v in vals)
    wall = sum(v[1] for v in vals)
    value = nnz / wall
    sample = "each step: " + _cpu_sample_text(cfg, last)
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, len(vals)), "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": config_dict(cfg, wl),
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["workers"], "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=5, choices=sorted(CONFIGS))
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (testing only)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--e2e-pieces", type=int, default=0, help="row pieces of C whose downloads overlap the remaining products "
                    "(e2e); 0: by GPU count (4 / 2 / 2 / 1 at 1 / 2 / 4 / 8 GPUs: the fewer bytes a rank downloads, the less there is to hide)")
    ap.add_argument("--e2e-pi-pieces", type=int, default=0, help="blocks of Γ cells Π is built in, each downloaded while the next is "
                    "built (e2e); 0: 2 on one GPU, else 1")
    args = ap.parse_args()
    if args.e2e_pieces <= 0:
        args.e2e_pieces = {1: 4, 2: 2, 4: 2}.get(args.gpus, 1)
    if args.e2e_pi_pieces <= 0:
        args.e2e_pi_pieces = 2 if args.gpus == 1 else 1
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
