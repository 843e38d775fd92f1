#!/usr/bin/env python
"""bench.py — Π assembly throughput (nnz/s) and ΠᵀM_ΓΠ time on B200, beside the host-CPU oracle.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C] [--impl ours|reference]

One "step" = one Π build (row frames -> locate+tabulate -> COO -> CSR) of the configuration's
synthetic workload, inputs resident in HBM.  Prints ONE JSON line (rank 0).  See DESIGN.md
§measurement for every key.  Nothing here reads /root/reference.
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
            gamma="line", nseg=63, op="circle", qdeg=5, radius=0.2),
    2: dict(name="cfg2 PointwiseTrace, P1, 64^3 tets (1.57M), 10k-segment line network", n=(64, 64, 64), p=1,
            gamma="network", nseg=10_000, op="trace"),
    3: dict(name="cfg3 Circle 64-pt, P2, 128^3 tets (12.6M), 100k-segment vessel tree", n=(128, 128, 128), p=2,
            gamma="tree", nseg=100_000, op="circle", qdeg=126),
    4: dict(name="cfg4 Disk 49-pt, P1, 256^3 tets (100.7M), 1M-segment vessel tree", n=(256, 256, 256), p=1,
            gamma="tree", nseg=1_000_000, op="disk", qdeg=7),
    5: dict(name="cfg5 Circle 8-pt, P1, 256^3 tets, 1M-segment tree, full block products", n=(256, 256, 256), p=1,
            gamma="tree", nseg=1_000_000, op="circle", qdeg=15),
}
K_QUAD_DEGREE = 10  # quadrature element on Γ (demos/coupled_poisson.py:68-72): 6 Gauss points / segment


def build_workload(cfg: dict, rank: int, pinned: bool):
    """Synthetic Ω, V, Γ, K and the operator.  Γ differs per rank (weak scaling: every GPU gets a
    full-size Γ shard); Ω is replicated."""
    from fenicsx_ii_b200 import Circle, Disk, PointwiseTrace
    from fenicsx_ii_b200 import synthetic as sy
    from fenicsx_ii_b200.mesh import Mesh, functionspace

    nx, ny, nz = cfg["n"]
    mesh = sy.kuhn_box(nx, ny, nz)
    V = functionspace(mesh, ("Lagrange", 1)) if cfg["p"] == 1 else sy.kuhn_p2_space(mesh)
    h = 1.0 / nx
    seed = sy.SEED + 7919 * rank
    seg_r = None
    if cfg["gamma"] == "line":
        gamma = sy.straight_line(cfg["nseg"] + 1)
    elif cfg["gamma"] == "network":
        gamma = sy.polyline_network(cfg["nseg"], h, seed=seed)
    else:
        gamma, seg_r = sy.vessel_tree(cfg["nseg"], h, seed=seed)
    if pinned:
        import torch

        def pin(a):
            t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
            return t.numpy()

        g2 = Mesh(pin(gamma.geometry.x), pin(gamma.geometry.dofmap), tdim=1, name=gamma.name)
        gamma = g2
    K = functionspace(gamma, ("Quadrature", K_QUAD_DEGREE))
    if pinned:
        K.dofmap.list = pin(K.dofmap.list)
        if seg_r is not None:
            seg_r = pin(seg_r)
    if cfg["op"] == "trace":
        op = PointwiseTrace(gamma)
    elif cfg["op"] == "circle":
        op = Circle(gamma, seg_r if seg_r is not None else cfg["radius"], degree=cfg["qdeg"])
    else:
        op = Disk(gamma, seg_r if seg_r is not None else cfg["radius"], degree=cfg["qdeg"])
    return mesh, V, gamma, K, op, seg_r


def scaled(cfg: dict, scale: float) -> dict:
    if scale == 1.0:
        return cfg
    c = dict(cfg)
    c["n"] = tuple(max(4, int(round(v * scale))) for v in cfg["n"])
    c["nseg"] = max(8, int(round(cfg["nseg"] * scale**3))) if cfg["gamma"] != "line" else cfg["nseg"]
    c["name"] = cfg["name"] + f" [scaled x{scale}]"
    return c


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


# algorithmic bytes
def locate_bytes_per_point(nd: int) -> int:
    return 24 + 16 + 96 + 4 * nd + 12 * nd  # point + cell nodes + 4 vertices + dofmap row ; COO (col i32 + val f64)


def csr_bytes(E: int, nnz: int, n_rows: int) -> int:
    return 12 * E + 12 * nnz + 8 * (n_rows + 1)


def spgemm_bytes(nnz_a: int, ip: int, nnz_c: int, m: int) -> int:
    return 12 * nnz_a + 12 * ip + 12 * nnz_c + 16 * (m + 1)


def run_ours(args):
    import torch
    import torch.distributed as dist

    from fenicsx_ii_b200 import _lib, create_interpolation_matrix
    from fenicsx_ii_b200.interpolation import device_rows, device_space

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        # keep stdout clean for the single JSON line (NCCL prints its version banner there)
        os.environ["NCCL_DEBUG"] = os.environ.get("BENCH_NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    cfg = scaled(CONFIGS[args.config], args.scale)
    mesh, V, gamma, K, op, seg_r = build_workload(cfg, rank, pinned=True)
    lib = _lib.load()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # -- cold part, once: Ω upload + LBVH build (the reference rebuilds its tree on every call)
    t0 = time.perf_counter()
    dspace = device_space(V)
    torch.cuda.synchronize()
    mesh_build_s = time.perf_counter() - t0
    mesh_info = dspace.mesh.info()

    rows = device_rows(K, op)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.int32, device="cuda")  # 256 MiB > 126 MB L2

    def step():
        return rows.assemble(dspace)

    # W untimed warm-up steps; short steps get extra ones so that first-use costs (module load,
    # memory-pool growth) never leak into the K timed steps
    warm = args.warmup
    for i in range(max(args.warmup, 10)):
        flush.zero_()
        t_w = time.perf_counter()
        csr, st = step()
        torch.cuda.synchronize()
        if i + 1 >= warm and time.perf_counter() - t_w > 5e-3:
            break
    barrier()
    sampler = ClockSampler(local_rank, period=float(os.environ.get("BENCH_CLOCK_PERIOD", "0.02")))
    sampler.start()
    launches0 = lib.fxii_launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    stage = {"ms_frames": 0.0, "ms_locate": 0.0, "ms_csr": 0.0, "ms_csr_count": 0.0, "ms_csr_fill": 0.0}
    launches = 0
    for a, b in ev:
        flush.zero_()  # L2 flush between timed iterations, outside the event pair
        l0 = lib.fxii_launch_count()
        a.record()
        csr, st = step()
        b.record()
        launches += lib.fxii_launch_count() - l0
        for k in stage:
            stage[k] += st[k]
    barrier()
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(step_ms))
    nnz_local = csr.nnz
    E = st["n_entries"]
    Np = st["n_points"]
    # max over ranks of the timed total; sum over ranks of the work
    t_all = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
    w_all = torch.tensor([float(nnz_local), float(Np)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
        dist.all_reduce(w_all, op=dist.ReduceOp.SUM)
    total_ms_max = float(t_all.item())
    nnz_total, np_total = float(w_all[0].item()), float(w_all[1].item())
    ms_per_step = total_ms_max / args.steps
    value = nnz_total / (ms_per_step * 1e-3)

    # -- e2e: the public API with HOST (pinned) inputs, H2D of Γ and D2H of the CSR inside the timed region
    indptr_h = torch.empty(K.num_dofs + 1, dtype=torch.int64).pin_memory().numpy()
    indices_h = torch.empty(max(nnz_local, 1), dtype=torch.int32).pin_memory().numpy()
    values_h = torch.empty(max(nnz_local, 1), dtype=torch.float64).pin_memory().numpy()

    def e2e_step():
        # the public call with caller-owned pinned result arrays: Γ upload + build + CSR download, one C call each
        A, _, _ = create_interpolation_matrix(V, K, op, host_out=(indptr_h, indices_h, values_h))
        return A

    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        e2e_step()
    barrier()
    e_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(e2e_steps)]
    for a, b in e_ev:
        flush.zero_()
        torch.cuda.synchronize()
        a.record()
        e2e_step()
        b.record()
    barrier()
    e2e_ms = float(sum(a.elapsed_time(b) for a, b in e_ev)) / e2e_steps
    e_all = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e_all, op=dist.ReduceOp.MAX)
    e2e_value = nnz_total / (float(e_all.item()) * 1e-3)
    gx, gc = gamma.geometry.x, gamma.geometry.dofmap
    # Γ geometry + reference points; cells_K and the row numbering are the identity here and are generated on the device
    h2d = gx.nbytes + gc.nbytes + 8 * K.element.interpolation_points.shape[0]
    if cfg["op"] != "trace":
        h2d += 32 * op.num_points + (8 * K.num_dofs if seg_r is not None else 8)
    d2h = 8 * (K.num_dofs + 1) + 12 * nnz_local

    # -- block product ΠᵀM_ΓΠ (matrix_assembler.py:225-258 with a diagonal quadrature mass)
    from fenicsx_ii_b200.distributed import product_pt_m_p

    try:
        prod = product_pt_m_p(csr, gamma, K, world, rank, steps=max(3, min(args.steps, 10)), flush=flush)
    except Exception as exc:  # e.g. a product too large for one GPU: report it, keep the Π numbers
        prod = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        if world > 1:
            raise

    # -- roofline of the dominant kernel of the step
    peak, peak_src = measured_peak_gbs()
    nd = dspace.nd
    k_loc = stage["ms_locate"] / args.steps
    if st.get("fused", 0) == 3:  # single cooperative launch: ms_locate + ms_csr apportion ONE kernel's event time
        k_loc += stage["ms_csr"] / args.steps
    k_cnt = stage["ms_csr_count"] / args.steps
    k_fill = stage["ms_csr_fill"] / args.steps
    fused = bool(st.get("fused", 0))
    cands = {
        ("pi_fused_kernel" if fused else "pi_locate_tabulate_kernel"): (locate_bytes_per_point(nd) * Np, k_loc),
        "rows_fill_kernel": (12 * E + 12 * nnz_local, k_fill),
        "rows_count_kernel": (4 * E + 8 * K.num_dofs, k_cnt),
    }
    dom = max(cands, key=lambda k: cands[k][1])
    dom_bytes, dom_ms = cands[dom]
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    traffic = None
    try:  # per-launch DRAM bytes of the dominant kernel from the committed ncu capture of this workload
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            traffic = json.load(f).get(f"cfg{args.config}", {}).get(dom) if args.scale == 1.0 else None
    except Exception:
        traffic = None
    roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes, "kernel_ms": dom_ms,
                "stages_ms": {k: v / args.steps for k, v in stage.items()},
                "all": {k: {"GB/s": (b / (ms * 1e-3) / 1e9 if ms > 0 else 0.0), "ms": ms} for k, (b, ms) in cands.items()}}

    out = None
    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["name"], "omega_tets": int(mesh_info["n_cells"]), "v_dofs": V.num_dofs,
                       "gamma_segments_per_gpu": int(gamma.geometry.dofmap.shape[0]), "rows_per_gpu": K.num_dofs,
                       "points_per_gpu": int(Np), "coo_entries_per_gpu": int(E), "nnz_per_gpu": int(nnz_local),
                       "sharding": f"gamma rows x{world}, mesh replicated", "l2": "256 MiB flush between timed steps"},
            "points_per_s": np_total / (ms_per_step * 1e-3),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": float(e_all.item()),
                    "note": "create_interpolation_matrix(V,K,op,host_out=pinned arrays): pinned host Γ arrays in, host CSR out; Ω handle cached",
                    "cold_mesh_build_s": mesh_build_s},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "products": prod,
            "mesh": {"build_s": mesh_build_s, "device_bytes": int(mesh_info["device_bytes"]),
                     "lbvh_depth": int(mesh_info["max_depth"]), "cells_per_s": mesh_info["n_cells"] / mesh_build_s},
        }
        if args.cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(cfg, args.cpu_seconds, threads=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


# ---- CPU side: the oracle port timed on the host cores ------------------------------------------
_CPU = {}  # workload arrays + locator, inherited copy-on-write by forked workers


def _cpu_prepare(cfg):
    from oracle import pi_oracle as orc

    if _CPU.get("cfg_name") == cfg["name"]:
        return
    mesh, V, gamma, K, op, seg_r = build_workload(cfg, 0, pinned=False)
    t0 = time.perf_counter()
    locator = orc.GridLocator(mesh.geometry.x, mesh.geometry.dofmap)
    _CPU.update(cfg_name=cfg["name"], mesh=mesh, V=V, gamma=gamma, K=K, seg_r=seg_r, locator=locator,
                locator_build_s=time.perf_counter() - t0, kind=cfg["op"], qdeg=cfg.get("qdeg"), radius=cfg.get("radius"))


def _cpu_chunk(cells):
    """Oracle Π for a block of Γ cells; returns (nnz, points, seconds)."""
    from oracle import pi_oracle as orc

    c = _CPU
    cells_k = np.asarray(cells, dtype=np.int32)
    mesh, V, gamma, K = c["mesh"], c["V"], c["gamma"], c["K"]
    radius = c["radius"]
    if c["seg_r"] is not None:
        per_row = np.repeat(c["seg_r"][cells_k], K.element.interpolation_points.shape[0])
        radius = lambda x: per_row  # noqa: E731
    t0 = time.perf_counter()
    r = orc.assemble_pi(mesh.geometry.x, mesh.geometry.dofmap, V.dofmap.list, V.element.degree, gamma.geometry.x,
                        gamma.geometry.dofmap, K.dofmap.list, K.element.interpolation_points, c["kind"], radius=radius,
                        quad_degree=c["qdeg"], cells_K=cells_k, n_rows=K.num_dofs, locator=c["locator"])
    return int(r["indices"].shape[0]), int(r["cell"].shape[0]), time.perf_counter() - t0


def cpu_run(cfg, seconds: float, threads: int) -> dict:
    """Times the numpy oracle on a bounded sample (the first cells of Γ) with `threads` forked workers.
    The GridLocator build (the analogue of dolfinx's tree build) is reported separately."""
    _cpu_prepare(cfg)
    n_cells = _CPU["gamma"].geometry.dofmap.shape[0]
    probe_n = min(n_cells, 32)
    _, _, dt = _cpu_chunk(np.arange(probe_n))
    rate = probe_n / max(dt, 1e-6)  # Γ cells per second on one thread
    n_sample = int(min(n_cells, max(probe_n, rate * seconds * threads)))
    chunks = [c for c in np.array_split(np.arange(n_sample, dtype=np.int32), threads) if c.size]
    t0 = time.perf_counter()
    if len(chunks) == 1:
        res = [_cpu_chunk(chunks[0])]
    else:
        import multiprocessing as mp

        with mp.get_context("fork").Pool(len(chunks)) as pool:
            res = pool.map(_cpu_chunk, chunks)
    wall = time.perf_counter() - t0
    return dict(nnz=sum(r[0] for r in res), points=sum(r[1] for r in res), wall_s=wall,
                locator_build_s=_CPU["locator_build_s"], sample_cells=n_sample, total_cells=n_cells)


def cpu_baseline(cfg, seconds: float, threads: int) -> dict:
    r = cpu_run(cfg, seconds, threads)
    return {"value": r["nnz"] / r["wall_s"], "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"first {r['sample_cells']} of {r['total_cells']} Γ cells of the same workload "
                      f"({r['points']} points) in {r['wall_s']:.1f} s; numpy oracle (CPU restatement, not the dolfinx "
                      f"reference); grid-locator build {r['locator_build_s']:.1f} s not included",
            "points_per_s": r["points"] / r["wall_s"]}


def run_reference(args):
    """--impl reference: the reference's algorithm on the host cores.  The dolfinx/PETSc stack is not
    installable in this image, so this is the oracle port with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = scaled(CONFIGS[args.config], args.scale)
    threads = os.cpu_count() or 1
    per_step = max(2.0, min(args.cpu_seconds, 60.0 / max(1, args.steps + args.warmup)))
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        last = cpu_run(cfg, per_step, threads)
        if i >= args.warmup:
            vals.append((last["nnz"], last["wall_s"]))
    nnz = sum(v[0] for v in vals)
    wall = sum(v[1] for v in vals)
    value = nnz / wall
    sample = (f"each step: first {last['sample_cells']} of {last['total_cells']} Γ cells ({last['points']} points), "
              f"numpy oracle port on {threads} forked workers; dolfinx/PETSc reference not installable here")
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, len(vals)), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": cfg["name"]},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the workload (testing only)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
