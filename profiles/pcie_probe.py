#!/usr/bin/env python
"""Host <-> device copy bandwidth of the box (what bounds bench.py's e2e and cold numbers):
pinned D2H / H2D with the allocating thread on each NUMA node, pageable H2D, threaded memcpy into pinned staging.

    python profiles/pcie_probe.py [--gib 1.0]
"""
import argparse
import glob
import os
import threading
import time

import numpy as np
import torch


def numa_nodes():
    nodes = {}
    for p in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
        try:
            with open(os.path.join(p, "cpulist")) as f:
                txt = f.read().strip()
        except OSError:
            continue
        cpus = []
        for part in txt.split(","):
            if not part:
                continue
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        nodes[int(os.path.basename(p)[4:])] = cpus
    return nodes


def gpu_numa(idx=0):
    try:
        bus = torch.cuda.get_device_properties(idx).pci_bus_id
        dom = torch.cuda.get_device_properties(idx).pci_domain_id
        dev = torch.cuda.get_device_properties(idx).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        with open(path) as f:
            return int(f.read().strip()), path
    except Exception as e:  # noqa: BLE001
        return None, repr(e)


def timed_copy(dst, src, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        dst.copy_(src, non_blocking=True)
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return src.numel() * src.element_size() / best / 1e6  # GB/s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=1.0)
    args = ap.parse_args()
    n = int(args.gib * (1 << 30))
    torch.cuda.set_device(0)
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    dev.zero_()
    nodes = numa_nodes()
    gn, how = gpu_numa(0)
    all_cpus = sorted(os.sched_getaffinity(0))
    print(f"cpus available {len(all_cpus)}; numa nodes { {k: len(v) for k, v in nodes.items()} }; gpu0 numa node {gn} ({how})")
    cases = [("default affinity", all_cpus)]
    for k, cpus in nodes.items():
        usable = [c for c in cpus if c in all_cpus]
        if usable:
            cases.append((f"thread on numa node {k}", usable))
    for name, cpus in cases:
        os.sched_setaffinity(0, cpus)
        host = torch.empty(n, dtype=torch.uint8).pin_memory()
        host.zero_()
        d2h = timed_copy(host, dev)
        h2d = timed_copy(dev, host)
        print(f"  pinned, {name:28s}: D2H {d2h:6.1f} GB/s   H2D {h2d:6.1f} GB/s")
        del host
    os.sched_setaffinity(0, all_cpus)
    # two concurrent D2H streams into two pinned buffers
    h1 = torch.empty(n // 2, dtype=torch.uint8).pin_memory()
    h2 = torch.empty(n // 2, dtype=torch.uint8).pin_memory()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(s1):
        h1.copy_(dev[: n // 2], non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(dev[n // 2:], non_blocking=True)
    torch.cuda.synchronize()
    print(f"  pinned, two streams concurrently      : D2H {n / (time.perf_counter() - t0) / 1e9:6.1f} GB/s")
    # pageable H2D (what a first call on numpy arrays pays)
    page = np.ones(n, dtype=np.uint8)
    pt = torch.from_numpy(page)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dev.copy_(pt)
    torch.cuda.synchronize()
    print(f"  pageable H2D (driver staging)         : {n / (time.perf_counter() - t0) / 1e9:6.1f} GB/s")
    # threaded memcpy into pinned staging
    stage = torch.empty(n, dtype=torch.uint8).pin_memory().numpy()
    for nt in (1, 2, 4, 8):
        parts = np.array_split(np.arange(n // 4096), nt)

        def work(p):
            lo, hi = int(p[0]) * 4096, (int(p[-1]) + 1) * 4096
            np.copyto(stage[lo:hi], page[lo:hi])

        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(p,)) for p in parts]
        [t.start() for t in th]
        [t.join() for t in th]
        print(f"  memcpy pageable -> pinned, {nt} threads  : {n / (time.perf_counter() - t0) / 1e9:6.1f} GB/s")


if __name__ == "__main__":
    main()
