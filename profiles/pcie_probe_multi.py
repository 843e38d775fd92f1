#!/usr/bin/env python
"""Aggregate device->host bandwidth of the box with ALL ranks copying at once (what bounds bench.py's e2e at N > 1):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29520 profiles/pcie_probe_multi.py
"""
import os

import torch
import torch.distributed as dist


def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    n = 512 << 20
    dev = torch.empty(n, dtype=torch.uint8, device="cuda").zero_()
    host = torch.empty(n, dtype=torch.uint8).pin_memory()
    res = {}
    for name, d, s in (("D2H", host, dev), ("H2D", dev, host)):
        best = 1e9
        for _ in range(4):
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            d.copy_(s, non_blocking=True)
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        res[name] = n / best / 1e6
    t = torch.tensor([res["D2H"], res["H2D"]], device="cuda", dtype=torch.float64)
    allr = torch.empty(2 * world, device="cuda", dtype=torch.float64)
    dist.all_gather_into_tensor(allr, t)
    allr = allr.view(world, 2).cpu().numpy()
    if rank == 0:
        print(f"{world} ranks copying 512 MiB each at the same time (pinned host memory, best of 4), cpus {os.cpu_count()}")
        print("  D2H per rank GB/s:", [round(float(v), 1) for v in allr[:, 0]], " aggregate", round(float(allr[:, 0].sum()), 1))
        print("  H2D per rank GB/s:", [round(float(v), 1) for v in allr[:, 1]], " aggregate", round(float(allr[:, 1].sum()), 1))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
