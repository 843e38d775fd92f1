#!/usr/bin/env python
"""All-gather-v variants over NCCL for the CSR shards of config 5 (84.3 M entries in total), run under torchrun:
uneven all_gather (coalesced broadcasts, in place), batch_isend_irecv, padded all_gather_into_tensor (+ unpack copies).
Prints ms per variant (max over ranks) and GB/s received per rank."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    total = int(os.environ.get("GATHER_NNZ", "84290030"))
    base, rem = divmod(total, world)
    sizes = [base + (1 if r < rem else 0) + 1000 * ((r * 7) % 5) for r in range(world)]  # slightly uneven
    off = [0]
    for s in sizes:
        off.append(off[-1] + s)
    n = sizes[rank]
    idx = torch.randint(0, 1 << 24, (n,), dtype=torch.int32, device="cuda")
    val = torch.rand(n, dtype=torch.float64, device="cuda")
    g_idx = torch.empty(off[-1], dtype=torch.int32, device="cuda")
    g_val = torch.empty(off[-1], dtype=torch.float64, device="cuda")

    def uneven():
        dist.all_gather([g_idx[off[r]:off[r + 1]] for r in range(world)], idx)
        dist.all_gather([g_val[off[r]:off[r + 1]] for r in range(world)], val)

    def p2p():
        ops = []
        for buf, loc in ((g_idx, idx), (g_val, val)):
            buf[off[rank]:off[rank + 1]].copy_(loc)
            for r in range(world):
                if r != rank:
                    ops.append(dist.P2POp(dist.isend, loc, r))
                    ops.append(dist.P2POp(dist.irecv, buf[off[r]:off[r + 1]], r))
        for w in dist.batch_isend_irecv(ops):
            w.wait()

    mx = max(sizes)
    pad_i = torch.zeros(mx, dtype=torch.int32, device="cuda")
    pad_v = torch.zeros(mx, dtype=torch.float64, device="cuda")
    rec_i = torch.empty(world * mx, dtype=torch.int32, device="cuda")
    rec_v = torch.empty(world * mx, dtype=torch.float64, device="cuda")

    def padded():
        pad_i[:n].copy_(idx)
        pad_v[:n].copy_(val)
        dist.all_gather_into_tensor(rec_i, pad_i)
        dist.all_gather_into_tensor(rec_v, pad_v)
        for r in range(world):
            g_idx[off[r]:off[r + 1]].copy_(rec_i[r * mx:r * mx + sizes[r]])
            g_val[off[r]:off[r + 1]].copy_(rec_v[r * mx:r * mx + sizes[r]])

    def padded_only():
        dist.all_gather_into_tensor(rec_i, pad_i)
        dist.all_gather_into_tensor(rec_v, pad_v)

    tiny = torch.zeros(6, dtype=torch.int64, device="cuda")
    tiny_all = torch.empty(6 * world, dtype=torch.int64, device="cuda")

    def tiny_gather_host():
        dist.all_gather_into_tensor(tiny_all, tiny)
        return tiny_all.cpu()

    def tiny_allreduce_stream():
        dist.all_reduce(tiny)

    def tiny_from_list():
        t = torch.tensor([1, 2, 3, 4, 5, 6], dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(tiny_all, t)
        return tiny_all.cpu()

    res = {}
    for name, fn in (("tiny_all_gather+host_read", tiny_gather_host), ("tiny_all_reduce_stream_ordered", tiny_allreduce_stream),
                     ("tiny_from_list+all_gather+host_read", tiny_from_list), ("uneven_all_gather", uneven), ("batch_isend_irecv", p2p), ("padded_all_gather+unpack", padded),
                     ("padded_all_gather_only", padded_only)):
        try:
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b) / reps], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            res[name] = float(t.item())
        except Exception as e:  # noqa: BLE001
            res[name] = f"failed: {type(e).__name__}: {e}"[:200]
    # peer-to-peer pulls out of the library's own allocations (CUDA IPC): raw copy rate, no collectives inside the timing
    try:
        import ctypes as C

        import numpy as np

        from fenicsx_ii_b200 import _lib
        from fenicsx_ii_b200.sparse import DeviceCSR

        lib = _lib.load()
        ptr_t = torch.arange(n + 1, dtype=torch.int64, device="cuda")
        mat = DeviceCSR.from_device_arrays(ptr_t, idx, val, (n, 1 << 24))  # copies into the library's allocator
        p = [C.c_void_p(), C.c_void_p(), C.c_void_p()]
        _lib.check(lib.fxii_csr_device_ptrs(mat._h, C.byref(p[0]), C.byref(p[1]), C.byref(p[2])))
        hbuf = (C.c_ubyte * 64)()
        _lib.check(lib.fxii_ipc_export(p[2], C.cast(hbuf, C.c_void_p)))
        hs = torch.frombuffer(bytearray(hbuf), dtype=torch.uint8).cuda()
        allh = torch.empty(64 * world, dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(allh, hs)
        allh = allh.view(world, 64).cpu().numpy()
        peer = (rank + 1) % world
        raw = (C.c_ubyte * 64).from_buffer_copy(allh[peer].tobytes())
        mapped = C.c_void_p()
        _lib.check(lib.fxii_ipc_open(C.cast(raw, C.c_void_p), C.byref(mapped)))
        nbytes = 8 * min(sizes)
        dst = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        stream = _lib.current_stream()

        def pull_kernel():
            _lib.check(lib.fxii_copy_dev_async(C.c_void_p(dst.data_ptr()), mapped, nbytes, stream))

        def pull_memcpy():
            torch.cuda.synchronize()  # (cudaMemcpyAsync through the runtime: use torch's peer-agnostic path)
            C.CDLL("libcudart.so").cudaMemcpyAsync(C.c_void_p(dst.data_ptr()), mapped, C.c_size_t(nbytes), 4, C.c_void_p(stream))

        def local_copy():
            _lib.check(lib.fxii_copy_dev_async(C.c_void_p(dst.data_ptr()), p[2], nbytes, stream))

        for name, fn in (("p2p_pull_sm_kernel", pull_kernel), ("p2p_pull_cudaMemcpyAsync", pull_memcpy), ("local_copy_sm_kernel", local_copy)):
            try:
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                dist.barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(5):
                    fn()
                b.record()
                torch.cuda.synchronize()
                t = torch.tensor([a.elapsed_time(b) / 5], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                res[name] = f"{float(t.item()):8.3f} ms  {nbytes / 1e9 / (float(t.item()) * 1e-3):7.1f} GB/s (one peer, {nbytes / 1e6:.0f} MB)"
            except Exception as e:  # noqa: BLE001
                res[name] = f"failed: {type(e).__name__}: {e}"[:200]
        dist.barrier()
    except Exception as e:  # noqa: BLE001
        res["p2p"] = f"failed: {type(e).__name__}: {e}"[:300]
    # correctness of the in-place variant
    uneven()
    chk = torch.tensor([float(g_val[off[rank]:off[rank + 1]].sum() - val.sum())], device="cuda")
    if rank == 0:
        recv_gb = 12 * (off[-1] - n) / 1e9
        print(f"world={world} entries={off[-1]} received per rank {recv_gb:.3f} GB")
        for k, v in res.items():
            print(f"  {k:28s} {v if isinstance(v, str) else f'{v:8.3f} ms  {recv_gb / (v * 1e-3):7.1f} GB/s'}")
        print("  self-check", float(chk.item()))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
