# scratch: where does the all-gather time go? (not part of the product)
import os, time, numpy as np, torch, torch.distributed as dist
from fenicsx_ii_b200.distributed import allgatherv_csr
from fenicsx_ii_b200.sparse import DeviceCSR
rank=int(os.environ["RANK"]); world=int(os.environ["WORLD_SIZE"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
os.environ["NCCL_DEBUG"]="WARN"
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{os.environ['LOCAL_RANK']}"))
n_rows=750000; nnz=13_000_000+rank*1000
indptr=torch.linspace(0,nnz,n_rows+1,device="cuda").to(torch.int64); indptr[-1]=nnz
indices=torch.randint(0,2_000_000,(nnz,),device="cuda",dtype=torch.int32); values=torch.rand(nnz,device="cuda",dtype=torch.float64)
def t(f,name,n=5):
    f(); torch.cuda.synchronize(); dist.barrier(); t0=time.perf_counter()
    for _ in range(n): r=f()
    torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/n*1e3
    if rank==0: print(f"{name}: {dt:.3f} ms")
    return r
g=t(lambda: allgatherv_csr(indptr,indices,values),"allgatherv_csr")
t(lambda: DeviceCSR.from_device_arrays(*g,(n_rows*world,2_000_000)),"from_device_arrays")
big=torch.empty(nnz*world,device="cuda",dtype=torch.float64)
def bc():
    ws=[dist.broadcast(big[r*nnz:(r+1)*nnz],src=r,async_op=True) for r in range(world)]
    [w.wait() for w in ws]
t(bc,"raw broadcasts f64")
out=torch.empty(nnz*world,device="cuda",dtype=torch.float64)
t(lambda: dist.all_gather_into_tensor(out, values[:nnz]),"all_gather_into_tensor f64 (equal sizes)")
dist.destroy_process_group()
