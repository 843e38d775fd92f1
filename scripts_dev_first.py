# scratch: first GPU shake-out (not part of the product)
import time, numpy as np, torch, traceback
from fenicsx_ii_b200 import Circle, Disk, PointwiseTrace, create_interpolation_matrix
from fenicsx_ii_b200 import synthetic as sy
from fenicsx_ii_b200.mesh import functionspace
from fenicsx_ii_b200.interpolation import device_space, device_rows
torch.cuda.set_device(0)
def run(name, V, K, op, reps=3):
    t=time.time(); ds=device_space(V); torch.cuda.synchronize(); print(name,'mesh build s',time.time()-t, ds.mesh.info())
    for i in range(reps):
        t=time.time(); rows=device_rows(K,op); torch.cuda.synchronize(); t1=time.time()
        csr,st=rows.assemble(ds); torch.cuda.synchronize(); t2=time.time()
        print(name, 'upload %.4f assemble %.4f'%(t1-t,t2-t1), dict(st), 'nnz',csr.nnz)
    return csr
try:
    mesh=sy.kuhn_box(64,64,64); V=functionspace(mesh,("Lagrange",1))
    g=sy.polyline_network(10000,1/64); K=functionspace(g,("Quadrature",10))
    P=run('cfg2',V,K,PointwiseTrace(g))
    g3,r3=sy.vessel_tree(100000,1/128)
    K3=functionspace(g3,("Quadrature",10))
    T=run('cfg2mesh-circle8',V,K3,Circle(g3,r3,degree=15))
    T=run('cfg2mesh-disk49',V,K3,Disk(g3,r3,degree=7))
    t=time.time(); Tt=T.transpose(); torch.cuda.synchronize(); print('transpose',time.time()-t, Tt.nnz)
    t=time.time(); Z,ip=Tt.matmult(T,return_ip=True); torch.cuda.synchronize(); print('spgemm',time.time()-t, Z.nnz, ip)
    t=time.time(); Z,ip=Tt.matmult(T,return_ip=True); torch.cuda.synchronize(); print('spgemm2',time.time()-t, Z.nnz, ip)
except Exception:
    traceback.print_exc()
