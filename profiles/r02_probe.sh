set -x
nproc; lscpu | grep -i "numa\|model name\|socket" ; nvidia-smi topo -m 2>/dev/null | head -20
python profiles/pcie_probe.py --gib 1.0 > gpurun_out/r2_pcie_probe.txt 2>&1; cat gpurun_out/r2_pcie_probe.txt
FXII_TRACE=1 python - > gpurun_out/r2_cold_trace.txt 2>&1 <<'P'
import time, numpy as np, torch, sys
sys.path.insert(0, '.')
import bench
from fenicsx_ii_b200.interpolation import device_space
cfg = bench.CONFIGS[5]
wl = bench.build_workload(cfg, pinned=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
ds = device_space(wl["V"])
torch.cuda.synchronize()
print("device_space total s", time.perf_counter() - t0)
P
cat gpurun_out/r2_cold_trace.txt | grep -v "spgemm\|transpose" | head -40
