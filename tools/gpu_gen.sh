#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gencube_device.py tests/test_gpu_parity.py -q -x -k "gencube or generated or random_choices or fixed_member or pipelined" --timeout 600 2>&1 | tail -30
python - <<'PY'
import time, torch
from python_stable_3d_truss_analysis_b200 import generate as G
for n in (1024, 65536):
    G.GenerateRandomCubeTrussesOnDevice(n, (5,5,5), (7,7), isDoStructuralAnalysis=True, seed=1, asNumpy=False)
    torch.cuda.synchronize(); t0=time.perf_counter()
    out=G.GenerateRandomCubeTrussesOnDevice(n, (5,5,5), (7,7), isDoStructuralAnalysis=True, seed=2, asNumpy=False)
    torch.cuda.synchronize(); t1=time.perf_counter()
    print(f"device generator + solve: {n} cube-7 trusses in {(t1-t0)*1e3:.2f} ms ({n/(t1-t0)/1e6:.2f} M trusses/s), solved {(out['info']==0).sum().item()}")
import random
t0=time.perf_counter(); G.GenerateRandomCubeTrusses(numCubeRange=(7,7), numEachRange=(1,50), isPrintMessage=False, seed=1); t1=time.perf_counter()
print(f"host generator: {(t1-t0)/50*1e3:.2f} ms per truss")
PY
