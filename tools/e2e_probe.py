"""Where the end-to-end time of tb_solve_host goes (bar-942 x1024): full call vs calls that skip outputs, raw copies."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from python_stable_3d_truss_analysis_b200 import _lib
from python_stable_3d_truss_analysis_b200.truss import Truss
import bench
B = 1024
data, F = bench.load_cases(B)
t = Truss(3).LoadFromJSON(data=data); xyz, sup, conn, aed, _ = t._pack(); plan = t._get_plan(sup, conn)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
h_xyz, h_aed, h_F = pin(xyz), pin(aed), pin(F)
N, M = plan.N, plan.M
def run(want, n=20):
    out = {k: pin(np.empty((B, N) if k in ("u", "ext") else (B, M) if k == "axial" else B)) for k in want}
    out["info"] = torch.empty(B, dtype=torch.int32).pin_memory().numpy()
    for _ in range(3): plan.solve_host(B, h_xyz, h_F, aed=h_aed, out=out, want=want)
    t0 = time.perf_counter()
    for _ in range(n): plan.solve_host(B, h_xyz, h_F, aed=h_aed, out=out, want=want)
    return (time.perf_counter() - t0) / n * 1e3
print("TB_HOST_CHUNKS =", os.environ.get("TB_HOST_CHUNKS", "(default 4)"))
print(f"all outputs       : {run(('u','ext','axial','weight')):.3f} ms")
print(f"weight only       : {run(('weight',)):.3f} ms   (H2D + compute, no bulk D2H)")
print(f"u only            : {run(('u',)):.3f} ms")
dev = torch.device('cuda:0')
big = torch.empty(B * (2 * N + M), dtype=torch.float64, device=dev); hbig = torch.empty(B * (2 * N + M), dtype=torch.float64).pin_memory()
torch.cuda.synchronize()
for _ in range(3): hbig.copy_(big, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): hbig.copy_(big, non_blocking=True); torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 20
print(f"raw D2H {hbig.numel()*8/1e6:.1f} MB : {dt*1e3:.3f} ms  ({hbig.numel()*8/dt/1e9:.1f} GB/s)")
hF = torch.from_numpy(h_F); dF = torch.empty_like(hF, device=dev)
t0 = time.perf_counter()
for _ in range(20): dF.copy_(hF, non_blocking=True); torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 20
print(f"raw H2D {hF.numel()*8/1e6:.1f} MB  : {dt*1e3:.3f} ms  ({hF.numel()*8/dt/1e9:.1f} GB/s)")
