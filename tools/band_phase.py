"""Per-phase cycle breakdown of k_band2 (instrumented build: -DTB_PHASE_TIMING, loaded through TB_LIB_PATH).
usage: python tools/band_phase.py [B ...]   (builds the instrumented library next to the product one if missing)"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
ALT = os.path.join(ROOT, "python_stable_3d_truss_analysis_b200", "csrc", "libtruss_b200_phase.so")
os.environ["TB_LIB_PATH"] = ALT
import numpy as np, torch
from python_stable_3d_truss_analysis_b200 import _lib
if not os.path.exists(ALT):
    _lib.build(out=ALT, extra_flags=["-DTB_PHASE_TIMING"])
from python_stable_3d_truss_analysis_b200.truss import Truss
dev = torch.device("cuda:0"); td = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
t = Truss(3).LoadFromJSON(os.path.join(ROOT, "tests/golden/ref_data/bar-942_input_0.json"))
xyz, sup, conn, aed, force = t._pack(); plan = t._get_plan()
names = ["F: diagonal products (d=1)", "F: P = K - S", "F: 16x16 factor + W", "-", "wait X",
         "trsm rb=1", "wait Y (for warp T)", "back substitution (all columns)"]
L = _lib.lib()
for B in [int(x) for x in sys.argv[1:]] or (1, 148, 1024):
    F = td(np.random.default_rng(0).uniform(-10, 10, size=(B, plan.N)))
    out = {k: torch.empty(B, plan.N if k in ("u", "ext") else plan.M, dtype=torch.float64, device=dev) for k in ("u", "ext", "axial")}
    out["weight"] = torch.empty(B, dtype=torch.float64, device=dev); out["info"] = torch.empty(B, dtype=torch.int32, device=dev)
    dx, da = td(xyz), td(aed)
    plan.solve_device(B, dx, F, aed=da, out=out); torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 16)(); L.tb_band_phase_read(buf)
    plan.solve_device(B, dx, F, aed=da, out=out); torch.cuda.synchronize()
    L.tb_band_phase_read(buf); v = np.array(buf[:8], dtype=np.float64) / B
    print(f"B={B}: cycles per system (factor warp), total {v.sum():.0f} = {v.sum()/44:.0f} per block column")
    for n, c in zip(names, v): print(f"   {n:26s} {c:10.0f}  {100*c/v.sum():5.1f}%  {c/44:8.0f}/col")
