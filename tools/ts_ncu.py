"""One batch of bar-942 through the band path (for ncu captures).  usage: python tools/ts_ncu.py B [reps]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from python_stable_3d_truss_analysis_b200.truss import Truss
dev = torch.device("cuda:0"); td = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
t = Truss(3).LoadFromJSON(os.path.join(ROOT, "tests/golden/ref_data/bar-942_input_0.json"))
xyz, sup, conn, aed, force = t._pack(); plan = t._get_plan()
F = td(np.random.default_rng(0).uniform(-10, 10, size=(B, plan.N)))
out = {k: torch.empty(B, plan.N if k in ("u", "ext") else plan.M, dtype=torch.float64, device=dev) for k in ("u", "ext", "axial")}
out["weight"] = torch.empty(B, dtype=torch.float64, device=dev); out["info"] = torch.empty(B, dtype=torch.int32, device=dev)
dx, da = td(xyz), td(aed)
for _ in range(reps):
    plan.solve_device(B, dx, F, aed=da, out=out)
torch.cuda.synchronize()
print("done", bool(out["info"].any()))
