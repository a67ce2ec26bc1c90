"""Batch-size sweep of the blocked pipeline (development aid)."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from python_stable_3d_truss_analysis_b200 import _lib
from python_stable_3d_truss_analysis_b200.truss import Truss
dev = torch.device("cuda:0"); td = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
t = Truss(3).LoadFromJSON(os.path.join(ROOT, "tests/golden/ref_data/bar-942_input_0.json"))
xyz, sup, conn, aed, force = t._pack(); plan = t._get_plan()
_lib.profile_enable(True)
for B in [int(x) for x in sys.argv[1:]] or (148, 296, 444, 592, 888, 1024, 1036, 1332):
    F = td(np.random.default_rng(0).uniform(-10, 10, size=(B, plan.N)))
    out = {k: torch.empty(B, plan.N if k in ("u", "ext") else plan.M, dtype=torch.float64, device=dev) for k in ("u", "ext", "axial")}
    out["weight"] = torch.empty(B, dtype=torch.float64, device=dev); out["info"] = torch.empty(B, dtype=torch.int32, device=dev)
    dx, da = td(xyz), td(aed)
    for _ in range(2): plan.solve_device(B, dx, F, aed=da, out=out)
    torch.cuda.synchronize(); _lib.profile_read()
    for _ in range(5): plan.solve_device(B, dx, F, aed=da, out=out)
    torch.cuda.synchronize(); pr = _lib.profile_read()
    ms = {k: v[0] / max(v[1], 1) for k, v in pr.items() if v[1]}
    print(f"B={B:5d}  chol {ms['chol']:.3f} ms  ({ms['chol']/B*148:.4f} ms per system-SM)  other: " + " ".join(f"{k}={v:.3f}" for k, v in ms.items() if k != 'chol'), flush=True)
