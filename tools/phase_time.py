"""Per-phase cycle breakdown of k_chol (needs the library built with -DTB_PHASE_TIMING)."""
import ctypes, os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from python_stable_3d_truss_analysis_b200 import _lib
from python_stable_3d_truss_analysis_b200.truss import Truss
dev = torch.device("cuda:0"); td = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
t = Truss(3).LoadFromJSON(os.path.join(ROOT, "tests/golden/ref_data/bar-942_input_0.json"))
xyz, sup, conn, aed, force = t._pack(); plan = t._get_plan()
names = ["diag gemm", "diag epilogue+entries", "subpanel update", "base case 16x16", "subpanel trsm", "publish+rhs", "panel gemm", "panel epilogue+entries", "panel trsm+store", "backsolve matvec", "backsolve diag", "  base: load rows", "  base: column loop", "  base: write back"]
L = _lib.lib()
for B in [int(x) for x in sys.argv[1:]] or (148, 888):
    F = td(np.random.default_rng(0).uniform(-10, 10, size=(B, plan.N)))
    out = {k: torch.empty(B, plan.N if k in ("u", "ext") else plan.M, dtype=torch.float64, device=dev) for k in ("u", "ext", "axial")}
    out["weight"] = torch.empty(B, dtype=torch.float64, device=dev); out["info"] = torch.empty(B, dtype=torch.int32, device=dev)
    dx, da = td(xyz), td(aed)
    plan.solve_device(B, dx, F, aed=da, out=out); torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 16)(); L.tb_phase_read(buf)
    plan.solve_device(B, dx, F, aed=da, out=out); torch.cuda.synchronize()
    L.tb_phase_read(buf); v = np.array(buf[:14], dtype=np.float64) / B
    print(f"B={B}: cycles per system (thread 0 of its CTA), total {v[:11].sum():.0f}")
    for n, c in zip(names, v): print(f"   {n:26s} {c:10.0f}  {100*c/v[:11].sum():5.1f}%")
