"""Development check of the fused two-sided band kernel: parity against the oracle on bar-942 and timings.
usage: python tools/ts_dev.py [--phase]      (TB_BAND_LEGACY=1 in the environment times the 16x16 kernels instead)"""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
PHASE = "--phase" in sys.argv
if PHASE:
    ALT = os.path.join(ROOT, "python_stable_3d_truss_analysis_b200", "csrc", "libtruss_b200_phase.so")
    os.environ["TB_LIB_PATH"] = ALT
import numpy as np, torch
from python_stable_3d_truss_analysis_b200 import _lib
from python_stable_3d_truss_analysis_b200.truss import Truss
from oracle import truss_oracle as orc

dev = torch.device("cuda:0"); td = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
data = json.load(open(os.path.join(ROOT, "tests/golden/ref_data/bar-942_input_0.json")))
t = Truss(3).LoadFromJSON(data=data)
xyz, sup, conn, aed, force = t._pack(); plan = t._get_plan()
prog = plan.ts_program()
print("program:", prog["info"] if prog else None, "legacy" if os.environ.get("TB_BAND_LEGACY") == "1" else "fused two-sided")
joints, support, conn_, aed_, force_ = orc.arrays_from_json(data, 3)
want = orc.solve(3, joints, support, conn_, aed_, force_)

def outs(B):
    out = {k: torch.empty(B, plan.N if k in ("u", "ext") else plan.M, dtype=torch.float64, device=dev) for k in ("u", "ext", "axial")}
    out["weight"] = torch.empty(B, dtype=torch.float64, device=dev); out["info"] = torch.empty(B, dtype=torch.int32, device=dev)
    return out

# parity: the fixture's own load vector and scaled copies
scales = (1.0, -2.0, 0.5, 3.0, 1.0, 1.0, 1.0)
F = td(np.stack([force.reshape(-1) * s for s in scales])); out = outs(len(scales))
plan.solve_device(len(scales), td(xyz), F, aed=td(aed), out=out); torch.cuda.synchronize()
print("info:", out["info"].cpu().numpy())
worst = 0.0
for b, s in enumerate(scales):
    for k in ("u", "ext", "axial"):
        err = orc.normwise_err(out[k][b].cpu().numpy(), want[k] * s); worst = max(worst, err)
        if err > 1e-9: print(f"  MISMATCH system {b} field {k}: {err:.3e}")
print(f"parity vs oracle, worst norm-wise error: {worst:.3e}", "OK" if worst <= 1e-9 else "FAIL")
same = all(torch.equal(out[k][0], out[k][j]) for k in ("u", "ext", "axial") for j in (4, 5, 6))
print("identical systems give identical bits:", same)

def timeit(fn, n=7, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts), sorted(ts)[len(ts) // 2]

names = ["setup", "column top (geometry, prefetch, wait)", "products", "assembly", "P = K - S, rows", "factor (8 pivots)",
         "stores + solves", "end of forward (hand-over)", "back substitution"]
L = _lib.lib()
for B in (1, 148, 1024, 2048, 8192):
    Fb = td(np.random.default_rng(0).uniform(-10, 10, size=(B, plan.N))); o = outs(B); dx, da = td(xyz), td(aed)
    _lib.profile_enable(True); _lib.profile_read()
    best, med = timeit(lambda: plan.solve_device(B, dx, Fb, aed=da, out=o))
    pr = _lib.profile_read(); _lib.profile_enable(False)
    best2, med2 = timeit(lambda: plan.solve_device(B, dx, Fb, aed=da, out=o))
    print(f"bar-942 x{B}: best {best2:.4f} ms median {med2:.4f} ms -> {B/best2*1e3:.0f} trusses/s; kernels(ms) " +
          str({k: round(v[0] / max(v[1], 1), 4) for k, v in pr.items() if v[1]}) + f"; info any={bool(o['info'].any())}")
    if PHASE:
        buf = (ctypes.c_ulonglong * 16)(); L.tb_ts_phase_read(buf)
        plan.solve_device(B, dx, Fb, aed=da, out=o); torch.cuda.synchronize()
        L.tb_ts_phase_read(buf); v = np.array(buf[:9], dtype=np.float64) / B
        ncol = prog["info"]["nblk"]
        print(f"  cycles per system, both warps summed: {v.sum():.0f}")
        for nme, c in zip(names, v): print(f"     {nme:40s} {c:10.0f}  {100*c/v.sum():5.1f}%  {c/ncol:8.0f}/col")
