"""Quick device-resident timings of the main configs (development aid, not the bench)."""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from python_stable_3d_truss_analysis_b200 import _lib
from python_stable_3d_truss_analysis_b200.truss import Truss
from python_stable_3d_truss_analysis_b200.batch import type_table
from python_stable_3d_truss_analysis_b200.type import MemberType

dev = torch.device("cuda:0")
td = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sorted(ts)[len(ts)//2]

for which, name in ((0, "DFMA"), (1, "DMMA m8n8k4"), (2, "DMMA m16n8k8")):
    tf, ms = _lib.fp64_peak(which, 8192)
    print(f"FP64 peak {name}: {tf:.2f} TFLOP/s ({ms:.3f} ms)")

G = os.path.join(ROOT, "tests", "golden", "ref_data")
# config 2: bar-942 x B load cases (independent K per system)
t = Truss(3).LoadFromJSON(os.path.join(G, "bar-942_input_0.json"))
xyz, sup, conn, aed, force = t._pack()
plan = t._get_plan()
for B in (1, 148, 1024, 8192):
    rng = np.random.default_rng(0)
    F = td(rng.uniform(-10, 10, size=(B, plan.N)))
    out = {k: torch.empty(B, plan.N if k in ("u", "ext") else plan.M, dtype=torch.float64, device=dev) for k in ("u", "ext", "axial")}
    out["weight"] = torch.empty(B, dtype=torch.float64, device=dev); out["info"] = torch.empty(B, dtype=torch.int32, device=dev)
    dx, da = td(xyz), td(aed)
    best, med = timeit(lambda: plan.solve_device(B, dx, F, aed=da, out=out))
    fl = B * (696**3 / 3 + 2 * 696**2)
    print(f"bar-942 x{B}: best {best:.3f} ms  median {med:.3f} ms  -> {B/best*1e3:.0f} trusses/s, {fl/best/1e9:.2f} TFLOP/s potrf-equivalent; info any={bool(out['info'].any())}")

for B in (1024, 8192):
    F = td(np.random.default_rng(0).uniform(-10, 10, size=(B, plan.N)))
    out = {k: torch.empty(B, plan.N if k in ("u", "ext") else plan.M, dtype=torch.float64, device=dev) for k in ("u", "ext", "axial")}
    out["weight"] = torch.empty(B, dtype=torch.float64, device=dev); out["info"] = torch.empty(B, dtype=torch.int32, device=dev)
    dx, da = td(xyz), td(aed)
    _lib.profile_enable(True); _lib.profile_read()
    best, med = timeit(lambda: plan.solve_device(B, dx, F, aed=da, out=out, shared_factor=True))
    pr = _lib.profile_read(); _lib.profile_enable(False)
    print(f"bar-942 x{B} load cases, shared factor: best {best:.3f} ms -> {B/best*1e3:.0f} load cases/s; kernels(ms) " +
          str({k: round(v[0] / max(v[1], 1), 4) for k, v in pr.items() if v[1]}) + f"; info any={bool(out['info'].any())}")

# config 3: bar-72 x 8192 genes fitness
import random
random.seed(0)
types = [MemberType(i, random.uniform(1e7, 3e7), random.uniform(0.1, 1.0)) for i in range(1, 21)]
genes = np.array([random.choices(range(20), k=72) for _ in range(8192)], dtype=np.int32)
t = Truss(3).LoadFromJSON(os.path.join(G, "bar-72_input_0.json"))
xyz, sup, conn, aed, force = t._pack()
plan = t._get_plan()
B = 8192
o = {"fitness": torch.empty(B, dtype=torch.float64, device=dev), "flags": torch.empty(B, 2, dtype=torch.uint8, device=dev), "info": torch.empty(B, dtype=torch.int32, device=dev)}
dx, df, dg, dt = td(xyz), td(force), td(genes), td(type_table(types))
best, med = timeit(lambda: plan.fitness_device(B, dx, df, dg, dt, 30000.0, 10.0, o), n=10)
print(f"bar-72 GA x{B}: best {best:.3f} ms median {med:.3f} ms -> {B/best*1e3:.0f} fitness/s")
for path in (1, 2):
    plan.set_path(path)
    best, med = timeit(lambda: plan.fitness_device(B, dx, df, dg, dt, 30000.0, 10.0, o), n=5)
    print(f"bar-72 GA x{B} (path {path}): best {best:.3f} ms -> {B/best*1e3:.0f} fitness/s")

# config 3 as a loop: GA generations on the device (tb_fitness + tb_ga_step per generation) vs the host GA class
from python_stable_3d_truss_analysis_b200.ga import GA
import time as _t
for nPop, nElite in ((8192, 1024),):
    ga = GA(t, types, allowStress=30000., allowDisplace=10., nIteration=40, nPatience=1000, nPop=nPop, nElite=nElite)
    ga.EvolveOnDevice(isPrintMessage=False, seed=1)
    torch.cuda.synchronize(); t0 = _t.perf_counter()
    g, info, pop, hist = ga.EvolveOnDevice(isPrintMessage=False, seed=1)
    dt = _t.perf_counter() - t0
    print(f"bar-72 GA nPop={nPop}: device loop {dt/40*1e3:.3f} ms per generation ({40/dt:.0f} generations/s, {nPop*40/dt/1e6:.1f} M fitness/s incl. ranking + update); best {hist[0]:.1f} -> {hist[-1]:.1f}")
    ga2 = GA(t, types, allowStress=30000., allowDisplace=10., nIteration=3, nPatience=1000, nPop=nPop, nElite=nElite)
    random.seed(1); t0 = _t.perf_counter(); ga2.Evolve(isPrintMessage=False); dt2 = _t.perf_counter() - t0
    print(f"bar-72 GA nPop={nPop}: host loop (GA.Evolve, batched fitness, Python operators) {dt2/3*1e3:.1f} ms per generation")

# config 4 as a pipeline: a pool of cube-7 trusses expanded to 65536 augmented trusses and solved, all on the device
import copy as _copy, glob as _glob
from python_stable_3d_truss_analysis_b200.generate import GenerateAugmentedDataset
pool = [Truss(3).LoadFromJSON(data=json.load(open(f))) for f in sorted(_glob.glob(os.path.join(ROOT, "tests", "golden", "ref_generate", "cube-7_case_*.json")))]
kw = dict(moveToCentroid=True, translateRange=(-30., 30.), noiseStds=[10., 10., 10.], resetPin=(5, 0.6), seed=42, asNumpy=False)
GenerateAugmentedDataset(pool, 65536, **kw); torch.cuda.synchronize()
t0 = _t.perf_counter(); ds = GenerateAugmentedDataset(pool, 65536, **kw); torch.cuda.synchronize(); dt = _t.perf_counter() - t0
ok = int((ds["info"] == 0).sum().item())
print(f"cube-7 pool of {len(pool)} -> 65536 augmented trusses generated + solved on the device: {dt*1e3:.2f} ms wall ({65536/dt/1e6:.2f} M trusses/s incl. pool upload and offsets), {ok} solved, {65536-ok} fail the counting rule")
