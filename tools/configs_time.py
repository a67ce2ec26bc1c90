"""Device-resident timings of BASELINE.json configs 3, 4 and 5 (config 2 is bench.py; config 1 is a single bar-6 solve).
Prints one line per config with the per-kernel CUDA-event breakdown; development aid + source of DESIGN.md's table."""
import json, os, random, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from python_stable_3d_truss_analysis_b200 import _lib
from python_stable_3d_truss_analysis_b200.truss import Truss
from python_stable_3d_truss_analysis_b200.batch import type_table
from python_stable_3d_truss_analysis_b200.type import MemberType
from tests import helpers as H
import ctypes as C

dev = torch.device("cuda:0"); td = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
G = os.path.join(ROOT, "tests", "golden", "ref_data")
which = sys.argv[1:] or ["3", "4", "5"]

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); _lib.profile_enable(True); _lib.profile_read()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    pr = _lib.profile_read(); _lib.profile_enable(False)
    return min(ts), {k: round(v[0] / max(v[1], 1), 4) for k, v in pr.items() if v[1]}

if "1" in which:
    t = Truss(3).LoadFromJSON(os.path.join(G, "bar-6_input_0.json"))
    t.Solve(); t0 = time.perf_counter()
    for _ in range(50): t.Solve()
    print(f"config 1 bar-6 single Truss.Solve() through the host API: {(time.perf_counter()-t0)/50*1e3:.3f} ms per call")

if "3" in which:
    random.seed(0)
    types = [MemberType(i, random.uniform(1e7, 3e7), random.uniform(0.1, 1.0)) for i in range(1, 21)]
    genes = np.array([random.choices(range(20), k=72) for _ in range(8192)], dtype=np.int32)
    t = Truss(3).LoadFromJSON(os.path.join(G, "bar-72_input_0.json"))
    xyz, sup, conn, aed, force = t._pack(); plan = t._get_plan(); B = 8192
    o = {"fitness": torch.empty(B, dtype=torch.float64, device=dev), "flags": torch.empty(B, 2, dtype=torch.uint8, device=dev), "info": torch.empty(B, dtype=torch.int32, device=dev)}
    dx, df, dg, dt = td(xyz), td(force), td(genes), td(type_table(types))
    for path in (0, 2):
        plan.set_path(path)
        best, pr = timeit(lambda: plan.fitness_device(B, dx, df, dg, dt, 30000.0, 10.0, o), n=10)
        print(f"config 3 bar-72 GA x{B} fitness, path {path}: {best:.3f} ms -> {B/best*1e3:.0f} fitness/s  kernels(ms) {pr}")

if "4" in which:
    pool = [Truss(3).LoadFromJSON(data={k: g[k] for k in ("joint", "force", "member")}) for g in H.load_json("live_cube7_aug.json")]
    B = 65536
    jo, mo, xyz, sup, conn, aed, force, _ = H.ragged_pool_arrays(pool, B)
    d = dict(jo=td(jo), mo=td(mo), xyz=td(xyz), sup=td(sup), conn=td(conn), aed=td(aed), f=td(force))
    SJ, SM = int(jo[-1]), int(mo[-1])
    out = dict(u=torch.empty(SJ * 3, dtype=torch.float64, device=dev), ext=torch.empty(SJ * 3, dtype=torch.float64, device=dev),
               axial=torch.empty(SM, dtype=torch.float64, device=dev), weight=torch.empty(B, dtype=torch.float64, device=dev), info=torch.empty(B, dtype=torch.int32, device=dev))
    ri = _lib.TbRaggedIn(3, B, d["jo"].data_ptr(), d["mo"].data_ptr(), d["xyz"].data_ptr(), d["sup"].data_ptr(), d["conn"].data_ptr(), d["aed"].data_ptr(), d["f"].data_ptr(),
                         int(np.diff(jo).max()), int(np.diff(mo).max()))
    bo = _lib.TbBatchOut(out["u"].data_ptr(), out["ext"].data_ptr(), out["axial"].data_ptr(), out["weight"].data_ptr(), out["info"].data_ptr())
    st = torch.cuda.current_stream()
    best, pr = timeit(lambda: _lib.check(_lib.lib().tb_solve_ragged(C.byref(ri), C.byref(bo), C.c_void_p(st.cuda_stream))))
    n_free_mean = 3 * SJ / B
    byts = xyz.nbytes + sup.nbytes + conn.nbytes + aed.nbytes + force.nbytes + 2 * SJ * 3 * 8 + SM * 8 + B * 12
    print(f"config 4 cube-7 ragged x{B}: {best:.3f} ms -> {B/best*1e3:.0f} trusses/s; compulsory I/O {byts/1e6:.1f} MB -> {byts/best/1e6:.1f} GB/s; info any={bool(out['info'].any())}  kernels(ms) {pr}")

if "5" in which:
    t = H.cube_truss(12)
    xyz, sup, conn, aed, force = t._pack(); plan = t._get_plan(); info = plan.info
    for B in (8, 64, 256):
        rng = np.random.default_rng(5)
        aedb = np.repeat(aed[None], B, axis=0).copy(); aedb[:, :, 0] = rng.uniform(1.0, 20.0, size=(B, plan.M))
        xyzb = np.repeat(xyz[None], B, axis=0) + rng.normal(0, 5.0, size=(B,) + xyz.shape)
        dx, da, df = td(xyzb), td(aedb), td(force)
        out = {k: torch.empty(B, plan.N if k in ("u", "ext") else plan.M, dtype=torch.float64, device=dev) for k in ("u", "ext", "axial")}
        out["weight"] = torch.empty(B, dtype=torch.float64, device=dev); out["info"] = torch.empty(B, dtype=torch.int32, device=dev)
        best, pr = timeit(lambda: plan.solve_device(B, dx, df, aed=da, out=out), n=3, warm=1)
        print(f"config 5 cube 12^3 (n={plan.n}, reordered={info.reordered}, half-bw {info.half_bandwidth}, {info.n_tiles_nonzero}/{info.n_tiles} tiles, {info.n_tile_products} products) x{B}: "
              f"{best:.2f} ms -> {B/best*1e3:.1f} trusses/s; block-sparse {B*info.chol_flops/best/1e9:.2f} TFLOP/s, envelope {B*info.envelope_flops/best/1e9:.2f} TFLOP/s, "
              f"dense-equivalent {B*(plan.n**3/3)/best/1e9:.1f} TFLOP/s; info any={bool(out['info'].any())}  kernels(ms) {pr}", flush=True)
        del dx, da, out; torch.cuda.empty_cache()
