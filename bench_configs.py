"""The other GPU configurations of BASELINE.json, measured inside bench.py next to the headline (config 2):

  config 3  bar-72 GA population of 8192 member-type variants (ga.py:139-160 fitness), sharded over the ranks (strong scaling)
  config 4  65 536 cube-7 trusses with augmentation-style joint noise, ragged batch (generate.py:354-357), sharded (strong)
  config 5  cube 12^3 full grid (n = 6084 free DOF, 14 868 members), 256 systems per GPU (weak scaling)

Each returns {"value", "unit", "ms_per_step", "scaling", "roofline", "e2e", "parity"}: device-resident throughput (CUDA
events, max over ranks), the dominant kernel against its roofline, the same batch through the host entry point of the C
ABI, and the error of sampled systems against the oracle (the checker only).  Inputs follow SURVEY.md section 8(d).
"""
from __future__ import annotations

import ctypes as C
import glob
import json
import os
import random
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
G = os.path.join(ROOT, "tests", "golden")


def _timed(torch, dist, world, fn, steps, flush, dev):
    """steps device-timed calls of fn (L2 flushed between them, outside the events); max over ranks of the total."""
    stream = torch.cuda.current_stream()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    for e0, e1 in ev:
        flush.zero_()
        e0.record(stream)
        fn()
        e1.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([sum(e0.elapsed_time(e1) for e0, e1 in ev)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / steps


def _wall(torch, dist, world, fn, steps, dev):
    fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / steps


def _kernel_ms(_lib, fn, torch, steps):
    """Average per-launch time of every library kernel over `steps` calls of fn (CUDA events inside the library)."""
    _lib.profile_enable(True)
    _lib.profile_read()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    pr = _lib.profile_read()
    _lib.profile_enable(False)
    return {k: v[0] / v[1] for k, v in pr.items() if v[1]}


def config3(torch, dist, rank, world, dev, steps, flush, peaks):
    from oracle import truss_oracle as orc
    from python_stable_3d_truss_analysis_b200 import _lib
    from python_stable_3d_truss_analysis_b200.batch import type_table
    from python_stable_3d_truss_analysis_b200.truss import Truss
    from python_stable_3d_truss_analysis_b200.type import MemberType

    POP, ALLOW_S, ALLOW_D = 8192, 30000.0, 10.0
    random.seed(0)                                           # example.py:175-205 recipe
    types = [MemberType(i, random.uniform(1e7, 3e7), random.uniform(0.1, 1.0)) for i in range(1, 21)]
    genes = np.array([random.choices(range(20), k=72) for _ in range(POP)], dtype=np.int32)
    data = json.load(open(os.path.join(G, "ref_data", "bar-72_input_0.json")))
    t = Truss(3).LoadFromJSON(data=data)
    xyz, sup, conn, aed, force = t._pack()
    plan = t._get_plan()
    B = POP // world
    mine = genes[rank * B:(rank + 1) * B]
    td = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    tt = type_table(types)
    dx, df, dg, dt = td(xyz), td(force), td(mine), td(tt)
    o = {"fitness": torch.empty(B, dtype=torch.float64, device=dev), "flags": torch.empty(B, 2, dtype=torch.uint8, device=dev),
         "info": torch.empty(B, dtype=torch.int32, device=dev)}
    gathered = [torch.empty_like(o["fitness"]) for _ in range(world)] if (world > 1 and rank == 0) else None

    def step():
        plan.fitness_device(B, dx, df, dg, dt, ALLOW_S, ALLOW_D, o)
        if world > 1:                                        # north_star: NCCL only gathers the fitness back to rank 0
            dist.gather(o["fitness"], gathered, dst=0)

    for _ in range(3):
        step()
    ms = _timed(torch, dist, world, step, steps, flush, dev)
    kms = _kernel_ms(_lib, lambda: plan.fitness_device(B, dx, df, dg, dt, ALLOW_S, ALLOW_D, o), torch, steps)
    # end to end: host genes in, fitness + flags out (tb_fitness_host)
    h_mine = _lib.pinned_empty(mine.shape, np.int32)         # page-locked genes in, fitness / flags out
    h_mine[...] = mine
    ho = {"fitness": _lib.pinned_empty((B,)), "flags": _lib.pinned_empty((B, 2), np.uint8), "info": _lib.pinned_empty((B,), np.int32)}
    e2e_s = _wall(torch, dist, world, lambda: plan.fitness_host(B, xyz, force, h_mine, tt, ALLOW_S, ALLOW_D, out=ho), max(3, steps), dev)
    # parity: sampled genes against the oracle's GA.GetFitness (ga.py:139-149)
    worst = 0.0
    if rank == 0:
        joints, support, conn_o, _, force_o = orc.arrays_from_json(data, 3)
        fit = o["fitness"].cpu().numpy()
        flags = o["flags"].cpu().numpy()
        for b in (0, B // 2, B - 1):
            want = orc.fitness(3, joints, support, conn_o, mine[b], tt, force_o, ALLOW_S, ALLOW_D)
            worst = max(worst, abs(fit[b] - want[0]) / max(1.0, abs(want[0])))
            assert (bool(flags[b][0]), bool(flags[b][1])) == (bool(want[1]), bool(want[2])), "GA feasibility flags differ from the oracle"
        assert worst <= 1e-9, f"config 3 parity: {worst:.3e}"
    n = plan.n
    flops = n ** 3 / 3.0 + n ** 2 / 2.0 + n / 6.0 + 2.0 * n * n
    kname = "k_dense16" if "small" in kms else "band/tiled pipeline"
    k_ms = kms.get("small", ms)
    ach = B * flops / (k_ms * 1e-3) / 1e12
    return {"workload": f"bar-72 (n={n}) GA population of {POP} member-type variants, fitness + feasibility flags (ga.py:139-160)",
            "value": POP / (ms * 1e-3), "unit": "fitness evaluations/s", "ms_per_step": ms, "batch_per_gpu": B, "scaling": "strong",
            "roofline": {"kernel": kname + " (fused assembly + Cholesky + recovery + fitness, one warp per truss)", "bound": "fp64",
                         "achieved": ach, "peak": peaks["dfma"], "unit": "TFLOP/s", "frac": ach / peaks["dfma"],
                         "flops_per_system": flops, "flops_model": "dense potrf + two triangular solves (SURVEY 8d)",
                         "ms_per_launch": k_ms, "peak_source": "FP64 DFMA microbenchmark of this run"},
            "e2e": {"value": POP / e2e_s, "unit": "fitness evaluations/s", "h2d_bytes_per_step": int(mine.nbytes + xyz.nbytes + force.nbytes + tt.nbytes),
                    "d2h_bytes_per_step": int(B * (8 + 2 + 4)), "api": "tb_fitness_host"},
            "parity": {"max_rel_err_fitness": worst, "systems": 3, "oracle": "oracle.truss_oracle.fitness"},
            "multi_gpu": "population sharded by contiguous blocks; NCCL gather of the fitness vector to rank 0 inside the step" if world > 1 else "single GPU"}


def config4(torch, dist, rank, world, dev, steps, flush, peaks):
    from oracle import truss_oracle as orc
    from python_stable_3d_truss_analysis_b200 import _lib
    from python_stable_3d_truss_analysis_b200.truss import Truss
    from tests import helpers as H

    TOTAL = 65536
    pool = [Truss(3).LoadFromJSON(data={k: g[k] for k in ("joint", "force", "member")}) for g in H.load_json("live_cube7_aug.json")]
    B = TOTAL // world
    # the pool tiled to TOTAL systems with fresh Gaussian joint jitter; this rank's block of it
    jo, mo, xyz, sup, conn, aed, force, which = H.ragged_pool_arrays(pool, TOTAL)
    j0, j1, m0, m1 = int(jo[rank * B]), int(jo[(rank + 1) * B]), int(mo[rank * B]), int(mo[(rank + 1) * B])
    jo_r = (jo[rank * B:(rank + 1) * B + 1] - j0).astype(np.int64)
    mo_r = (mo[rank * B:(rank + 1) * B + 1] - m0).astype(np.int64)
    xyz_r, sup_r, conn_r = xyz[3 * j0:3 * j1], sup[j0:j1], conn[2 * m0:2 * m1]
    aed_r, force_r = aed[3 * m0:3 * m1], force[3 * j0:3 * j1]
    td = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    d = dict(jo=td(jo_r), mo=td(mo_r), xyz=td(xyz_r), sup=td(sup_r), conn=td(conn_r), aed=td(aed_r), f=td(force_r))
    SJ, SM = j1 - j0, m1 - m0
    flat = torch.empty(2 * SJ * 3 + SM, dtype=torch.float64, device=dev)
    out = dict(u=flat[:SJ * 3], ext=flat[SJ * 3:2 * SJ * 3], axial=flat[2 * SJ * 3:], weight=torch.empty(B, dtype=torch.float64, device=dev),
               info=torch.empty(B, dtype=torch.int32, device=dev))
    ri = _lib.TbRaggedIn(3, B, d["jo"].data_ptr(), d["mo"].data_ptr(), d["xyz"].data_ptr(), d["sup"].data_ptr(), d["conn"].data_ptr(),
                         d["aed"].data_ptr(), d["f"].data_ptr(), int(np.diff(jo_r).max()), int(np.diff(mo_r).max()))
    bo = _lib.TbBatchOut(out["u"].data_ptr(), out["ext"].data_ptr(), out["axial"].data_ptr(), out["weight"].data_ptr(), out["info"].data_ptr())
    st = torch.cuda.current_stream()
    sizes = None
    gathered = None
    if world > 1:
        sz = torch.tensor([flat.numel()], dtype=torch.int64, device=dev)
        allsz = [torch.zeros_like(sz) for _ in range(world)]
        dist.all_gather(allsz, sz)
        sizes = [int(x.item()) for x in allsz]
        if rank == 0:
            gathered = [torch.empty(s, dtype=torch.float64, device=dev) for s in sizes]

    def solve():
        _lib.check(_lib.lib().tb_solve_ragged(C.byref(ri), C.byref(bo), C.c_void_p(st.cuda_stream)))

    def step():
        solve()
        if world > 1:                                        # ragged blocks differ in size: point-to-point sends to rank 0
            if rank == 0:
                reqs = [dist.irecv(gathered[r], src=r) for r in range(1, world)]
                for q in reqs:
                    q.wait()
            else:
                dist.send(flat, dst=0)

    for _ in range(3):
        step()
    ms = _timed(torch, dist, world, step, steps, flush, dev)
    kms = _kernel_ms(_lib, solve, torch, steps)
    def pin(a):                                              # page-locked copies: the host entry point copies at PCIe speed
        h = _lib.pinned_empty(a.shape, a.dtype)
        h[...] = a
        return h
    hp = [pin(np.ascontiguousarray(x)) for x in (jo_r, mo_r, xyz_r, sup_r, conn_r, aed_r, force_r)]
    ho = {"u": _lib.pinned_empty((SJ * 3,)), "ext": _lib.pinned_empty((SJ * 3,)), "axial": _lib.pinned_empty((SM,)),
          "weight": _lib.pinned_empty((B,)), "info": _lib.pinned_empty((B,), np.int32)}
    e2e_s = _wall(torch, dist, world,
                  lambda: _lib.solve_ragged_host(3, *hp, want=("u", "ext", "axial", "weight"), out=ho),
                  max(3, min(steps, 5)), dev)
    assert np.array_equal(ho["u"], out["u"].cpu().numpy()), "config 4: host and device entry points disagree"
    worst = 0.0
    info = out["info"].cpu().numpy()
    solved = int((info == 0).sum())
    if rank == 0:
        u, ext, ax = out["u"].cpu().numpy(), out["ext"].cpu().numpy(), out["axial"].cpu().numpy()
        for b in (0, B // 3, B - 1):
            if info[b] != 0:
                continue
            a0, a1, b0, b1 = int(jo_r[b]), int(jo_r[b + 1]), int(mo_r[b]), int(mo_r[b + 1])
            want = orc.solve(3, xyz_r[3 * a0:3 * a1].reshape(-1, 3), sup_r[a0:a1], conn_r[2 * b0:2 * b1].reshape(-1, 2),
                             aed_r[3 * b0:3 * b1].reshape(-1, 3), force_r[3 * a0:3 * a1])
            for k, got in (("u", u[3 * a0:3 * a1]), ("ext", ext[3 * a0:3 * a1]), ("axial", ax[b0:b1])):
                worst = max(worst, orc.normwise_err(got, want[k]))
        assert worst <= 1e-9, f"config 4 parity: {worst:.3e}"
    # compulsory I/O of the fused kernel: inputs that vary per truss + outputs
    byts = xyz_r.nbytes + sup_r.nbytes + conn_r.nbytes + aed_r.nbytes + force_r.nbytes + flat.numel() * 8 + B * 12
    k_ms = kms.get("small", ms)
    nbar = 3.0 * SJ / B
    flops = float(np.sum((3.0 * np.diff(jo_r)) ** 3 / 3.0 + 2.0 * (3.0 * np.diff(jo_r)) ** 2))   # upper bound: every DOF free
    ach = flops / (k_ms * 1e-3) / 1e12
    # the generator's whole loop on the device (generate.py:338-372): topologies drawn by tb_gencube, packed, solved
    from python_stable_3d_truss_analysis_b200 import generate as G
    gen_kw = dict(gridRange=(5, 5, 5), numCubeRange=(7, 7), isDoStructuralAnalysis=True, asNumpy=False)
    G.GenerateRandomCubeTrussesOnDevice(B, seed=100 + rank, **gen_kw)
    gen_s = _wall(torch, dist, world, lambda: G.GenerateRandomCubeTrussesOnDevice(B, seed=200 + rank, **gen_kw), max(3, min(steps, 5)), dev)
    gen = G.GenerateRandomCubeTrussesOnDevice(B, seed=200 + rank, **gen_kw)
    gen_solved = int((gen["info"] == 0).sum().item())
    return {"generated_on_device": {"value": TOTAL / gen_s, "unit": "trusses generated + solved/s", "ms_per_step": gen_s * 1e3,
                                    "solved": gen_solved, "of": B,
                                    "what": "GenerateRandomCubeTrussesOnDevice: tb_gencube (one thread per truss) + tb_gencube_pack + "
                                            "tb_solve_ragged, every topology distinct, wall clock incl. the offset prefix sums"},
            "workload": f"{TOTAL} cube-7 trusses (pool of {len(pool)} generated topologies, Gaussian joint jitter sigma 10), ragged batch, full results",
            "value": TOTAL / (ms * 1e-3), "unit": "trusses/s", "ms_per_step": ms, "batch_per_gpu": B, "scaling": "strong",
            "solved": solved, "mean_dof": nbar,
            "roofline": {"kernel": "k_dense16 (fused, one warp per truss, ragged)", "bound": "fp64", "achieved": ach, "peak": peaks["dfma"],
                         "unit": "TFLOP/s", "frac": ach / peaks["dfma"], "flops_model": "dense potrf + solves on d*nJ DOFs per truss (upper bound)",
                         "ms_per_launch": k_ms, "hbm_view": {"achieved": byts / (k_ms * 1e-3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                                                             "frac": byts / (k_ms * 1e-3) / 1e9 / peaks["hbm"], "bytes_per_launch": int(byts)},
                         "peak_source": "FP64 DFMA microbenchmark of this run; MEASURED_PEAKS.json hbm_gbs"},
            "e2e": {"value": TOTAL / e2e_s, "unit": "trusses/s", "h2d_bytes_per_step": int(xyz_r.nbytes + sup_r.nbytes + conn_r.nbytes + aed_r.nbytes + force_r.nbytes + jo_r.nbytes + mo_r.nbytes),
                    "d2h_bytes_per_step": int(flat.numel() * 8 + B * 12), "api": "tb_solve_ragged_host"},
            "parity": {"max_normwise_err": worst, "systems": 3, "oracle": "oracle.truss_oracle.solve"},
            "multi_gpu": "batch sharded by contiguous blocks; results sent to rank 0 over NCCL inside the step" if world > 1 else "single GPU"}


def config5(torch, dist, rank, world, dev, steps, flush, peaks):
    from oracle import truss_oracle as orc
    from python_stable_3d_truss_analysis_b200 import _lib
    from tests import helpers as H

    B = 256
    t = H.cube_truss(12)
    xyz, sup, conn, aed, force = t._pack()
    plan = t._get_plan()
    info = plan.info
    rng = np.random.default_rng(5 + rank)                    # member areas ~ U(1,20), joint jitter sigma 5
    aedb = np.repeat(aed[None], B, axis=0).copy()
    aedb[:, :, 0] = rng.uniform(1.0, 20.0, size=(B, plan.M))
    xyzb = np.repeat(xyz[None], B, axis=0) + rng.normal(0, 5.0, size=(B,) + xyz.shape)
    td = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    dx, da, df = td(xyzb), td(aedb), td(force)
    N, M = plan.N, plan.M
    flat = torch.empty(B * (2 * N + M), dtype=torch.float64, device=dev)
    out = {"u": flat[:B * N].view(B, N), "ext": flat[B * N:2 * B * N].view(B, N), "axial": flat[2 * B * N:].view(B, M),
           "weight": torch.empty(B, dtype=torch.float64, device=dev), "info": torch.empty(B, dtype=torch.int32, device=dev)}
    gathered = [torch.empty_like(flat) for _ in range(world)] if (world > 1 and rank == 0) else None

    def solve():
        plan.solve_device(B, dx, df, aed=da, out=out)

    def step():
        solve()
        if world > 1:
            dist.gather(flat, gathered, dst=0)

    step()
    nst = max(2, min(steps, 3))
    ms = _timed(torch, dist, world, step, nst, flush, dev)
    kms = _kernel_ms(_lib, solve, torch, 1)
    # end to end through tb_solve_host on a slice of the batch (the full batch's host buffers are 57 MB out, 42 MB in)
    Be = 32
    ho = {}
    e2e_s = _wall(torch, dist, world, lambda: plan.solve_host(Be, xyzb[:Be], force, aed=aedb[:Be], out=ho), 2, dev)
    worst = 0.0
    assert not bool(out["info"].any().item()), "config 5: a system failed to factorise"
    if rank == 0:
        for b in (0, B - 1):
            want = orc.solve_closed_form(3, xyzb[b], sup, conn, aedb[b], force)
            for k in ("u", "ext", "axial"):
                worst = max(worst, orc.normwise_err(out[k][b].cpu().numpy(), want[k]))
        assert worst <= 1e-9, f"config 5 parity: {worst:.3e}"
    k_ms = kms.get("chol", ms)
    ach = B * float(info.envelope_flops) / (k_ms * 1e-3) / 1e12
    return {"workload": f"cube 12^3 full grid (n={plan.n} free DOF, {M} members) x {B} systems per GPU, member areas and joint positions vary per system",
            "value": B * world / (ms * 1e-3), "unit": "trusses/s", "ms_per_step": ms, "batch_per_gpu": B, "scaling": "weak",
            "roofline": {"kernel": "k_chol (tiled 64x64 block-sparse Cholesky, fused assembly) + substitutions", "bound": "tensor", "achieved": ach,
                         "peak": peaks["dmma"], "unit": "TFLOP/s", "frac": ach / peaks["dmma"],
                         "flops_per_system": float(info.envelope_flops), "flops_model": "envelope Cholesky + two triangular solves (plan.envelope_flops)",
                         "executed_block_sparse_tflops": B * float(info.chol_flops) / (k_ms * 1e-3) / 1e12,
                         "ms_per_launch": k_ms, "half_bandwidth": int(info.half_bandwidth), "tiles_nonzero": int(info.n_tiles_nonzero),
                         "tiles": int(info.n_tiles), "peak_source": "FP64 DMMA microbenchmark of this run"},
            "kernels_ms": kms,
            "e2e": {"value": Be * world / e2e_s, "unit": "trusses/s", "batch": Be, "h2d_bytes_per_step": int(xyzb[:Be].nbytes + aedb[:Be].nbytes + force.nbytes),
                    "d2h_bytes_per_step": int(Be * (2 * N + M + 1) * 8 + Be * 4), "api": "tb_solve_host"},
            "parity": {"max_normwise_err": worst, "systems": 2, "oracle": "oracle.truss_oracle.solve_closed_form"},
            "multi_gpu": "256 systems per rank; NCCL gather of u/ext/axial to rank 0 inside the step" if world > 1 else "single GPU"}


def run_configs(torch, dist, rank, world, dev, steps, flush, peaks, which=("3", "4", "5")):
    out = {}
    for key, fn in (("3", config3), ("4", config4), ("5", config5)):
        if key not in which:
            continue
        t0 = time.perf_counter()
        try:
            r = fn(torch, dist, rank, world, dev, steps, flush, peaks)
            r["bench_wall_s"] = time.perf_counter() - t0
        except Exception as exc:   # noqa: BLE001  -- the headline line must survive a failing extra
            r = {"error": repr(exc)}
        out["config" + key] = r
        torch.cuda.empty_cache()
    return out
