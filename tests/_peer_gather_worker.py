"""torchrun worker of tests/test_gpu_multi.py: every rank solves its block of load cases with the outputs placed in
rank 0's memory (parallel.PeerGather); rank 0 compares what arrived with a local solve of the whole batch."""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from python_stable_3d_truss_analysis_b200 import parallel
from python_stable_3d_truss_analysis_b200.truss import Truss

rank, world, local = parallel.init("nccl")
dev = torch.device("cuda", local)
t = Truss(3).LoadFromJSON(os.path.join(ROOT, "tests", "golden", "ref_data", "bar-942_input_0.json"))
xyz, support, conn, aed, _ = t._pack(); plan = t._get_plan(support, conn)
N, M, B = plan.N, plan.M, 48
F_all = np.random.default_rng(11).uniform(-10, 10, size=(B * world, N))
td = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)

def views(flat, nb):
    return {"u": flat[:nb * N].view(nb, N), "ext": flat[nb * N:2 * nb * N].view(nb, N), "axial": flat[2 * nb * N:nb * (2 * N + M)].view(nb, M),
            "weight": torch.empty(nb, dtype=torch.float64, device=dev), "info": torch.empty(nb, dtype=torch.int32, device=dev)}

pg = parallel.PeerGather(B * (2 * N + M), torch.float64, dev, dst=0)
lo, hi = parallel.shard_range(B * world, rank, world)
out = views(pg.local, B)
plan.solve_device(B, td(xyz), td(F_all[lo:hi]), aed=td(aed), out=out)
pg.barrier()
torch.cuda.synchronize()
assert not bool(out["info"].any().item())
if rank == 0:
    full = torch.empty(B * world * (2 * N + M), dtype=torch.float64, device=dev)
    ref = views(full, B * world)
    plan.solve_device(B * world, td(xyz), td(F_all), aed=td(aed), out=ref)
    torch.cuda.synchronize()
    for r in range(world):
        got = views(pg.slices[r], B)
        for k in ("u", "ext", "axial"):
            assert torch.equal(got[k], ref[k][r * B:(r + 1) * B]), (r, k)
    print("peer gather ok", world)
dist.barrier()
dist.destroy_process_group()
