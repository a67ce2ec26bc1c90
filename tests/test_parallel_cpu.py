"""Host-side multi-GPU logic on CPU: world_size 2 over gloo.  The solver itself needs a GPU, so the
per-rank "solve" here is a deterministic stand-in; what is tested is the partition, the ragged gather and the ordering."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from python_stable_3d_truss_analysis_b200 import parallel as par


def test_shard_range_partitions_exactly():
    for n in (0, 1, 2, 7, 8, 1024, 8192, 65536, 65537):
        for ws in (1, 2, 3, 4, 8):
            spans = [par.shard_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
            assert sizes == par.shard_sizes(n, ws)
    with pytest.raises(ValueError):
        par.shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_solve(F):
    """Stand-in for a per-rank batched solve: rows in, dict of row-aligned arrays out."""
    F = np.asarray(F)
    return {"u": F * 2.0, "axial": F[:, :3].sum(axis=1, keepdims=True) * np.ones((1, 5)), "weight": F[:, 0].copy(),
            "info": (F[:, 0] > 0).astype(np.int32)}


def _worker(rank, ws, port, n_total, ret):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(ws), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    par.init("gloo")
    rng = np.random.default_rng(0)
    F = rng.standard_normal((n_total, 6))
    got = par.sharded_call(_fake_solve, n_total, F)
    if rank == 0:
        want = _fake_solve(F)
        ok = all(np.array_equal(got[k].numpy(), want[k]) for k in want)
        ret.put(bool(ok) and set(got) == set(want))
    else:
        ret.put(got is None)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 7, 1])
def test_sharded_call_gathers_in_order_world2(n_total):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, ret)) for r in range(2)]
    for p in procs:
        p.start()
    res = [ret.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [True, True]


def test_single_process_is_passthrough():
    F = np.arange(12.0).reshape(4, 3)
    got = par.sharded_call(lambda a: {"u": a + 1}, 4, F)
    assert torch.equal(got["u"], torch.from_numpy(F + 1))
