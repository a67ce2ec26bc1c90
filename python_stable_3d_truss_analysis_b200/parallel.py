"""Multi-GPU plumbing: one process per GPU, the batch sharded by contiguous blocks, results gathered to rank 0.

Every truss / gene / load case is an independent linear system, so there is no collective on
the data path: each rank solves rows ``shard_range(B, rank, world)`` of the batch with its own copy of the plan, and
one gather (NCCL over NVLink on GPUs, gloo in the CPU tests) brings ``u / ext / axial / weight / info`` or the GA
fitness back to rank 0.  The reference has no counterpart (its GA loop is sequential, ga.py:139-160).
"""
from __future__ import annotations

import os

import numpy as np


def world():
    """(rank, world_size, local_rank) from the torchrun environment (1 process = (0, 1, 0))."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(n_total: int, rank: int, world_size: int):
    """Contiguous block partition: the first ``n_total % world_size`` ranks get one extra row."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world size")
    base, extra = divmod(int(n_total), world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_total: int, world_size: int):
    return [shard_range(n_total, r, world_size)[1] - shard_range(n_total, r, world_size)[0] for r in range(world_size)]


def init(backend: str | None = None):
    """Initialise torch.distributed from the torchrun environment (NCCL with a GPU, gloo without)."""
    import torch
    import torch.distributed as dist

    rank, ws, local = world()
    if ws == 1 or dist.is_initialized():
        return rank, ws, local
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group(backend)
    return rank, ws, local


def gather_rows(local_rows, n_total: int, dst: int = 0):
    """Gather the row shards of a [n_total, ...] array to rank ``dst`` (tensor in, tensor out; None elsewhere).

    Shards may differ by one row, so every rank pads to the largest shard and rank ``dst`` trims while concatenating."""
    import torch
    import torch.distributed as dist

    t = local_rows if isinstance(local_rows, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local_rows))
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return t
    rank, ws = dist.get_rank(), dist.get_world_size()
    sizes = shard_sizes(n_total, ws)
    if t.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {t.shape[0]} rows, its shard has {sizes[rank]}")
    width = max(sizes)
    if t.shape[0] < width:
        pad = torch.zeros((width - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        t = torch.cat([t, pad], dim=0)
    t = t.contiguous()
    bufs = [torch.empty_like(t) for _ in range(ws)] if rank == dst else None
    dist.gather(t, bufs, dst=dst)
    if rank != dst:
        return None
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)


def gather_results(local: dict, n_total: int, dst: int = 0):
    """Gather a dict of row-sharded arrays (u, ext, axial, weight, info, fitness, flags ...) to rank ``dst``."""
    out = {k: gather_rows(v, n_total, dst) for k, v in sorted(local.items()) if v is not None}
    rank = world()[0]
    return out if rank == dst else None


def sharded_call(fn, n_total: int, *row_arrays, dst: int = 0, **kwargs):
    """Run ``fn(*shards, **kwargs) -> dict`` on this rank's rows of every array in ``row_arrays`` and gather the dict.

    ``fn`` is e.g. ``lambda F: SolveLoadCases(truss, F, raise_on_error=False)`` or a FitnessBatch closure."""
    import torch.distributed as dist

    rank, ws = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
    lo, hi = shard_range(n_total, rank, ws)
    local = fn(*[a[lo:hi] for a in row_arrays], **kwargs)
    return gather_results(local, n_total, dst)


def SolveLoadCasesSharded(truss, forces, dst: int = 0):
    """batch.SolveLoadCases over all ranks: rank r solves its block of load cases, rank ``dst`` gets everything."""
    from .batch import SolveLoadCases

    forces = np.ascontiguousarray(forces, dtype=np.float64).reshape(-1, truss.nJoint * truss.dim)
    return sharded_call(lambda F: SolveLoadCases(truss, F, raise_on_error=False), forces.shape[0], forces, dst=dst)


def FitnessBatchSharded(truss, genes, memberTypeList, allowStress, allowDisplace, dst: int = 0):
    """batch.FitnessBatch (GA.GetFitness for a population, ga.py:139-149) with the population sharded over the ranks."""
    from .batch import FitnessBatch

    genes = np.ascontiguousarray(genes, dtype=np.int32).reshape(-1, truss.nMember)
    return sharded_call(lambda g: FitnessBatch(truss, g, memberTypeList, allowStress, allowDisplace), genes.shape[0], genes,
                        dst=dst)
