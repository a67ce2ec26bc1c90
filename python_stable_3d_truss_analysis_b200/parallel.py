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


class PeerGather:
    """Gather without a collective: every rank's result buffer IS a slice of one buffer in rank ``dst``'s HBM.

    The buffer is allocated as symmetric memory and mapped into every process of the node (NVLink 5 / NVSwitch peer
    access), so the stores of the recovery kernel (k_recover / the fused kernels) travel to rank ``dst`` as they are
    issued -- compute and transfer are one kernel -- and ``barrier()`` (a device-side signal barrier on the current
    stream) is all that follows.  ``local`` is this rank's slice (pass views of it as the solver's outputs);
    ``slices`` on rank ``dst`` lists every rank's slice.  Requires one process per GPU on one node (NCCL group)."""

    def __init__(self, per_rank_elems: int, dtype=None, device=None, dst: int = 0, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        dtype = torch.float64 if dtype is None else dtype
        group = dist.group.WORLD if group is None else group
        self.rank, self.world, self.dst = dist.get_rank(group), dist.get_world_size(group), dst
        self.per_rank = int(per_rank_elems)
        self._pool = symm_mem.empty(self.world * self.per_rank, dtype=dtype, device=device)
        self._hdl = symm_mem.rendezvous(self._pool, group)
        self.local = self._hdl.get_buffer(dst, (self.per_rank,), dtype, self.rank * self.per_rank)
        self.slices = ([self._pool[r * self.per_rank:(r + 1) * self.per_rank] for r in range(self.world)]
                       if self.rank == dst else None)

    def barrier(self):
        """Enqueue the device-side barrier: after it, rank ``dst`` holds every rank's results."""
        self._hdl.barrier()


class PipelinedPeerGather(PeerGather):
    """PeerGather with the transfer taken off the critical path: every rank computes into one of TWO local result buffers
    and a copy stream pushes the finished one into its slice of rank ``dst``'s buffer (peer-mapped symmetric memory, NVLink /
    NVSwitch) while the next step is already computing into the other.  With eight GPUs the 7 x 20 MB that converge on rank
    ``dst``'s NVLink ingress every step (0.15-0.2 ms at the measured 770 GB/s) then ride under the next step's
    factorisation instead of stalling the recovery kernel's stores.

        buf = pg.acquire(i, stream)      # result buffer of step i (waits, on `stream`, until step i-2's copy has left it)
        ... enqueue the solve of step i into views of buf on `stream` ...
        pg.submit(i, stream)             # copy stream: wait for the solve, push buf to rank dst
        pg.drain(stream)                 # after the last step: `stream` waits for the outstanding copies, then the barrier

    Rank ``dst`` computes straight into its own slice (nothing to copy)."""

    def __init__(self, per_rank_elems: int, dtype=None, device=None, dst: int = 0, group=None):
        import torch

        super().__init__(per_rank_elems, dtype, device, dst, group)
        dtype = torch.float64 if dtype is None else dtype
        self.is_dst = self.rank == dst
        self.bufs = [self.local, self.local] if self.is_dst else [torch.empty(self.per_rank, dtype=dtype, device=device) for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=device)
        self.ev_solved = [torch.cuda.Event() for _ in range(2)]
        self.ev_copied = [torch.cuda.Event() for _ in range(2)]
        self._pending = [False, False]

    def acquire(self, i: int, stream):
        k = i & 1
        if self._pending[k]:
            stream.wait_event(self.ev_copied[k])
        return self.bufs[k]

    def submit(self, i: int, stream):
        import torch

        k = i & 1
        if self.is_dst:
            return
        self.ev_solved[k].record(stream)
        self.copy_stream.wait_event(self.ev_solved[k])
        with torch.cuda.stream(self.copy_stream):
            self.local.copy_(self.bufs[k], non_blocking=True)
        self.ev_copied[k].record(self.copy_stream)
        self._pending[k] = True

    def drain(self, stream):
        for k in range(2):
            if self._pending[k]:
                stream.wait_event(self.ev_copied[k])
                self._pending[k] = False
        self.barrier()


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
