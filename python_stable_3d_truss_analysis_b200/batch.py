"""Batched entry points next to the reference's one-truss-at-a-time API.

The reference solves exactly one truss per ``Truss.Solve()`` call and its callers loop in Python
(``ga.py:155-160`` over genes, ``generate.py:342-357`` over generated trusses).  These functions
hand the whole batch to the CUDA library in one call:

  ``SolveBatch(trusses)``                    any list of Truss objects (shared or ragged topology)
  ``SolveLoadCases(truss, forces)``          one truss, many dense load vectors (same K, many f)
  ``SolveMemberTypes(truss, genes, types)``  one truss, many member-type assignments (GA population)
  ``FitnessBatch(...)``                      GA.GetFitness for a population (ga.py:139-149)

All of them return dense numpy arrays (``u [B,N]``, ``ext [B,N]``, ``axial [B,M]``, ``weight [B]``,
``info [B]``); ``SolveBatch`` also stores the results back into the Truss objects.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .truss import Truss, raise_for_info
from .type import MemberType


def type_table(memberTypeList):
    """[T,3] float64 (a, e, density) table from MemberType objects or triples."""
    return np.array([t.Serialize() if isinstance(t, MemberType) else list(t) for t in memberTypeList],
                    dtype=np.float64).reshape(-1, 3)


def _check_infos(info, raise_on_error):
    if raise_on_error:
        bad = np.nonzero(info)[0]
        if bad.size:
            raise_for_info(int(info[bad[0]]))


def SolveLoadCases(truss: Truss, forces, raise_on_error=True, independent=False):
    """One truss under B load cases.  ``forces`` is [B, N] dense (index jointID*dim + axis).

    Equivalent to ``for f in forces: truss.SetForces(f); truss.Solve()`` (truss.py:329-364).  The stiffness matrix is the
    same for every load case, so on the band path it is assembled and factorised once (tb_solve_loadcases_host);
    ``independent=True`` factorises every system on its own, as a batch of unrelated trusses would (tb_solve_host)."""
    xyz, support, conn, aed, _ = truss._pack()
    forces = np.ascontiguousarray(forces, dtype=np.float64).reshape(-1, truss.nJoint * truss.dim)
    plan = truss._get_plan(support, conn)
    out = plan.solve_host(forces.shape[0], xyz, forces, aed=aed, shared_factor=not independent)
    _check_infos(out["info"], raise_on_error)
    return out


def SolveMemberTypes(truss: Truss, genes, memberTypeList, raise_on_error=True):
    """One truss under B member-type assignments; ``genes`` is int [B, M] of indices into the list."""
    xyz, support, conn, _, force = truss._pack()
    genes = np.ascontiguousarray(genes, dtype=np.int32).reshape(-1, truss.nMember)
    plan = truss._get_plan(support, conn)
    out = plan.solve_host(genes.shape[0], xyz, force, gene=genes, type_table=type_table(memberTypeList))
    _check_infos(out["info"], raise_on_error)
    return out


def FitnessBatch(truss: Truss, genes, memberTypeList, allowStress, allowDisplace, full=False):
    """GA.GetFitness (ga.py:139-149) for a whole population in one call.

    Returns dict(fitness [B], flags [B,2] (stress ok, displacement ok), info [B]) (+ full results)."""
    xyz, support, conn, _, force = truss._pack()
    genes = np.ascontiguousarray(genes, dtype=np.int32).reshape(-1, truss.nMember)
    plan = truss._get_plan(support, conn)
    return plan.fitness_host(genes.shape[0], xyz, force, genes, type_table(memberTypeList), allowStress,
                             allowDisplace, full=full)


def pack_ragged(trusses):
    """Pack Truss objects back to back in the tb_ragged_in layout."""
    dim = trusses[0].dim
    packs = [t._pack() for t in trusses]
    joint_off = np.zeros(len(trusses) + 1, np.int64)
    member_off = np.zeros(len(trusses) + 1, np.int64)
    joint_off[1:] = np.cumsum([p[0].shape[0] for p in packs])
    member_off[1:] = np.cumsum([p[2].shape[0] for p in packs])
    cat = lambda i, dt: np.concatenate([np.asarray(p[i]).reshape(-1) for p in packs]).astype(dt)  # noqa: E731
    return dim, joint_off, member_off, cat(0, np.float64), cat(1, np.uint8), cat(2, np.int32), cat(3, np.float64), \
        cat(4, np.float64)


def SolveBatch(trusses, raise_on_error=True):
    """Solve a list of Truss objects in one call and store the results into them.

    Same topology everywhere -> one plan, uniform batch.  Otherwise, if every truss fits the fused
    shared-memory kernel, one ragged batch (the generator's case); else grouped by topology.
    Returns the per-truss info codes."""
    trusses = list(trusses)
    if not trusses:
        return np.zeros(0, np.int32)
    dim = trusses[0].dim
    if any(t.dim != dim for t in trusses):
        raise ValueError("all trusses of a batch must have the same dimension")
    packs = [t._pack() for t in trusses]
    info = np.zeros(len(trusses), np.int32)

    def store(idx, u, ext, axial, inf):
        for k, i in enumerate(idx):
            info[i] = inf[k]
            if inf[k] == 0:
                trusses[i]._set_dense_results(u[k], ext[k], axial[k])

    def signature(p):
        return (p[0].shape[0], p[2].shape[0], p[1].tobytes(), p[2].tobytes())

    groups = {}
    for i, p in enumerate(packs):
        groups.setdefault(signature(p), []).append(i)

    # one ragged batch needs the LARGEST truss of the batch to fit the fused kernels' shared memory
    fits_small = _lib.small_path_fits(dim, max(p[0].shape[0] for p in packs), max(p[2].shape[0] for p in packs))
    if len(groups) > 1 and fits_small:
        _, jo, mo, xyz, sup, conn, aed, force = pack_ragged(trusses)
        out = _lib.solve_ragged_host(dim, jo, mo, xyz, sup, conn, aed, force, want=("u", "ext", "axial"))
        for i in range(len(trusses)):
            info[i] = out["info"][i]
            if info[i] == 0:
                trusses[i]._set_dense_results(out["u"][jo[i] * dim:jo[i + 1] * dim], out["ext"][jo[i] * dim:jo[i + 1] * dim],
                                              out["axial"][mo[i]:mo[i + 1]])
    else:
        for idx in groups.values():
            t0, p0 = trusses[idx[0]], packs[idx[0]]
            plan = t0._get_plan(p0[1], p0[2])
            B = len(idx)
            xyz = np.stack([packs[i][0] for i in idx])
            aed = np.stack([packs[i][3] for i in idx])
            force = np.stack([packs[i][4] for i in idx])
            out = plan.solve_host(B, xyz, force, aed=aed, want=("u", "ext", "axial"))
            store(idx, out["u"], out["ext"], out["axial"], out["info"])
    _check_infos(info, raise_on_error)
    return info
