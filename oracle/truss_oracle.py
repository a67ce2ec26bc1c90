"""numpy restatement of the reference's direct-stiffness solve (TEST INFRASTRUCTURE ONLY).

This file restates, on plain arrays, what ``slientruss3d`` v2.0.3 computes in
``Truss.Solve()`` and the GA fitness.  It exists to check the CUDA path and to
serve as the timed CPU baseline ("port") in ``bench.py``; the product never
imports it.  Every function cites the reference lines it follows
(paths relative to the reference root, e.g. ``slientruss3d/truss.py:329-364``).

Parity pin: ``tests/test_oracle_golden.py`` checks this oracle against every
golden file the reference ships (``data/bar-*_output_*.json``,
``generate/cube-7_case_*.json``, copied to ``tests/golden/ref_data``) and
against vectors produced by importing the live reference in the build
container (``tests/golden/make_golden.py`` -> ``tests/golden/live_*.json``).

Array conventions (shared with the C ABI in ``include/truss_b200.h``):
    dim      2 or 3
    joints   float64 [nJ, dim]
    support  uint8   [nJ]      SupportType ints: NO=0 PIN=1 ROLLER_X=2 ROLLER_Y=3 ROLLER_Z=4
    conn     int32   [M, 2]    (joint0, joint1) per member, member id = row
    aed      float64 [M, 3]    (a, e, density) per member
    force    float64 [N]       dense load vector, N = dim*nJ, index j*dim+axis
"""
from __future__ import annotations

import numpy as np

ZERO_EPS = 1e-10  # utils.py:79-84 (IsZero / IsZeroVector default eps)

NO, PIN, ROLLER_X, ROLLER_Y, ROLLER_Z = 0, 1, 2, 3, 4


class NotStable(Exception):
    """Counting rule of truss.py:158-164 failed (reference raises TrussNotStableError)."""


# --------------------------------------------------------------------------- supports
def resistance_number(support_type: int, dim: int) -> int:
    """type.py:37-46 -- PIN resists ``dim`` DOFs, a roller one, NO none."""
    if support_type == PIN:
        return dim
    if support_type in (ROLLER_X, ROLLER_Y, ROLLER_Z):
        return 1
    if support_type == NO:
        return 0
    raise ValueError(f"no such support type [{support_type}]")


def resistance_mask(support_type: int, dim: int) -> np.ndarray:
    """type.py:48-74 -- per-axis "is restrained" flags of one joint.

    ROLLER_Z is invalid in 2D (the reference raises there); an unknown type in
    3D silently yields None in the reference (type.py:62) -- we raise for both.
    """
    m = np.zeros(dim, dtype=bool)
    if support_type == PIN:
        m[:] = True
    elif support_type == ROLLER_X:
        m[0] = True
    elif support_type == ROLLER_Y:
        m[1] = True
    elif support_type == ROLLER_Z and dim == 3:
        m[2] = True
    elif support_type != NO:
        raise ValueError(f"no such {dim}D support type [{support_type}]")
    return m


def free_mask(dim: int, support: np.ndarray) -> np.ndarray:
    """truss.py:319-326 -- bool[N], True where the displacement is unknown."""
    n_joint = len(support)
    mask = np.ones(n_joint * dim, dtype=bool)
    for j in range(n_joint):
        mask[j * dim:(j + 1) * dim] = ~resistance_mask(int(support[j]), dim)
    return mask


def dof_maps(dim: int, support: np.ndarray):
    """The "bit-exact DOF map": boolean-mask indexing (truss.py:343,348) orders the
    free DOFs -- and the supported ones -- by ascending DOF index.

    Returns (free_idx int32[n], dof2free int32[N] (-1 at supported), sup_idx int32[s]).
    """
    mask = free_mask(dim, support)
    free_idx = np.nonzero(mask)[0].astype(np.int32)
    sup_idx = np.nonzero(~mask)[0].astype(np.int32)
    dof2free = np.full(mask.shape[0], -1, dtype=np.int32)
    dof2free[free_idx] = np.arange(free_idx.shape[0], dtype=np.int32)
    return free_idx, dof2free, sup_idx


def is_stable(dim: int, support: np.ndarray, n_member: int) -> bool:
    """truss.py:154-164 -- the counting rule (not a rank test)."""
    n_res = sum(resistance_number(int(s), dim) for s in support)
    n_joint = len(support)
    if dim == 2:
        return n_member + n_res >= n_joint * dim
    return n_res >= 6 and (n_member + n_res >= n_joint * dim)


# --------------------------------------------------------------------------- elements
def member_length(x0, x1) -> float:
    """truss.py:19,98 -- builtin sum() of the squares in axis order, then ** 0.5.
    (builtin sum is Neumaier-compensated for floats since CPython 3.12, which moves
    3-term lengths by an ulp versus a plain loop; we follow whatever the running
    interpreter does, exactly like the reference.)"""
    return sum((float(x1[i]) - float(x0[i])) ** 2. for i in range(len(x0))) ** 0.5


def member_k(a: float, e: float, length: float) -> float:
    """truss.py:56-58 -- (e*a)/L."""
    return e * a / length


def member_cosines(x0, x1, length: float):
    """truss.py:60-63."""
    return [(float(x1[i]) - float(x0[i])) / length for i in range(len(x0))]


def member_matK(x0, x1, a: float, e: float) -> np.ndarray:
    """truss.py:65-86 -- k * [[cc^T, -cc^T], [-cc^T, cc^T]]; the cosine products are
    formed first, negated where needed, and only then scaled by k."""
    length = member_length(x0, x1)
    c = member_cosines(x0, x1, length)
    d = len(c)
    cc = np.empty((d, d))
    for i in range(d):
        for j in range(d):
            cc[i, j] = c[i] ** 2. if i == j else c[i] * c[j]
    blk = np.empty((2 * d, 2 * d))
    blk[:d, :d] = cc
    blk[:d, d:] = -cc
    blk[d:, :d] = -cc
    blk[d:, d:] = cc
    return member_k(a, e, length) * blk


def assemble_K(dim, joints, conn, aed) -> np.ndarray:
    """truss.py:307-316 -- dense N x N, members in ascending id, four d x d block adds."""
    n_joint = joints.shape[0]
    K = np.zeros((n_joint * dim, n_joint * dim))
    for m in range(conn.shape[0]):
        j0, j1 = int(conn[m, 0]), int(conn[m, 1])
        ke = member_matK(joints[j0], joints[j1], float(aed[m, 0]), float(aed[m, 1]))
        for i, x in ((0, j0 * dim), (dim, j1 * dim)):
            for j, y in ((0, j0 * dim), (dim, j1 * dim)):
                K[x:x + dim, y:y + dim] += ke[i:i + dim, j:j + dim]
    return K


def weight(joints, conn, aed) -> float:
    """truss.py:52-54,166-168 -- sum over members of (a*L)*density, member order."""
    return float(sum(float(aed[m, 0]) * member_length(joints[int(conn[m, 0])], joints[int(conn[m, 1])]) * float(aed[m, 2])
                     for m in range(conn.shape[0])))


# --------------------------------------------------------------------------- solve
def solve(dim, joints, support, conn, aed, force, check_stable=True):
    """truss.py:329-364 restated on dense arrays.

    Returns dict(u[N], ext[N], axial[M], weight) with *dense* vectors; the
    reference's sparse dicts are ``sparse_*`` of these (1e-10 filter).
    ``ext`` follows truss.py:347-349: the load vector with every supported DOF
    overwritten by K[sup,:] @ u.  ``axial`` follows truss.py:353-359: the force
    the member applies at joint1, signed + for tension.
    """
    joints = np.asarray(joints, dtype=np.float64)
    conn = np.asarray(conn)
    aed = np.asarray(aed, dtype=np.float64)
    if check_stable and not is_stable(dim, support, conn.shape[0]):
        raise NotStable("The truss is not stable !")
    K = assemble_K(dim, joints, conn, aed)
    f = np.array(force, dtype=np.float64).reshape(-1).copy()
    mask = free_mask(dim, support)

    u = np.zeros(joints.shape[0] * dim)
    u[mask] = np.linalg.solve(K[mask, :][:, mask], f[mask])          # truss.py:342-343

    sup = ~mask
    ext = f
    ext[sup] = (K[sup, :] @ u.reshape(-1, 1)).ravel()                 # truss.py:348-349

    axial = np.zeros(conn.shape[0])
    for m in range(conn.shape[0]):                                    # truss.py:354-359
        j0, j1 = int(conn[m, 0]), int(conn[m, 1])
        idx = list(range(j0 * dim, (j0 + 1) * dim)) + list(range(j1 * dim, (j1 + 1) * dim))
        ke = member_matK(joints[j0], joints[j1], float(aed[m, 0]), float(aed[m, 1]))
        v = (ke[dim:] @ u[idx].reshape(-1, 1)).ravel()
        tension = np.dot(joints[j1] - joints[j0], v) > 0              # truss.py:89-91
        axial[m] = (1. if tension else -1.) * float((v ** 2).sum() ** 0.5)
    return {"u": u, "ext": ext, "axial": axial, "weight": weight(joints, conn, aed)}


def solve_closed_form(dim, joints, support, conn, aed, force):
    """Vectorised equivalent: Ke = k g g^T with g = [c, -c]
    as seen from joint1, N_m = k c.(u_j1 - u_j0), reactions summed member-wise.
    Used as a fast cross-check at sizes where the per-member Python loops of
    ``solve`` take too long; agrees with ``solve`` to rounding."""
    joints = np.asarray(joints, dtype=np.float64)
    conn = np.asarray(conn).astype(np.int64)
    aed = np.asarray(aed, dtype=np.float64)
    n_joint, M = joints.shape[0], conn.shape[0]
    N = n_joint * dim
    dx = joints[conn[:, 1]] - joints[conn[:, 0]]
    L = np.sqrt((dx ** 2).sum(axis=1))
    c = dx / L[:, None]
    k = aed[:, 1] * aed[:, 0] / L
    g = np.concatenate([-c, c], axis=1)                               # [M, 2d]
    ke = k[:, None, None] * (g[:, :, None] * g[:, None, :])
    dofs = (conn[:, :, None] * dim + np.arange(dim)[None, None, :]).reshape(M, 2 * dim)
    K = np.zeros((N, N))
    np.add.at(K, (dofs[:, :, None], dofs[:, None, :]), ke)
    mask = free_mask(dim, support)
    f = np.array(force, dtype=np.float64).reshape(-1).copy()
    u = np.zeros(N)
    u[mask] = np.linalg.solve(K[np.ix_(mask, mask)], f[mask])
    du = u.reshape(n_joint, dim)[conn[:, 1]] - u.reshape(n_joint, dim)[conn[:, 0]]
    axial = k * (c * du).sum(axis=1)
    ext = f
    ext[~mask] = (K[~mask, :] @ u)
    w = float((aed[:, 0] * L * aed[:, 2]).sum())
    return {"u": u, "ext": ext, "axial": axial, "weight": w}


# --------------------------------------------------------------------------- sparse views
def sparse_joint_dict(vec, dim):
    """truss.py:344-345 / 350-351 -- keep joints with any |component| >= 1e-10."""
    v = np.asarray(vec).reshape(-1, dim)
    return {j: v[j].copy() for j in range(v.shape[0]) if not (np.abs(v[j]) < ZERO_EPS).all()}


def sparse_member_dict(axial):
    """truss.py:358-359 -- keep members with |N| >= 1e-10."""
    return {m: float(x) for m, x in enumerate(axial) if not abs(x) < ZERO_EPS}


# --------------------------------------------------------------------------- GA fitness
def stress_violation(axial, area, limit):
    """truss.py:429-433 with isGetSumViolation=True: (allowed, sum of (sigma-limit))."""
    v = sum(s - limit for m, n_m in sparse_member_dict(axial).items() if (s := abs(n_m) / float(area[m])) > limit)
    return abs(v) < ZERO_EPS, float(v)


def displacement_violation(u, dim, limit):
    """truss.py:447-451 with isGetSumViolation=True: (allowed, sum of (|d_j|-limit))."""
    v = sum(l - limit for d in sparse_joint_dict(u, dim).values() if (l := float((d ** 2).sum() ** 0.5)) > limit)
    return abs(v) < ZERO_EPS, float(v)


def fitness(dim, joints, support, conn, gene, type_table, force, allow_stress, allow_displace):
    """ga.py:132-149 -- gene -> member types -> Solve -> penalised weight.

    Returns (fitness, stress_ok, displace_ok)."""
    type_table = np.asarray(type_table, dtype=np.float64)
    aed = type_table[np.asarray(gene, dtype=np.int64)]
    r = solve(dim, joints, support, conn, aed, force)
    ok_s, vio_s = stress_violation(r["axial"], aed[:, 0], allow_stress)
    ok_d, vio_d = displacement_violation(r["u"], dim, allow_displace)
    fit = r["weight"]
    if not ok_s:
        fit += vio_s / allow_stress * 1e5
    if not ok_d:
        fit += vio_d / allow_displace * 1e5
    return fit, bool(ok_s), bool(ok_d)


# --------------------------------------------------------------------------- JSON helpers
SUPPORT_NAMES = {"NO": NO, "PIN": PIN, "ROLLER_X": ROLLER_X, "ROLLER_Y": ROLLER_Y, "ROLLER_Z": ROLLER_Z}


def arrays_from_json(data: dict, dim: int):
    """truss.py:401-413 + 174-187 on arrays: the reference JSON layout
    (detail/combine_with_JSON.md:71-163) -> (joints, support, conn, aed, force).
    Zero force vectors are dropped by AddExternalForce (truss.py:181-182), which
    on a dense vector is a no-op."""
    joints = np.array([[float(v) for v in j[0][:dim]] for j in data["joint"]], dtype=np.float64).reshape(-1, dim)
    support = np.array([SUPPORT_NAMES[j[1]] for j in data["joint"]], dtype=np.uint8)
    conn = np.array([m[0] for m in data["member"]], dtype=np.int32).reshape(-1, 2)
    aed = np.array([m[1] for m in data["member"]], dtype=np.float64).reshape(-1, 3)
    force = np.zeros(joints.shape[0] * dim)
    for jid, vec in data["force"]:
        force[jid * dim:(jid + 1) * dim] = [float(v) for v in vec[:dim]]
    return joints, support, conn, aed, force


def dense_from_sparse(pairs, n, dim=None):
    """Densify a reference ``[[id, value], ...]`` result list."""
    if dim is None:
        out = np.zeros(n)
        for i, v in pairs:
            out[i] = v
        return out
    out = np.zeros((n, dim))
    for i, v in pairs:
        out[i] = v
    return out.reshape(-1)


def normwise_err(a, b) -> float:
    """SURVEY.md section 4 trap 1: compare densified vectors with max|a-b| / max|b|."""
    a, b = np.asarray(a, dtype=np.float64).ravel(), np.asarray(b, dtype=np.float64).ravel()
    scale = np.abs(b).max() if b.size else 0.0
    if scale == 0.0:
        return float(np.abs(a).max()) if a.size else 0.0
    return float(np.abs(a - b).max() / scale)
