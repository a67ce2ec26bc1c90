"""The two-sided band program of the fused band kernel (csrc/tb_tsplan.cu), replayed in numpy (tests/ts_replay.py).

The program is pure integer data built on the host, so everything but the thread-level layout of ``k_band_ts`` can be
checked without a GPU: the assembly pass's lists in the kernel's entry order (contributions in ascending member order,
truss.py:307-316, 343), structural masks, ring slots, the hand-over
to the separator, chunk offsets, both back substitutions -- against the oracle's dense solve on every fixture, with
the automatic split and with forced ones.
"""
import os

import numpy as np
import pytest

from oracle import truss_oracle as orc
from python_stable_3d_truss_analysis_b200 import _lib
from tests import helpers as H
from tests import ts_replay


def cases():
    out = [(n, d, data) for n, d, data, _ in H.shipped_cases()]
    out += [(f"random{i}", c["dim"], c["data"]) for i, c in enumerate(H.load_json("live_random.json"))]
    out += [(n, d, data) for n, d, data, _ in H.cube7_shipped()[:2]]
    return out


def run_case(dim, data, env=None):
    joints, support, conn, aed, force = orc.arrays_from_json(data, dim)
    old = {k: os.environ.get(k) for k in (env or {})}
    os.environ.update(env or {})
    try:
        plan = _lib.Plan(dim, conn, support)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    prog = plan.ts_program()
    return plan, prog, (joints, support, conn, aed, force)


def check(plan, prog, arrays, dim):
    joints, support, conn, aed, force = arrays
    u_dof, dbg = ts_replay.replay(prog, dim, joints, conn, aed, force.reshape(-1), plan.N)
    want = orc.solve(dim, joints, support, conn, aed, force)
    err = orc.normwise_err(u_dof, want["u"])
    assert err <= 1e-9, f"replayed program vs oracle: {err:.3e}"
    # assembled K values: the program must add the contributions of every entry in the scatter map's order (ascending
    # member, truss.py:310-314) -- bit-exact against a replay of the scatter map with the device's arithmetic (squares as
    # products; the reference's ``l ** 2.`` goes through libm pow, which is an ulp off the product now and then, and
    # CPython's compensated builtin sum() moves three-term lengths by an ulp, so against the oracle's
    # GetKMatrix()[mask][:, mask] the bound is a few ulp of the contributions)
    K = orc.assemble_K(dim, joints, conn, aed)
    mask = orc.free_mask(dim, support)
    Kff = K[mask][:, mask]
    row, col, ptr, mem, loc = plan.scatter()
    src = prog["ent_src"]
    live = src >= 0                     # (holes of the entry schedule carry -1)
    assert sorted(src[live].tolist()) == list(range(len(row))), "every scatter-map entry is assembled exactly once"
    got = np.zeros(len(row))
    got[src[live]] = dbg["kv"][live]
    prods = [ts_replay.member_products(dim, joints, conn[m], aed[m][0], aed[m][1]) for m in range(conn.shape[0])]
    want_k = np.zeros(len(row))
    mag = np.zeros(len(row))
    for i in range(len(row)):
        v = 0.0
        for q in range(ptr[i], ptr[i + 1]):
            a, b = divmod(int(loc[q]), 2 * dim)
            k, c = prods[mem[q]]
            t = k * (c[a % dim] * c[b % dim])
            v = v + (-t if (a // dim) != (b // dim) else t)
            mag[i] += abs(t)
        want_k[i] = v
    assert np.array_equal(got, want_k), "program order differs from the scatter map's ascending-member order"
    ref = Kff[row, col]
    assert np.all(np.abs(got - ref) <= 16 * np.finfo(float).eps * mag), "assembled K is more than a few ulp off GetKMatrix()"


@pytest.mark.parametrize("name,dim,data", cases(), ids=lambda v: v if isinstance(v, str) else None)
def test_program_replay_matches_oracle(name, dim, data):
    plan, prog, arrays = run_case(dim, data)
    if prog is None:
        pytest.skip("band too wide for the two-sided kernel")
    check(plan, prog, arrays, dim)


def test_bar942_is_split_in_two():
    name, dim, data, _ = next(c for c in H.shipped_cases() if c[0].startswith("bar-942"))
    plan, prog, arrays = run_case(dim, data)
    info = prog["info"]
    assert info["nS"] >= 1 and info["nB"] >= 1, info
    # the dependent chain: max(top, bottom) + separator block columns, well below the one-sided count
    chain = max(info["bT"], info["nB"]) + info["nS"]
    assert chain <= 0.65 * info["nblk"], info
    assert info["nb_top"] <= 8 and info["nb_bottom"] <= 8


@pytest.mark.parametrize("split", [3, 10, 20, 40, 60, 75])
def test_forced_splits_bar942(split):
    name, dim, data, _ = next(c for c in H.shipped_cases() if c[0].startswith("bar-942"))
    plan, prog, arrays = run_case(dim, data, env={"TB_TS_SPLIT": str(split)})
    assert prog is not None
    if prog["info"]["nS"] > 0:
        assert prog["info"]["bT"] == split
    check(plan, prog, arrays, dim)


def test_one_sided_program():
    name, dim, data, _ = next(c for c in H.shipped_cases() if c[0].startswith("bar-942"))
    plan, prog, arrays = run_case(dim, data, env={"TB_TS_ONE_SIDED": "1"})
    assert prog["info"]["nS"] == 0 and prog["info"]["nB"] == 0
    check(plan, prog, arrays, dim)


@pytest.mark.parametrize("i", range(4))
def test_forced_splits_small_trusses(i):
    """Every fixture with at least four block columns, split after its first / middle block column."""
    done = 0
    for name, dim, data in cases():
        joints, support, conn, aed, force = orc.arrays_from_json(data, dim)
        n = int(orc.free_mask(dim, support).sum())
        if n < 32:
            continue
        for split in (1, max(1, (n // 8) // 2)):
            plan, prog, arrays = run_case(dim, data, env={"TB_TS_SPLIT": str(split)})
            if prog is None:
                continue
            check(plan, prog, arrays, dim)
            done += 1
        if done > 2 * (i + 1):
            break
    assert done > 0
