"""Random cube-truss topologies generated on the device (csrc/tb_gencube.cu, generate.GenerateRandomCubeTrussesOnDevice)
against the reference generator's definitions (slientruss3d/generate.py:152-336, 338-372).

The kernel cannot replay Python's ``random`` stream, so it exports the random choices of every truss (the cell sequence of
the walk, the diagonal pick of every face, the three cell lengths); the deterministic rest -- joint numbering, member
linking with the duplicate filter, positions, pin supports -- is replayed through the host classes CubeTruss / CubeGrid
(which are RNG-compatible with the reference: seed 42 reproduces its shipped cube-7 files) and must agree exactly.  The
random choices themselves are checked against the rules of the walk and for their distributions; solved results against
the oracle."""
import numpy as np
import pytest

from oracle import truss_oracle as orc
from python_stable_3d_truss_analysis_b200 import generate as G
from python_stable_3d_truss_analysis_b200.dataset import PackedDataset
from python_stable_3d_truss_analysis_b200.type import GenerateMethod, LinkType

pytestmark = pytest.mark.gpu

GRID = (4, 3, 3)
TYPES = ((1., 1e7, 0.1), (2.5, 2e7, 0.2), (0.5, 1e7, 0.3))


class _Picks:
    """stands in for the ``random`` module inside LinkMember: hands out the exported diagonal picks"""

    def __init__(self, picks):
        self.it = iter(picks)

    def sample(self, population, k):
        assert len(population) == 3 and k == 1
        return [int(next(self.it))]


def _replay(ds, o, allow_parallel, link_type, add_pin=True):
    gx, gy, gz = GRID
    cells = [int(c) for c in ds["cells"][o] if c >= 0]
    coords = [(c % gx, (c // gx) % gy, c // (gx * gy)) for c in cells]
    grid = G.CubeGrid(*GRID)
    cubes = []
    for coord in coords:
        grid[coord] = True
        cubes.append(G.CubeTruss(coord, grid._usedDict))
    real = G.random
    G.random = _Picks(ds["picks"][o][:len(cells)].reshape(-1))
    try:
        data = grid.CubesToTruss(cubes, list(ds["length"][o]), add_pin, allow_parallel, link_type)
    finally:
        G.random = real
    return coords, data


def _slice(ds, o):
    j0, j1, m0, m1 = ds["joint_off"][o], ds["joint_off"][o + 1], ds["member_off"][o], ds["member_off"][o + 1]
    return (ds["xyz"][3 * j0:3 * j1].reshape(-1, 3), ds["support"][j0:j1], ds["conn"][2 * m0:2 * m1].reshape(-1, 2),
            ds["aed"][3 * m0:3 * m1].reshape(-1, 3), ds["force"][3 * j0:3 * j1].reshape(-1, 3))


@pytest.mark.parametrize("method", [GenerateMethod.DFS, GenerateMethod.BFS, GenerateMethod.Random])
@pytest.mark.parametrize("allow_parallel,link_type", [(False, LinkType.Random), (True, LinkType.Random), (False, LinkType.Cross),
                                                      (False, LinkType.LeftBottom_RightTop)])
def test_generated_topologies_replay_through_the_host_classes(method, allow_parallel, link_type):
    n = 300
    ds = G.GenerateRandomCubeTrussesOnDevice(n, GRID, (1, 9), (50, 150), nForceRange=None, method=method, linkType=link_type,
                                             memberTypes=TYPES, isAllowParallel=allow_parallel, seed=7 + method, export=True)
    assert ds["joint_off"][-1] * 3 == ds["xyz"].size and not ds["gen_info"].any()
    table = np.array(TYPES)
    for o in range(0, n, 3):
        coords, data = _replay(ds, o, allow_parallel, link_type)
        xyz, sup, conn, aed, force = _slice(ds, o)
        # --- the walk: distinct cells inside the grid, every cell after the first a face neighbour of an earlier one
        assert len(set(coords)) == len(coords) and 1 <= len(coords) <= 9
        for i, c in enumerate(coords[1:], 1):
            assert any(sum(abs(a - b) for a, b in zip(c, p)) == 1 for p in coords[:i]), (o, coords)
        # --- joints, supports and members: exactly what the host classes build from the same choices
        assert xyz.shape[0] == len(data["joint"]) and conn.shape[0] == len(data["member"])
        assert np.array_equal(xyz, np.array([j[0] for j in data["joint"]]))
        assert [("PIN" if s else "NO") for s in sup] == [j[1] for j in data["joint"]]
        assert conn.tolist() == [m[0] for m in data["member"]]
        # --- loads on unsupported joints only, inside the ranges; member types from the table; the counting rule
        loaded = np.nonzero(force.any(axis=1))[0]
        assert 1 <= len(loaded) <= int((sup == 0).sum()) and not sup[loaded].any()
        assert np.all(np.abs(force) <= 30000)
        assert all(any(np.array_equal(row, t) for t in table) for row in aed)
        assert conn.shape[0] + 3 * int(sup.sum()) >= 3 * xyz.shape[0]
        if link_type != LinkType.Random:
            assert set(ds["picks"][o][:len(coords)].reshape(-1).tolist()) == {int(link_type)}


def test_random_choices_are_uniform_and_reproducible():
    n = 6000
    kw = dict(gridRange=GRID, numCubeRange=(2, 7), memberTypes=TYPES, export=True)
    a = G.GenerateRandomCubeTrussesOnDevice(n, seed=11, **kw)
    b = G.GenerateRandomCubeTrussesOnDevice(n, seed=11, **kw)
    c = G.GenerateRandomCubeTrussesOnDevice(n, seed=12, **kw)
    for k in ("xyz", "conn", "force", "aed", "cells", "picks"):
        assert np.array_equal(a[k], b[k]), k
    assert not np.array_equal(a["cells"], c["cells"])
    ok = a["gen_info"] == 0
    first = a["cells"][ok, 0]
    cnt = np.bincount(first, minlength=36)
    assert cnt.min() > 0.6 * ok.sum() / 36 and cnt.max() < 1.4 * ok.sum() / 36          # GetRandomFeasible: uniform over the cells
    ncube = (a["cells"] >= 0).sum(axis=1)
    assert set(ncube.tolist()) <= set(range(1, 8)) and np.bincount(ncube[ok], minlength=8)[2:8].min() > 0.05 * ok.sum()
    picks = a["picks"][ok, 0, :].reshape(-1)
    pc = np.bincount(picks, minlength=3) / picks.size
    assert np.all(np.abs(pc - 1 / 3) < 0.02)                                                # random.sample(range(3), 1)
    lens = a["length"][ok]
    assert lens.min() >= 50 and lens.max() <= 150 and abs(lens.mean() - 100) < 2
    types = a["aed"].reshape(-1, 3)[:, 0]
    frac = np.array([(types == t[0]).mean() for t in TYPES])
    assert np.all(np.abs(frac - 1 / 3) < 0.02)                                              # random.choice(memberTypes)


def test_generated_and_solved_in_one_pass_matches_the_oracle():
    n = 2000
    ds = G.GenerateRandomCubeTrussesOnDevice(n, (5, 5, 5), (5, 5), memberTypes=TYPES, isDoStructuralAnalysis=True, seed=3)
    assert (ds["gen_info"] == 0).all()
    solved = np.nonzero(ds["info"] == 0)[0]
    assert solved.size > 0.9 * n                          # (the counting rule is necessary, not sufficient: a few are singular)
    pk = PackedDataset(3, ds)
    for o in solved[:: max(1, solved.size // 12)]:
        xyz, sup, conn, aed, force = _slice(ds, o)
        try:
            want = orc.solve(3, xyz, sup, conn, aed, force)
        except np.linalg.LinAlgError:
            continue
        j0, j1, m0, m1 = ds["joint_off"][o], ds["joint_off"][o + 1], ds["member_off"][o], ds["member_off"][o + 1]
        if np.abs(want["u"]).max() > 1e6:                 # a mechanism the counting rule lets through: ill-posed, skip
            continue
        assert orc.normwise_err(ds["u"][3 * j0:3 * j1], want["u"]) <= 1e-7
        assert orc.normwise_err(ds["axial"][m0:m1], want["axial"]) <= 1e-7
        t = pk.truss(int(o))
        assert t.isSolved and t.nJoint == j1 - j0 and t.nMember == m1 - m0
