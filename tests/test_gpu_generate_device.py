"""Dataset generation on the device (csrc/tb_augment.cu, generate.GenerateAugmentedDataset) against the reference's augmenter
definitions (slientruss3d/generate.py:12-148) and, for the solved results, against the oracle on the augmented arrays."""
import copy

import numpy as np
import pytest

from oracle import truss_oracle as orc
from python_stable_3d_truss_analysis_b200.generate import GenerateAugmentedDataset, MoveToCentroid
from python_stable_3d_truss_analysis_b200.truss import Truss
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _pool():
    cases = H.cube7_shipped()
    return [Truss(3).LoadFromJSON(data=copy.deepcopy(data)) for _, _, data, _ in cases], cases


def _slice(ds, o):
    d = ds["dim"]
    j0, j1, m0, m1 = ds["joint_off"][o], ds["joint_off"][o + 1], ds["member_off"][o], ds["member_off"][o + 1]
    return (ds["xyz"][d * j0:d * j1].reshape(-1, d), ds["support"][j0:j1], ds["conn"][2 * m0:2 * m1].reshape(-1, 2),
            ds["aed"][3 * m0:3 * m1].reshape(-1, 3), ds["force"][d * j0:d * j1])


def test_no_augmentation_reproduces_the_pool_and_the_shipped_results():
    pool, cases = _pool()
    ds = GenerateAugmentedDataset(pool, 25)
    assert np.array_equal(ds["src"], np.arange(25) % 10) and not ds["info"].any()
    for o in range(25):
        xyz, sup, conn, aed, force = _slice(ds, o)
        pxyz, psup, pconn, paed, pforce = pool[o % 10]._pack()
        assert np.array_equal(xyz, pxyz) and np.array_equal(sup, psup) and np.array_equal(conn, pconn)
        assert np.array_equal(aed, paed) and np.array_equal(force, pforce)
        gold = cases[o % 10][3]
        j0, j1, m0, m1 = ds["joint_off"][o], ds["joint_off"][o + 1], ds["member_off"][o], ds["member_off"][o + 1]
        got = {"u": ds["u"][3 * j0:3 * j1], "ext": ds["ext"][3 * j0:3 * j1], "axial": ds["axial"][m0:m1], "weight": ds["weight"][o]}
        H.assert_close(got, gold, what=f"output {o}")


def test_centroid_is_bit_exact_and_translation_is_one_vector_per_truss():
    pool, _ = _pool()
    ds = GenerateAugmentedDataset(pool, 10, moveToCentroid=True, isDoStructuralAnalysis=False)
    for o in range(10):
        want = MoveToCentroid()(pool[o].Serialize())           # the host augmenter: the reference's arithmetic
        assert np.array_equal(_slice(ds, o)[0], np.array([j[0] for j in want["joint"]]))
    dt = GenerateAugmentedDataset(pool, 200, translateRange=(-30., 30.), seed=5, isDoStructuralAnalysis=False)
    shifts = []
    for o in range(200):
        delta = _slice(dt, o)[0] - pool[o % 10]._pack()[0]
        assert np.abs(delta - delta[0]).max() <= 1e-11 and np.all(np.abs(delta[0]) <= 30.)
        shifts.append(delta[0])
    shifts = np.array(shifts)
    assert len({tuple(np.round(s, 6)) for s in shifts}) == 200 and abs(shifts.mean()) < 4.0 and shifts.std() > 12.0
    again = GenerateAugmentedDataset(pool, 200, translateRange=(-30., 30.), seed=5, isDoStructuralAnalysis=False)
    assert np.array_equal(again["xyz"], dt["xyz"])


def test_joint_noise_statistics_and_pin_reset_rules():
    pool, _ = _pool()
    n = 3000
    ds = GenerateAugmentedDataset(pool, n, noiseMeans=[1., -2., 0.], noiseStds=[10., 5., 0.5], seed=2, isDoStructuralAnalysis=False)
    deltas = np.concatenate([_slice(ds, o)[0] - pool[o % 10]._pack()[0] for o in range(n)])          # [sum nJ, 3]
    cnt = deltas.shape[0]
    for ax, (mu, sd) in enumerate(((1., 10.), (-2., 5.), (0., 0.5))):
        assert abs(deltas[:, ax].mean() - mu) < 5 * sd / np.sqrt(cnt) and abs(deltas[:, ax].std() - sd) < 0.02 * sd
    assert abs(np.corrcoef(deltas[:, 0], deltas[:, 1])[0, 1]) < 0.02
    dp = GenerateAugmentedDataset(pool, n, resetPin=(5, 0.6), seed=3, isDoStructuralAnalysis=False)
    counts = []
    for o in range(n):
        sup = _slice(dp, o)[1]
        t = pool[o % 10]
        assert set(np.unique(sup)) <= {0, 1}
        lo = max(5, int(np.ceil((t.nJoint * 3 - t.nMember) / 3)))
        hi = int(0.6 * t.nJoint)
        k = int(sup.sum())
        assert lo <= k <= hi, (o, k, lo, hi)
        counts.append((k - lo) / max(1, hi - lo))
    assert 0.4 < np.mean(counts) < 0.6                          # k uniform over its range
    first = np.array([_slice(dp, o)[1][0] for o in range(0, n, 10)])
    assert 0.1 < first.mean() < 0.9                             # the pinned joints move around


def test_full_recipe_solved_on_the_device_matches_the_oracle():
    pool, _ = _pool()
    ds = GenerateAugmentedDataset(pool, 4096, moveToCentroid=True, translateRange=(-30., 30.), noiseStds=[10., 10., 10.],
                                  resetPin=(5, 0.6), seed=42)
    rng = np.random.default_rng(0)
    checked = 0
    for o in rng.choice(4096, size=24, replace=False):
        xyz, sup, conn, aed, force = _slice(ds, o)
        stable = orc.is_stable(3, sup, conn.shape[0])
        assert (ds["info"][o] == -1) == (not stable)
        if ds["info"][o] != 0:
            continue
        want = orc.solve(3, xyz, sup, conn, aed, force)
        j0, j1, m0, m1 = ds["joint_off"][o], ds["joint_off"][o + 1], ds["member_off"][o], ds["member_off"][o + 1]
        got = {"u": ds["u"][3 * j0:3 * j1], "ext": ds["ext"][3 * j0:3 * j1], "axial": ds["axial"][m0:m1]}
        for k in H.FIELDS:
            assert orc.normwise_err(got[k], want[k]) <= H.TOL, (o, k)
        checked += 1
    assert checked >= 12
    # the packed result goes into the binary container as it is, and comes back as Truss objects / reference JSON
    from python_stable_3d_truss_analysis_b200.dataset import PackedDataset
    pd_ = PackedDataset(3, ds)
    o = int(np.nonzero(ds["info"] == 0)[0][5])
    t = pd_.truss(o)
    j0, j1 = ds["joint_off"][o], ds["joint_off"][o + 1]
    assert t.isSolved and np.array_equal(t._dense["u"], ds["u"][3 * j0:3 * j1])
    again = Truss(3).LoadFromJSON(data=pd_.json(o), isOutputFile=True)
    assert again.GetInternalForces() == t.GetInternalForces()


def test_bulk_pipeline_json_files_to_solved_container():
    """N reference JSON files -> packed arrays (native loader) -> one ragged GPU batch -> the shipped results, without a
    single Truss object on the way (truss.py:401-421 + generate.py:354-357 in bulk)."""
    import glob
    import os
    from python_stable_3d_truss_analysis_b200.dataset import PackedDataset
    files = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_generate", "cube-7_case_*.json")))
    gold = PackedDataset.from_json_files(files, 3, isOutputFile=True)       # the shipped results
    ds = PackedDataset.from_json_files(files, 3)                              # inputs only
    assert not ds.solved
    info = ds.solve()
    assert not info.any() and ds.solved
    for k in ("u", "ext", "axial"):
        assert orc.normwise_err(ds.a[k], gold.a[k]) <= 1e-9, k
    assert np.allclose(ds.a["weight"], gold.a["weight"], rtol=1e-12)
    j = ds.json(3)
    assert set(j) >= {"joint", "force", "member", "displace", "external", "internal", "weight"}
