"""Bulk I/O (python_stable_3d_truss_analysis_b200/dataset.py): the packed binary container against the reference's per-truss
JSON format, using the solved cube-7 files the reference ships (generate/cube-7_case_*.json) as known answers."""
import copy
import json

import numpy as np

from python_stable_3d_truss_analysis_b200.dataset import PackedDataset
from python_stable_3d_truss_analysis_b200.truss import Truss
from tests import helpers as H


def _shipped():
    cases = H.cube7_shipped()
    trusses = [Truss(3).LoadFromJSON(data=copy.deepcopy(data), isOutputFile=True) for _, _, data, _ in cases]
    return cases, trusses


def test_packed_views_reproduce_the_shipped_json():
    cases, trusses = _shipped()
    ds = PackedDataset.from_trusses(trusses, names=[c[0] for c in cases])
    assert len(ds) == 10 and ds.solved
    for i, (_, _, gold, _) in enumerate(cases):
        got = ds.json(i)
        assert got["joint"] == gold["joint"] and got["member"] == gold["member"]
        assert {j: tuple(v) for j, v in got["force"]} == {j: tuple(v) for j, v in gold["force"]}
        for key in ("displace", "external", "internal"):
            assert got[key] == gold[key], key                   # same sparse entries, same order, same doubles
        # Truss.weight is the reference's sum(member.weight ...) (truss.py:166-168); since Python 3.12 the built-in sum()
        # of floats is compensated, so it differs from the shipped file (written under an older Python) in the last bit
        assert abs(got["weight"] - gold["weight"]) <= 1e-14 * gold["weight"]


def test_container_round_trip_and_truss_views(tmp_path):
    cases, trusses = _shipped()
    ds = PackedDataset.from_trusses(trusses, names=[c[0] for c in cases])
    for compressed in (False, True):
        back = PackedDataset.load(ds.save(str(tmp_path / f"d{int(compressed)}.npz"), compressed=compressed))
        assert back.dim == 3 and back.names == ds.names
        for k, v in ds.a.items():
            assert np.array_equal(back.a[k], v), k
    t = ds.truss(4)
    ref = trusses[4]
    assert t.isSolved and t.GetInternalForces() == ref.GetInternalForces() and t.weight == ref.weight
    assert {j: tuple(v) for j, v in t.GetDisplacements().items()} == {j: tuple(v) for j, v in ref.GetDisplacements().items()}
    assert {j: tuple(v) for j, v in t.GetResistances().items()} == {j: tuple(v) for j, v in ref.GetResistances().items()}
    paths = ds.dump_json(str(tmp_path / "json"), indices=[0, 9])
    assert [p.rsplit("/", 1)[1] for p in paths] == [cases[0][0] + ".json", cases[9][0] + ".json"]
    again = Truss(3).LoadFromJSON(paths[1], isOutputFile=True)
    assert again.GetInternalForces() == trusses[9].GetInternalForces()
    assert json.load(open(paths[0]))["internal"] == cases[0][2]["internal"]


def test_unsolved_and_failed_entries():
    cases, trusses = _shipped()
    plain = [Truss(3).LoadFromJSON(data={k: copy.deepcopy(v) for k, v in c[2].items() if k in ("joint", "force", "member")}) for c in cases[:3]]
    ds = PackedDataset.from_trusses(plain)
    assert not ds.solved and "displace" not in ds.json(0) and not ds.truss(1).isSolved
    full = PackedDataset.from_trusses(trusses[:3])
    full.a["info"] = np.array([0, 3, 0], np.int32)              # a failed system carries no results
    assert "displace" in full.json(0) and "displace" not in full.json(1) and not full.truss(1).isSolved
