"""Bulk I/O (python_stable_3d_truss_analysis_b200/dataset.py): the packed binary container against the reference's per-truss
JSON format, using the solved cube-7 files the reference ships (generate/cube-7_case_*.json) as known answers."""
import copy
import json
import os

import numpy as np
import pytest

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


# ---------------------------------------------------------------------------- bulk JSON loader (csrc/tb_json.cu)
def _ref_files(kind):
    import glob
    root = os.path.join(os.path.dirname(__file__), "golden")
    if kind == "inputs":
        return sorted(glob.glob(os.path.join(root, "ref_data", "bar-*_input_*.json")))
    if kind == "outputs":
        return sorted(glob.glob(os.path.join(root, "ref_data", "bar-*_output_*.json")))
    return sorted(glob.glob(os.path.join(root, "ref_generate", "cube-7_case_*.json")))


def _dim_of(path):
    return len(json.load(open(path))["joint"][0][0])


@pytest.mark.parametrize("kind,is_output", [("inputs", False), ("outputs", True), ("cube", True), ("cube", False)])
def test_bulk_json_loader_equals_per_truss_loading(kind, is_output):
    """N files -> packed arrays through the native parser == Truss.LoadFromJSON per file + from_trusses (truss.py:401-421),
    bit for bit, inputs and output files, 2D and 3D."""
    files = _ref_files(kind)
    for dim in (2, 3):
        sel = [f for f in files if _dim_of(f) == dim]
        if not sel:
            continue
        got = PackedDataset.from_json_files(sel, dim, isOutputFile=is_output, threads=3)
        want = PackedDataset.from_trusses([Truss(dim).LoadFromJSON(f, isOutputFile=is_output) for f in sel])
        assert len(got) == len(sel) and got.names == [os.path.basename(f)[:-5] for f in sel]
        for k in ("joint_off", "member_off", "xyz", "support", "conn", "aed", "force"):
            assert np.array_equal(got.a[k], want.a[k]), k
        assert got.solved == is_output
        if is_output:
            for k in ("u", "ext", "axial"):
                assert np.array_equal(got.a[k], want.a[k]), k
            assert np.allclose(got.a["weight"], want.a["weight"], rtol=1e-13)
            for i in (0, len(sel) - 1):      # (files without a "weight" key: recomputed vectorised, equal to rounding)
                a, b = got.json(i), want.json(i)
                wa, wb = a.pop("weight"), b.pop("weight")
                assert abs(wa - wb) <= 1e-12 * abs(wb)
                assert a == b


def test_bulk_json_loader_reference_semantics_and_errors():
    from python_stable_3d_truss_analysis_b200.utils import InvaildJointError, InvalidSupportTypeError
    base = {"joint": [[[0, 0, 0], "PIN"], [[1.5, 0, 0], "ROLLER_Z"], [[0, 2e0, 1]  , "NO"]],
            "force": [[2, [0, 0, -1e3]], [2, [0.0, 0.0, 0.0]], [1, [1, 2, 3]], [1, [4, 5, 6]]],
            "member": [[[0, 1], [1, 1e7, 0.1]], [[1, 2], [2.5, 1e7, 0.1]], [[0, 2], [1, 3e7, 0.2]]], "extra": {"a": [1, {"b": None}], "s": 'x"y'}}
    text = json.dumps(base, indent=2).encode()
    ds = PackedDataset.from_json_texts([text, text], 3)
    t = Truss(3).LoadFromJSON(data=base)
    ref = PackedDataset.from_trusses([t, t])
    for k in ("xyz", "support", "conn", "aed", "force", "joint_off", "member_off"):
        assert np.array_equal(ds.a[k], ref.a[k]), k
    # a zero load vector is ignored (truss.py:181), a later entry of the same joint replaces the earlier one
    assert ds.a["force"].reshape(2, 3, 3)[0].tolist() == [[0, 0, 0], [4, 5, 6], [0, 0, -1e3]]
    empty = PackedDataset.from_json_texts([], 3)
    assert len(empty) == 0

    def broken(**kw):
        d = json.loads(json.dumps(base))
        d.update(kw)
        return json.dumps(d).encode()

    with pytest.raises(InvalidSupportTypeError):
        PackedDataset.from_json_texts([text, broken(joint=[[[0, 0, 0], "HINGE"]])], 3)
    with pytest.raises(InvaildJointError):
        PackedDataset.from_json_texts([broken(member=[[[0, 7], [1, 1, 1]]])], 3)
    with pytest.raises(InvaildJointError):
        PackedDataset.from_json_texts([broken(force=[[5, [1, 0, 0]]])], 3)
    for bad in (b"", b"[1, 2]", text[:-5], broken(member=[[[0, 1], [1, 1]]]), json.dumps({"joint": []}).encode(),
                broken(joint=[[[0, 0], "PIN"]])):
        with pytest.raises(ValueError):
            PackedDataset.from_json_texts([bad], 3, names=["bad"])
    with pytest.raises(ValueError):          # an input file read as an output file lacks the result lists
        PackedDataset.from_json_texts([text], 3, isOutputFile=True)
