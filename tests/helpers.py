"""Shared loaders for the parity fixtures under tests/golden (see make_golden.py)."""
from __future__ import annotations

import glob
import json
import os

import numpy as np

from oracle import truss_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-9          # north_star: relative (norm-wise) tolerance on displacements, forces, reactions
FIELDS = ("u", "ext", "axial")


def dim_of(name: str) -> int:
    return 2 if name.startswith(("bar-10_", "bar-47_")) else 3


def shipped_cases():
    """[(name, dim, input_dict, golden dense dict)] for the reference's data/bar-* files."""
    out = []
    for f in sorted(glob.glob(os.path.join(GOLDEN, "ref_data", "*_input_*.json"))):
        name = os.path.basename(f)[:-5]
        dim = dim_of(name)
        data = json.load(open(f))
        gold = json.load(open(f.replace("_input_", "_output_")))
        out.append((name, dim, data, dense_from_output(gold, dim)))
    return out


def dense_from_output(gold: dict, dim: int):
    nj, nm = len(gold["joint"]), len(gold["member"])
    return {"u": orc.dense_from_sparse(gold["displace"], nj, dim),
            "ext": orc.dense_from_sparse(gold["external"], nj, dim),
            "axial": orc.dense_from_sparse(gold["internal"], nm),
            "weight": float(gold["weight"])}


def cube7_shipped():
    out = []
    for f in sorted(glob.glob(os.path.join(GOLDEN, "ref_generate", "cube-7_case_*.json")),
                    key=lambda p: int(p.rsplit("_", 1)[1][:-5])):
        gold = json.load(open(f))
        out.append((os.path.basename(f)[:-5], 3, gold, dense_from_output(gold, 3)))
    return out


def load_json(name):
    return json.load(open(os.path.join(GOLDEN, name)))


def assert_close(got: dict, want: dict, tol=TOL, what=""):
    for k in FIELDS:
        err = orc.normwise_err(got[k], want[k])
        assert err <= tol, f"{what} field {k}: norm-wise err {err:.3e} > {tol:g}"
    w = want["weight"]
    assert abs(got["weight"] - w) <= tol * max(1.0, abs(w)), f"{what} weight {got['weight']} vs {w}"


def key_sets_match(dense_got, dense_want, rel=1e-8):
    """SURVEY section 4 trap 2: sparse-dict key sets are only reproducible away from the 1e-10 cutoff."""
    a, b = np.asarray(dense_got).ravel(), np.asarray(dense_want).ravel()
    scale = max(np.abs(b).max(), 1e-300)
    sig = np.abs(b) > rel * scale
    return bool(np.all((np.abs(a[sig]) >= orc.ZERO_EPS) == (np.abs(b[sig]) >= orc.ZERO_EPS)))
