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


def cube_truss(n_side: int, seed: int = 1):
    """Full n^3 cube-truss grid (n_side = 12 -> nJ 2197, M 14868, n 6084)."""
    from python_stable_3d_truss_analysis_b200.generate import GenerateRandomCubeTrusses
    from python_stable_3d_truss_analysis_b200.type import LinkType

    ts = GenerateRandomCubeTrusses(gridRange=(n_side,) * 3, numCubeRange=(n_side ** 3,) * 2, numEachRange=(1, 1),
                                   lengthRange=(100, 200), forceRange=[(-1000, 1000)] * 3,
                                   linkType=LinkType.LeftBottom_RightTop, seed=seed, isPrintMessage=False)
    return ts[0]


def ragged_pool_arrays(pool, n_total: int, sigma: float = 10.0, seed: int = 1):
    """Config 4 recipe: tile a pool of generated trusses to n_total systems with fresh Gaussian joint
    jitter (AddJointNoise-style, sigma, default_rng(seed)).  Returns the tb_ragged_in arrays + the pool index of each system."""
    packs = [t._pack() for t in pool]
    which = np.arange(n_total) % len(pool)
    nj = np.array([p[0].shape[0] for p in packs])[which]
    nm = np.array([p[2].shape[0] for p in packs])[which]
    joint_off = np.zeros(n_total + 1, np.int64); joint_off[1:] = np.cumsum(nj)
    member_off = np.zeros(n_total + 1, np.int64); member_off[1:] = np.cumsum(nm)
    cat = lambda i, dt: np.concatenate([np.asarray(packs[w][i]).reshape(-1) for w in which]).astype(dt)  # noqa: E731
    xyz = cat(0, np.float64)
    rng = np.random.default_rng(seed)
    xyz = xyz + rng.normal(0.0, sigma, size=xyz.shape)
    return joint_off, member_off, xyz, cat(1, np.uint8), cat(2, np.int32), cat(3, np.float64), cat(4, np.float64), which
