"""CPU-side checks of the C-ABI library and the host logic (no compute calls: there is no GPU here).

 * the library loads and exports every function include/truss_b200.h declares
 * the plan's DOF maps are bit-exact against the oracle's boolean-mask order (truss.py:319-326,343)
 * the plan's scatter map, replayed in numpy, rebuilds the oracle's reduced K bit for bit (truss.py:307-316)
 * argument errors map to the documented codes; solving without a device fails loudly
"""
import json
import os
import re

import numpy as np
import pytest

from oracle import truss_oracle as orc
from python_stable_3d_truss_analysis_b200 import _lib
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def all_cases():
    cases = [(n, d, data) for n, d, data, _ in H.shipped_cases()]
    cases += [(f"random{i}", c["dim"], c["data"]) for i, c in enumerate(H.load_json("live_random.json"))]
    cases += [(n, d, data) for n, d, data, _ in H.cube7_shipped()[:3]]
    return cases


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "truss_b200.h")).read()
    declared = set(re.findall(r"\b(tb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    L = _lib.lib()
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, f"library does not export {missing}"
    assert set(_lib.EXPORTS) == declared
    assert L.tb_version() == 100


@pytest.mark.parametrize("name,dim,data", all_cases(), ids=lambda v: v if isinstance(v, str) else None)
def test_plan_dof_maps_bit_exact(name, dim, data):
    joints, support, conn, aed, force = orc.arrays_from_json(data, dim)
    plan = _lib.Plan(dim, conn, support)
    free_idx, dof2free, sup_idx = plan.maps()
    want = orc.dof_maps(dim, support)
    assert np.array_equal(free_idx, want[0]) and free_idx.dtype == np.int32
    assert np.array_equal(dof2free, want[1])
    assert np.array_equal(sup_idx, want[2])
    assert plan.stable == orc.is_stable(dim, support, conn.shape[0])
    assert plan.info.n_resist == sum(orc.resistance_number(int(s), dim) for s in support)


@pytest.mark.parametrize("name,dim,data", all_cases(), ids=lambda v: v if isinstance(v, str) else None)
def test_scatter_map_rebuilds_reference_K(name, dim, data):
    joints, support, conn, aed, force = orc.arrays_from_json(data, dim)
    plan = _lib.Plan(dim, conn, support)
    row, col, ptr, mem, loc = plan.scatter()
    # entries are unique, lower-triangular, row-major sorted
    assert np.all(row >= col)
    order = np.lexsort((col, row))
    assert np.array_equal(order, np.arange(len(row)))
    # contributions of one entry come in ascending member order
    for e in range(len(row)):
        seg = mem[ptr[e]:ptr[e + 1]]
        assert np.all(np.diff(seg) >= 0)
    # replay: k * (+-(c_i c_j)) summed in list order == reference K[mask][:, mask]
    d = dim
    K = np.zeros((plan.n, plan.n))
    for e in range(len(row)):
        v = 0.0
        for p in range(ptr[e], ptr[e + 1]):
            m = mem[p]
            la, lb = divmod(int(loc[p]), 2 * d)
            A, i = divmod(la, d)
            B, j = divmod(lb, d)
            L = orc.member_length(joints[conn[m, 0]], joints[conn[m, 1]])
            c = orc.member_cosines(joints[conn[m, 0]], joints[conn[m, 1]], L)
            pr = c[i] ** 2. if i == j else c[i] * c[j]
            if A != B:
                pr = -pr
            v = v + orc.member_k(aed[m, 0], aed[m, 1], L) * pr
        K[row[e], col[e]] = v
    mask = orc.free_mask(dim, support)
    Kref = orc.assemble_K(dim, joints, conn, aed)[np.ix_(mask, mask)]
    assert np.array_equal(np.tril(Kref), K), "scatter map does not rebuild the reference matrix bit for bit"
    natural = int(np.abs(row - col).max()) if len(row) else 0
    if plan.info.reordered:       # internal RCM elimination order: never wider than the reference's order
        assert plan.info.half_bandwidth <= natural
    else:
        assert plan.info.half_bandwidth == natural


def test_argument_errors():
    conn = np.array([[0, 1]], np.int32)
    sup = np.array([1, 0], np.uint8)
    with pytest.raises(_lib.TrussLibError) as e:
        _lib.Plan(4, conn, sup)
    assert e.value.code == -2
    with pytest.raises(_lib.TrussLibError) as e:
        _lib.Plan(3, np.array([[0, 2]], np.int32), sup)
    assert e.value.code == -4
    with pytest.raises(_lib.TrussLibError) as e:
        _lib.Plan(2, conn, np.array([4, 0], np.uint8))      # ROLLER_Z does not exist in 2D (type.py:66-74)
    assert e.value.code == -5
    with pytest.raises(_lib.TrussLibError) as e:
        _lib.Plan(3, conn, np.array([7, 0], np.uint8))
    assert e.value.code == -5
    assert _lib.lib().tb_strerror(-7).decode().startswith("no CUDA device")


def _has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_cuda(), reason="only meaningful on a machine without a GPU")
def test_no_cpu_fallback():
    from python_stable_3d_truss_analysis_b200.truss import Truss
    data = json.load(open(f"{H.GOLDEN}/ref_data/bar-6_input_0.json"))
    t = Truss(3).LoadFromJSON(data=data)
    with pytest.raises(_lib.NoCudaDeviceError):
        t.Solve()
    assert not t.isSolved


def test_path_selection_and_padding():
    for name, dim, data, _ in H.shipped_cases():
        _, support, conn, _, _ = orc.arrays_from_json(data, dim)
        plan = _lib.Plan(dim, conn, support)
        big = name.startswith("bar-942")
        assert plan.path == (2 if big else 0)        # bar-942: half-bandwidth 56 -> band pipeline
        assert plan.info.n_pad % 64 == 0 and plan.info.n_pad >= plan.n
        assert plan.info.band_blocks >= 1 and plan.info.envelope_size <= plan.n * (plan.n + 1) // 2
        if big:
            assert plan.info.band_blocks == 4 and plan.info.half_bandwidth == 56
            assert plan.info.n_tiles_nonzero == 21 and plan.info.n_tile_products == 10
        for path in (1, 2, 0 if not big else 2):
            plan.set_path(path)
            assert plan.path == path


def test_ga_and_augment_argument_errors_without_a_device():
    """tb_ga_* / tb_augment_ragged validate their arguments before touching the GPU, and fail loudly without one."""
    import ctypes as C
    L = _lib.lib()
    ok = _lib.TbGaParams(64, 8, 12, 3, 0.7, 0.1, 0.1, 1)
    dummy = np.zeros(64 * 12, np.int32)
    assert L.tb_ga_init(None, None, dummy.ctypes.data, None) == -1                      # TB_ERR_NULL
    bad = _lib.TbGaParams(64, 80, 12, 3, 0.7, 0.1, 0.1, 1)                               # more elites than population
    assert L.tb_ga_init(C.byref(bad), None, dummy.ctypes.data, None) == -3               # TB_ERR_SIZE
    bad = _lib.TbGaParams(64, 8, 12, 1, 0.7, 0.1, 0.1, 1)                                # one member type (OnlyOneMemberTypeError)
    assert L.tb_ga_init(C.byref(bad), None, dummy.ctypes.data, None) == -3
    bad = _lib.TbGaParams(64, 8, 12, 3, 0.7, 0.3, 0.1, 1)                                # probabilities above one
    assert L.tb_ga_step(C.byref(bad), 0, dummy.ctypes.data, None, dummy.ctypes.data, None, dummy.ctypes.data, None, None) == -3
    assert L.tb_ga_step(C.byref(ok), 0, None, None, dummy.ctypes.data, None, dummy.ctypes.data, None, None) == -1
    prm = _lib.TbAugmentParams()
    assert L.tb_augment_ragged(None, 4, None, None, None, C.byref(prm), None, None, None, None, None, None) == -1
    if not _has_cuda():
        assert L.tb_ga_init(C.byref(ok), None, dummy.ctypes.data, None) == _lib.TB_ERR_NO_DEVICE


def test_fused_path_is_chosen_by_shared_memory_not_by_size_alone():
    """53 joints / 600 members (N = 159 <= 160, M <= 1024) passes the size limits of the fused kernels but
    needs more shared memory than either of them has: the plan must route it to a blocked pipeline, and a ragged batch
    holding such a truss must be refused up front instead of failing at launch time."""
    rng = np.random.default_rng(3)
    nj, m = 53, 600
    pairs = set()
    while len(pairs) < m:
        a, b = rng.integers(0, nj, 2)
        if a != b:
            pairs.add((int(min(a, b)), int(max(a, b))))
    conn = np.array(sorted(pairs), np.int32)
    support = np.zeros(nj, np.uint8)
    support[:1] = orc.PIN                              # n = 156 free DOFs
    plan = _lib.Plan(3, conn, support)
    assert plan.info.path in (1, 2), "a truss that does not fit the fused kernels' shared memory was routed to them"
    with pytest.raises(_lib.TrussLibError) as ei:
        plan.set_path(0)
    assert ei.value.code == _lib.TB_ERR_TOO_LARGE
    assert not _lib.small_path_fits(3, nj, m)
    assert _lib.small_path_fits(3, 32, 119)            # the cube-7 trusses of the generator
    # 2-D: 80 joints, 400 members
    assert not _lib.small_path_fits(2, 80, 400) or _lib.Plan(2, conn[conn.max(axis=1) < 80][:400], np.r_[np.ones(2, np.uint8), np.zeros(78, np.uint8)]).info.path == 0


def test_gencube_limits_and_argument_errors():
    """tb_gencube_limits is host arithmetic (strides of the generator's fixed-stride scratch); bad parameters are refused
    before any device work (generate.py:152-336 on the device, csrc/tb_gencube.cu)."""
    import ctypes as C
    L = _lib.lib()
    prm = _lib.TbGencubeParams()
    for i, g in enumerate((5, 5, 5)):
        prm.grid[i] = g
    prm.ncube_lo, prm.ncube_hi = 7, 7
    prm.method, prm.link_type, prm.n_type, prm.max_attempts = 2, 3, 1, 8
    prm.length_lo, prm.length_hi = 50.0, 150.0
    mj, mm, mc = C.c_int32(), C.c_int32(), C.c_int32()
    assert L.tb_gencube_limits(C.byref(prm), C.byref(mj), C.byref(mm), C.byref(mc)) == 0
    assert (mj.value, mm.value, mc.value) == (56, 168, 7)          # 8 corners and 24 links per cube at most
    prm.ncube_hi = 1000                                           # more cubes than cells: capped by the grid
    assert L.tb_gencube_limits(C.byref(prm), C.byref(mj), C.byref(mm), C.byref(mc)) == 0
    assert mc.value == 125 and mj.value == 216
    prm.grid[0] = 50                                              # grid beyond TB_GEN_MAX_CELLS
    assert L.tb_gencube_limits(C.byref(prm), None, None, None) == _lib.TB_ERR_TOO_LARGE
    prm.grid[0] = 5
    prm.ncube_lo = 0
    assert L.tb_gencube_limits(C.byref(prm), None, None, None) == -3
    prm.ncube_lo, prm.ncube_hi, prm.method = 7, 7, 9               # unknown GenerateMethod
    assert L.tb_gencube(C.byref(prm), 4, *([None] * 13)) == -3
    prm.method = 2
    assert L.tb_gencube(C.byref(prm), 4, *([None] * 13)) == -1     # NULL outputs
    assert L.tb_gencube(C.byref(prm), 0, *([None] * 13)) == 0      # empty batch
    if not _has_cuda():
        buf = (C.c_double * 8)()
        p = C.cast(buf, C.c_void_p)
        assert L.tb_gencube(C.byref(prm), 1, p, p, p, p, p, p, p, p, p, None, None, None, None) == _lib.TB_ERR_NO_DEVICE


def test_compact_output_fields_are_part_of_the_abi():
    """tb_batch_out carries the compact pair as its last two fields; the ragged entry points refuse them (no common n / s)."""
    import ctypes as C
    assert [f[0] for f in _lib.TbBatchOut._fields_] == ["u", "ext", "axial", "weight", "info", "u_free", "react"]
    assert C.sizeof(_lib.TbBatchOut) == 7 * C.sizeof(C.c_void_p)
    jo, mo = np.array([0, 3], np.int64), np.array([0, 2], np.int64)
    ri = _lib.TbRaggedIn(3, 1, jo.ctypes.data, mo.ctypes.data, 0, 0, 0, 0, 0, 3, 2)
    bo = _lib.TbBatchOut(None, None, None, None, None, 1234, None)
    assert _lib.lib().tb_solve_ragged_host(C.byref(ri), C.byref(bo)) == -3
    assert _lib.lib().tb_solve_ragged(C.byref(ri), C.byref(bo), None) == -3
