"""numpy replay of the two-sided band program (csrc/tb_ts.cuh) -- test infrastructure only.

Executes, block by block and with the kernel's own ring-slot arithmetic, what ``k_band_ts`` (csrc/tb_bandts.cu) does with
the program ``Plan.ts_program()`` returns: per-column assembly from member products (truss.py:56-86, 307-316), left-looking
block products driven by the structural masks, the hand-over of the bottom side's Schur complement to the separator,
the chunk layout of the factor and the two back substitutions.  It checks the integer program (orders, flips, masks,
slots, chunk offsets) on the CPU; the thread-level layout of the kernel is covered by the GPU parity tests.
"""
from __future__ import annotations

import numpy as np

BT = 8


def member_products(dim, xyz, conn_pair, a, e):
    """(k, c) of one member with the reference's roundings (truss.py:19, 56-63)."""
    j0, j1 = conn_pair
    dx = [xyz[j1][i] - xyz[j0][i] for i in range(dim)]
    l2 = dx[0] * dx[0]
    for i in range(1, dim):
        l2 = l2 + dx[i] * dx[i]
    length = float(np.sqrt(l2))
    k = e * a / length
    c = [dx[i] / length for i in range(dim)]
    return k, c


def _term(dim, prods, pk):
    """One contribution of the assembly pass's lists: member << 4 | negate << 3 | index of the cosine product."""
    k, c = prods[pk >> 4]
    ij = pk & 7
    if dim == 3:
        i = (ij >= 3) + (ij >= 5)
        j = ij if ij < 3 else (ij - 2 if ij < 5 else 2)
    else:
        i = int(ij >= 2)
        j = int(ij >= 1)
    t = k * (c[i] * c[j])
    return -t if (pk & 8) else t


def assemble_program_order(prog, dim, xyz, conn, aed):
    """kv[e]: what the assembly pass (k_prep on tq_first / tq_multi / tq_ptr / tq_pack) writes, in the kernel's entry
    order: contributions summed in list order (ascending member, truss.py:310-314)."""
    prods = [member_products(dim, xyz, conn[m], aed[m][0], aed[m][1]) for m in range(len(conn))]
    ptr, pack, first = prog["tq_ptr"], prog["tq_pack"], prog["tq_first"]
    kv = np.zeros(len(ptr) - 1)
    multi = set(int(q) for q in prog["tq_multi"])
    for q in range(len(kv)):
        if ptr[q + 1] == ptr[q]:          # hole of the bank-conflict-free entry schedule: skipped by both loops of the pass
            assert first[q] == -2**31 and q not in multi
            continue
        if q in multi:
            assert first[q] < 0
            v = 0.0
            for t in range(int(ptr[q]), int(ptr[q + 1])):
                v = v + _term(dim, prods, int(pack[t]))
        else:
            assert first[q] >= 0 and ptr[q + 1] - ptr[q] == 1 and pack[ptr[q]] == first[q]
            v = 0.0 + _term(dim, prods, int(first[q]))
        kv[q] = v
    return kv


NSTAGE = 3


def main_doubles(nb, chunk_max):
    """csrc/tb_ts.cuh ts_main_doubles: ring of live blocks, reused by the back substitution's chunk buffers + u ring."""
    return max(nb * (nb + 1) // 2 * 64, NSTAGE * chunk_max + 9 * BT)


def _unpack_pos(epos, c, mainsz):
    """Staging position (byte offset from the side's shared-memory base, resolved by the plan) -> (rb, row, col); checks
    that a ring position is the slot the kernel's pointer ring hands to block (c+rb, c) (csrc/tb_ts.cuh ts_ring_slot)."""
    assert epos % 8 == 0
    off8 = epos // 8
    if off8 >= mainsz:
        rb, off = 0, off8 - mainsz
        assert off < 64
    else:
        slot, off = divmod(off8, 64)
        rb = 1
        while rb * (rb + 1) // 2 <= slot:
            rb += 1
        assert slot == rb * (rb - 1) // 2 + (rb - 1 - c % rb), "entry staged in a slot the ring does not give to this block"
    k = ((off >> 5) << 2) | (off & 3)
    r = ((off >> 2) & 7) ^ ((off >> 5) << 1)      # rows of the second k-slab are stored permuted (csrc/tb_ts.cuh ts_b8_off)
    return rb, r, k


class _Side:
    def __init__(self, prog, s):
        d = prog["side"][s]
        self.d = d
        self.s = s
        self.nb = prog["info"]["nb_top" if s == 0 else "nb_bottom"]
        self.tot = len(d["colmask"])
        nS = prog["info"]["nS"]
        self.own = self.tot - nS if self.tot else 0
        self.ring = {}          # slot -> 8x8 block
        self.idx = [0] * 10
        self.nzprev = [0] * 10
        self.ys = []            # y blocks of the processed columns (virtual order)
        self.chunks = {}        # c -> dict(Z, y, blocks{rb})

    def slot(self, e, p):
        return e * (e - 1) // 2 + p % e


def replay(prog, dim, xyz, conn, aed, force, n_dof):
    """Returns (u per DOF, dict with debug data).  ``aed`` is [M,3]."""
    info = prog["info"]
    kv = assemble_program_order(prog, dim, xyz, conn, aed)
    epos = prog["epos"]
    nS = info["nS"]
    sides = [_Side(prog, 0), _Side(prog, 1)]
    X = {}
    Zx = np.zeros((max(nS, 1), BT))
    two = sides[1].tot > 0
    NBK = max(info["nb_top"], info["nb_bottom"])

    def forward_column(S, c):
        d = S.d
        nb = S.nb
        own = c < S.own
        xcol = S.s == 1 and not own
        nzc, srcc, xm = int(d["colmask"][c]), int(d["srcmask"][c]), int(d["xmask"][c])
        acc = {rb: np.zeros((BT, BT)) for rb in range(nb + 1)}
        tp = np.zeros(BT)
        # the kernel takes the liveness of every block product from the column record (bit = flat index of (d, rb) for NB =
        # the wider side's band): it must be exactly "both blocks of column c-d exist"
        rec = d["colrec"][c]
        assert (int(rec[0]) & 0xffffffff) == (nzc | (xm << 9) | (int(d["colent"][c][1] - d["colent"][c][0]) << 18))
        assert int(rec[1]) == int(d["lofs"][c])
        pmask = (int(rec[2]) & 0xffffffff) | ((int(rec[3]) & 0xffffffff) << 32)
        idx, want = 0, 0
        for dd in range(1, NBK + 1):
            for rb in range(0, NBK - dd + 1):
                nzp = S.nzprev[dd] if dd <= nb else 0
                if (nzp >> dd) & 1 and (nzp >> (rb + dd)) & 1:
                    want |= 1 << idx
                idx += 1
        assert pmask == want, "product mask of the column record disagrees with the block structure"
        for dd in range(1, nb + 1):
            nzp = S.nzprev[dd]
            if not (nzp >> dd) & 1:
                continue
            Bm = S.ring[dd * (dd - 1) // 2 + S.idx[dd]]
            tp += Bm @ S.ys[c - dd]
            for rb in range(0, nb - dd + 1):
                e = rb + dd
                if not (nzp >> e) & 1:
                    continue
                sl = S.idx[e] - dd
                if sl < 0:
                    sl += e
                A = S.ring[e * (e - 1) // 2 + sl]
                acc[rb] += A @ Bm.T
        if xcol:
            jq = c - S.own
            for rb in range(nb + 1):
                if (nzc >> rb) & 1:
                    J = nS - 1 - jq - rb
                    assert J >= 0
                    X[(J, rb)] = acc[rb][::-1, ::-1].T.copy()     # element (r, k) -> (7-k, 7-r)
            Zx[nS - 1 - jq] = tp[::-1]
            S.ys.append(np.zeros(BT))
        else:
            if S.s == 0 and not own and two:
                J = c - S.own
                for rb in range(nb + 1):
                    if (xm >> rb) & 1:
                        acc[rb] += X[(J, rb)]
                tp += Zx[J]
            # staging: block rb in the slot of the dead block (c, c-rb)
            stage = {rb: np.zeros((BT, BT)) for rb in range(nb + 1) if (nzc >> rb) & 1}
            e_lo, e_hi = int(d["colent"][c][0]), int(d["colent"][c][1])
            nlive = int(np.sum(epos[e_lo:e_hi] >= 0))
            assert e_hi - e_lo <= 32 * ((nlive + 31) // 32), "the entry schedule takes more rounds than the entries need"
            for e in range(e_lo, e_hi):
                if epos[e] < 0:
                    assert prog["ent_src"][e] < 0
                    continue
                rb, r, k = _unpack_pos(int(epos[e]), c, main_doubles(S.nb, info["chunk_max"]))
                assert rb in stage, "entry in a block the mask calls zero"
                assert stage[rb][r, k] == 0.0, "two entries in one position"
                stage[rb][r, k] = kv[e]
            for r in range(BT):
                if d["rowdof"][c * BT + r] < 0:
                    stage[0][r, r] = 1.0
            P = {rb: stage[rb] - acc[rb] for rb in stage}
            fr = np.array([force[d["rowdof"][c * BT + r]] if d["rowdof"][c * BT + r] >= 0 else 0.0 for r in range(BT)])
            t = fr - tp
            Pd = np.tril(P[0]) + np.tril(P[0], -1).T
            Ld = np.linalg.cholesky(Pd)
            Zm = np.linalg.inv(Ld).T                      # Z = L_D^{-T}
            y = Zm.T @ t
            S.ys.append(y)
            blocks = {}
            for rb in range(1, nb + 1):
                if (nzc >> rb) & 1:
                    Lb = P[rb] @ Zm
                    blocks[rb] = Lb
                    S.ring[rb * (rb - 1) // 2 + S.idx[rb]] = Lb
            S.chunks[c] = dict(Z=Zm, y=y, blocks=blocks, lofs=int(d["lofs"][c]),
                               size=64 + 8 + 64 * bin(nzc >> 1).count("1"))
        for e in range(nb + 1, 1, -1):
            S.nzprev[e] = S.nzprev[e - 1]
        S.nzprev[1] = srcc
        for e in range(1, nb + 2):
            S.idx[e] = 0 if S.idx[e] + 1 == e else S.idx[e] + 1

    # the two sides are independent until the separator: bottom first (its hand-over must exist), then top
    for c in range(sides[1].tot):
        forward_column(sides[1], c)
    for c in range(sides[0].tot):
        forward_column(sides[0], c)

    # chunk offsets: contiguous, non-overlapping
    spans = sorted((ch["lofs"], ch["size"]) for S in sides for ch in S.chunks.values())
    for (o0, s0), (o1, _) in zip(spans, spans[1:]):
        assert o0 + s0 == o1, "factor chunks overlap or leave gaps"
    assert not spans or spans[-1][0] + spans[-1][1] == info["l_per_sys"]
    assert max((s for _, s in spans), default=0) == info["chunk_max"]

    u_dof = np.zeros(n_dof)
    us = [dict(), dict()]

    def back(S, cols):
        d = S.d
        for c in cols:
            ch = S.chunks[c]
            t = np.zeros(BT)
            for rb, Lb in ch["blocks"].items():
                t += Lb.T @ us[S.s][c + rb]
            u = ch["Z"] @ (ch["y"] - t)
            us[S.s][c] = u
            for r in range(BT):
                dof = d["rowdof"][c * BT + r]
                if dof >= 0:
                    u_dof[dof] = u[r]

    T = sides[0]
    back(T, range(T.tot - 1, T.own - 1, -1))
    if two:
        Bt = sides[1]
        for c in range(Bt.own, Bt.tot):
            us[1][c] = us[0][T.own + (nS - 1 - (c - Bt.own))][::-1]
        back(Bt, range(Bt.own - 1, -1, -1))
    back(T, range(T.own - 1, -1, -1))
    return u_dof, dict(sides=sides, X=X, kv=kv)
