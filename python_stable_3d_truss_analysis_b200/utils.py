"""Exceptions, tolerances and small helpers of the truss API.

Mirrors the hot-path part of the reference's ``slientruss3d/utils.py`` (exceptions ``:52-67``,
``CheckDim :70-74``, ``IsZero/IsZeroVector :79-84``, ``GetLength :87-88``, misc ``:91-121``).
The reference's plotting helpers (``Arrow3D`` ...) and its stray tkinter import are not part of
the solve path and are not reproduced.
"""
from __future__ import annotations

import numpy as np

INF = float("inf")
ZERO_EPS = 1e-10


class InvalidSupportTypeError(Exception): pass
class InvalidMetapathTypeError(Exception): pass
class InvalidTaskTypeError(Exception): pass
class InvalidLinkTypeError(Exception): pass
class InvalidGenerateMethodError(Exception): pass
class TrussNotStableError(Exception): pass
class TrussNotSolvedError(Exception): pass
class DimensionError(Exception): pass
class InvaildJointError(Exception): pass
class EliteNumberTooMuchError(Exception): pass
class ProbabilityGreaterThanOneError(Exception): pass
class OnlyOneMemberTypeError(Exception): pass
class MinStressTooLargeError(Exception): pass
class MinDisplaceTooLargeError(Exception): pass
class NotAllBeSetError(Exception): pass
class PinNotEnoughError(Exception): pass


def CheckDim(dim):
    if dim not in (2, 3):
        raise DimensionError(f"Dimension of truss and member must be 2 or 3, but got [{dim}].")
    return dim


def IsZero(num, eps=ZERO_EPS):
    return abs(num) < eps


def IsZeroVector(vec, eps=ZERO_EPS):
    return bool((np.abs(np.asarray(vec, dtype=np.float64)) < eps).all())


def GetLength(vec):
    v = np.asarray(vec, dtype=np.float64)
    return float((v * v).sum() ** 0.5)


def MinNorm(vec, minNorm=1.):
    return vec * max(1., minNorm / np.linalg.norm(vec))


def GetPowerset(s):
    for bits in range(1 << len(s)):
        yield [s[j] for j in range(len(s)) if bits >> j & 1]


def GetCenter(position0, position1):
    return [0.5 * (a + b) for a, b in zip(position0, position1)]


def GetAngles(position0, position1):
    lo, hi = (position0, position1) if position0[-1] < position1[-1] else (position1, position0)
    vec = [b - a for a, b in zip(lo, hi)]
    full = sum(v ** 2. for v in vec) ** 0.5
    flat = sum(v ** 2. for v in vec[:2]) ** 0.5
    if IsZero(flat):
        return flat / full, vec[2] / full, 0., 0.
    return flat / full, vec[2] / full, vec[1] / flat, vec[0] / flat


def InfinteLoop():
    i = 0
    while True:
        yield i
        i += 1
