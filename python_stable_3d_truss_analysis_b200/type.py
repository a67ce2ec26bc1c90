"""Value types of the truss API: ``MemberType`` and ``SupportType`` plus the small enums.

Same names, codes and semantics as the reference's ``slientruss3d/type.py`` (``MemberType :5-27``,
``SupportType :30-89``, enums ``:91-111``); the integer codes of ``SupportType`` are also the codes
the CUDA kernels read (``include/truss_b200.h``).
"""
from __future__ import annotations

import numpy as np

from .utils import CheckDim, IsZero, InvalidSupportTypeError


class MemberType:
    """Cross-section area ``a``, Young's modulus ``e``, density."""

    __slots__ = ("a", "e", "density")

    def __init__(self, a=1., e=1., density=1.):
        self.a, self.e, self.density = float(a), float(e), float(density)

    def __repr__(self):
        return f"MemberType(a={self.a}, e={self.e}, density={self.density})"

    def __eq__(self, other):
        return all(IsZero(x - y) for x, y in zip(self.Serialize(), other.Serialize()))

    def __hash__(self):
        return hash((self.a, self.e, self.density))

    def Set(self, other):
        self.a, self.e, self.density = other.a, other.e, other.density

    def Serialize(self):
        return [self.a, self.e, self.density]

    def Copy(self):
        return MemberType(self.a, self.e, self.density)


class SupportType:
    NO = 0
    PIN = 1
    ROLLER_X = 2
    ROLLER_Y = 3
    ROLLER_Z = 4

    _NAMES = ("NO", "PIN", "ROLLER_X", "ROLLER_Y", "ROLLER_Z")

    @staticmethod
    def _axis_flags(supportType, dim):
        """Restrained axes of one joint; raises for codes that do not exist in this dimension."""
        CheckDim(dim)
        if supportType == SupportType.PIN:
            return [True] * dim
        if supportType == SupportType.NO:
            return [False] * dim
        axis = supportType - SupportType.ROLLER_X if isinstance(supportType, (int, np.integer)) else -1
        if 0 <= axis < dim:
            return [i == axis for i in range(dim)]
        raise InvalidSupportTypeError(f"[GetResistanceMask] No such {dim}D-support type [{supportType}] !")

    @staticmethod
    def GetResistanceNumber(supportType, dim):
        if supportType == SupportType.PIN:
            return dim
        if supportType in (SupportType.ROLLER_X, SupportType.ROLLER_Y, SupportType.ROLLER_Z):
            return 1
        if supportType == SupportType.NO:
            return 0
        raise InvalidSupportTypeError(f"[GetResistanceNumber] No such support type [{supportType}] !")

    @staticmethod
    def GetResistanceMask(supportType, dim):
        return np.array(SupportType._axis_flags(supportType, dim))

    @staticmethod
    def GetFromString(string):
        if isinstance(string, str) and string in SupportType._NAMES:
            return SupportType._NAMES.index(string)
        raise InvalidSupportTypeError(f"[GetFromString] No such support type [{string}] !")

    @staticmethod
    def GetFromType(supportType):
        if isinstance(supportType, (int, np.integer)) and 0 <= supportType < len(SupportType._NAMES):
            return SupportType._NAMES[supportType]
        return None


class MetapathType:
    USE_IMPLICIT = 0
    NO_IMPLICIT = 1


class TaskType:
    OPTIMIZATION = 0
    REGRESSION = 1


class LinkType:
    LeftBottom_RightTop = 0
    RightBottom_LeftTop = 1
    Cross = 2
    Random = 3


class GenerateMethod:
    DFS = 0
    BFS = 1
    Random = 2
