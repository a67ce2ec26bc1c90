"""Random cube-truss dataset generator with augmentation; the solves run as ONE batched GPU call.

Same public surface as the reference's ``slientruss3d/generate.py`` (``GenerateRandomCubeTrusses
:314-376``, ``CubeGrid :236-311``, ``CubeTruss :152-231``, augmenters ``:13-148``; documented in
``detail/gen_truss.md``).  Topology generation stays host Python and replays the reference's
``random`` call sequence draw for draw, so ``seed=42`` reproduces the shipped
``generate/cube-7_case_*.json`` inputs bit for bit (tests/test_generate.py).  What changes is the
solve site (``generate.py:354-357``): the reference builds and solves truss by truss; here every
truss is generated first (the stability retry uses the host-side counting rule, exactly the test
``Solve()`` applies) and the whole list goes through ``SolveBatch`` -- a ragged batch on the fused
shared-memory kernel.

Not reproduced: ``isPlotTruss`` (matplotlib figures are outside the hot path; a warning is printed).
Augmenters called on a ``Truss`` rebuild it in place; the reference re-loads into the same object
and so duplicates every joint and member (``generate.py:57-58,74-75,94-95,135-136``).
"""
from __future__ import annotations

import json
import os
import random
from math import ceil

from .batch import SolveBatch
from .truss import Truss
from .type import GenerateMethod, LinkType, MemberType
from .utils import PinNotEnoughError, TrussNotStableError


# --------------------------------------------------------------------------- augmentation
class TrussDataAugmenter:
    @staticmethod
    def IsTrussClass(trussData):
        if isinstance(trussData, Truss):
            return True, trussData.Serialize()
        return False, trussData

    @staticmethod
    def GetCentroid(jointDict):
        total = [0., 0., 0.]
        for position, _ in jointDict:
            total = [total[i] + position[i] for i in range(3)]
        return [v / len(jointDict) for v in total]

    @staticmethod
    def GetStableMinNumPin(trussData):
        return ceil((len(trussData['joint']) * 3 - len(trussData['member'])) / 3)

    # helper shared by the concrete augmenters: edit the dict, write back into a Truss if one was given
    def _apply(self, trussData, edit):
        isTruss, data = self.IsTrussClass(trussData)
        edit(data)
        if isTruss:
            solved = trussData.isSolved
            trussData.__init__(trussData.dim)
            trussData.LoadFromJSON(data=data, isOutputFile=solved)
        return trussData


class NoChange(TrussDataAugmenter):
    """Do nothing to the truss."""
    def __call__(self, trussData):
        return trussData


class AddJointNoise(TrussDataAugmenter):
    """Add gaussian noise to the position of every joint."""
    def __init__(self, noiseMeans=[0., 0., 0.], noiseStds=[1., 1., 1.]):
        self.noiseMeans, self.noiseStds = noiseMeans, noiseStds

    def __call__(self, trussData):
        def edit(data):
            for joint in data['joint']:
                joint[0][:] = [joint[0][i] + random.gauss(self.noiseMeans[i], self.noiseStds[i]) for i in range(3)]
        return self._apply(trussData, edit)


class MoveToCentroid(TrussDataAugmenter):
    """Move the centroid of the truss to the origin."""
    def __call__(self, trussData):
        def edit(data):
            centroid = self.GetCentroid(data['joint'])
            for joint in data['joint']:
                joint[0][:] = [joint[0][i] - centroid[i] for i in range(3)]
        return self._apply(trussData, edit)


class Translation(TrussDataAugmenter):
    """Translate the whole truss."""
    def __init__(self, translation):
        self.translation = translation

    def __call__(self, trussData):
        def edit(data):
            for joint in data['joint']:
                joint[0][:] = [joint[0][i] + self.translation[i] for i in range(3)]
        return self._apply(trussData, edit)


class RandomTranslation(TrussDataAugmenter):
    """Translate the whole truss by a random vector."""
    def __init__(self, translateRange=[-1., 1.]):
        self.translateRange = translateRange

    def __call__(self, trussData):
        return Translation([random.uniform(*self.translateRange) for _ in range(3)])(trussData)


class RandomResetPin(TrussDataAugmenter):
    """Re-draw which joints are pin supports (at most nJoint * maxNumPinRatio of them)."""
    def __init__(self, minNumPin=3, maxNumPinRatio=None):
        if minNumPin < 3:
            raise PinNotEnoughError("Number of pins must >= 3.")
        self.minNumPin, self.maxNumPinRatio = minNumPin, maxNumPinRatio

    def __call__(self, trussData):
        def edit(data):
            joints = data['joint']
            need = self.GetStableMinNumPin(data)
            lo = need if self.minNumPin is None else max(self.minNumPin, need)
            hi = len(joints) if self.maxNumPinRatio is None else int(self.maxNumPinRatio * len(joints))
            pins = set(random.sample(range(len(joints)), k=random.choice(range(lo, hi + 1))))
            for jointID, joint in enumerate(joints):
                joint[-1] = "PIN" if jointID in pins else "NO"
        return self._apply(trussData, edit)


class TrussDataAugmenterList(TrussDataAugmenter):
    def __init__(self, *augmenters):
        self.augmenters = augmenters

    def __call__(self, trussData):
        for augmenter in self.augmenters:
            trussData = augmenter(trussData)
        return trussData


# --------------------------------------------------------------------------- cube blocks
# Corner i of a unit cube at (x, y, z) is (x + bit0(i), y + bit1(i), z + bit2(i)).
_FACE_DIAGONALS = (((0, 5), (1, 4)), ((1, 7), (3, 5)), ((3, 6), (2, 7)), ((2, 4), (0, 6)), ((4, 7), (5, 6)), ((0, 3), (1, 2)))
_EDGES = ((4, 5), (5, 7), (6, 7), (4, 6),      # top cycle
          (0, 1), (0, 2), (1, 3), (2, 3),      # bottom cycle
          (0, 4), (1, 5), (2, 6), (3, 7))      # side cycle


class CubeTruss:
    """One cube block: the joint ids of its 8 corners (shared corners reuse the ids in ``usedDict``)."""

    def __init__(self, coordinate, usedDict=None):
        self._coord = coordinate
        self.jointIDs = [None] * 8
        self.GenerateNew({} if usedDict is None else usedDict)

    def __repr__(self):
        return str(self.jointIDs)

    def __getitem__(self, i):
        return self.jointIDs[i]

    def __setitem__(self, i, val):
        self.jointIDs[i] = val

    def GetCubeVertices(self):
        return [tuple(v + (corner >> axis & 1) for axis, v in enumerate(self._coord)) for corner in range(1 << len(self._coord))]

    def GenerateNew(self, usedDict=None):
        usedDict = {} if usedDict is None else usedDict
        nextID = max(usedDict.values()) + 1 if usedDict else 0
        for i, vertex in enumerate(self.GetCubeVertices()):
            if vertex not in usedDict:
                usedDict[vertex] = nextID
                nextID += 1
            self[i] = usedDict[vertex]

    def LinkMember(self, linkType, hasLinked):
        links = []

        def add(pair):
            link = [self[pair[0]], self[pair[1]]]
            if hasLinked is None:
                links.append(link)
            elif tuple(link) not in hasLinked:
                links.append(link)
                hasLinked.add(tuple(link))

        for first, second in _FACE_DIAGONALS:      # one (or both) diagonals per face
            pick = random.sample(range(3), k=1)[0] if linkType == LinkType.Random else linkType
            for pair in ((first,), (second,), (first, second))[pick]:
                add(pair)
        for pair in _EDGES:
            add(pair)
        return links


class CubeGrid:
    def __init__(self, xMax, yMax, zMax):
        self._shape = (xMax, yMax, zMax)
        self._usedDict = {}
        self.grid = [[[False] * zMax for _ in range(yMax)] for _ in range(xMax)]

    def __getitem__(self, coordinate):
        x, y, z = coordinate
        return self.grid[x][y][z]

    def __setitem__(self, coordinate, isUsed):
        x, y, z = coordinate
        self.grid[x][y][z] = isUsed

    def IsOutOfRange(self, coordinate):
        return any(not 0 <= v < m for v, m in zip(coordinate, self._shape))

    def GetRandomFeasible(self):
        xMax, yMax, zMax = self._shape
        return random.choice([(x, y, z) for z in range(zMax) for y in range(yMax) for x in range(xMax) if not self[(x, y, z)]])

    def GetNextFeasibles(self, coordinate, isSuffle=True):
        out = []
        for axis in range(3):
            for step in (-1, 1):
                nxt = tuple(v + (step if i == axis else 0) for i, v in enumerate(coordinate))
                if not self.IsOutOfRange(nxt) and not self[nxt]:
                    out.append(nxt)
        if isSuffle:
            random.shuffle(out)
        return out

    def RandomGenerateCubes(self, numCube=None, method=GenerateMethod.DFS):
        xMax, yMax, zMax = self._shape
        if numCube is None:
            numCube = random.randint(1, xMax * yMax * zMax)
        self._usedDict.clear()
        cubes, frontier = [], [self.GetRandomFeasible()]
        while len(cubes) < numCube and frontier:
            if method == GenerateMethod.DFS:
                coord = frontier.pop()
            elif method == GenerateMethod.BFS:
                coord = frontier.pop(0)
            else:
                coord = frontier.pop() if random.random() <= 0.5 else frontier.pop(0)
            self[coord] = True
            frontier.extend([c for c in self.GetNextFeasibles(coord) if c not in frontier])
            cubes.append(CubeTruss(coord, self._usedDict))
        return cubes

    def ProcessPinSupport(self, isAddPinSupport, length):
        minZ = min((z for _, _, z in self._usedDict), default=float("inf"))
        lx, ly, lz = (float(v) for v in length)
        joints = [None] * len(self._usedDict)
        for (x, y, z), jointID in self._usedDict.items():
            support = "PIN" if (isAddPinSupport and z == minZ) else "NO"
            joints[jointID] = [[float(x * lx), float(y * ly), float((z - minZ) * lz)], support]
        return joints

    def CubesToTruss(self, cubes, length, isAddPinSupport=True, isAllowParallel=True, linkType=LinkType.Random,
                     memberType=[1., 1e7, 0.1]):
        joints = self.ProcessPinSupport(isAddPinSupport, length)
        hasLinked = None if isAllowParallel else set()
        members = [[link, memberType] for cube in cubes for link in cube.LinkMember(linkType, hasLinked)]
        return {'joint': joints, 'force': {}, 'member': members}


# --------------------------------------------------------------------------- the generator
def _assign_random_forces(trussData, forceRange, nForceRange):
    free = [j for j, (_, support) in enumerate(trussData['joint']) if support == "NO"]
    lo, hi = (1, len(free)) if nForceRange is None else (1 if nForceRange[0] is None else nForceRange[0],
                                                           len(free) if nForceRange[1] is None else nForceRange[1])
    nForce = random.randint(lo, hi)
    trussData['force'] = [[j, [random.uniform(*forceRange[i]) for i in range(3)]] for j in sorted(random.sample(free, nForce))]
    return trussData


def _assign_random_member_types(trussData, memberTypes):
    for member in trussData['member']:
        choice = random.choice(memberTypes)
        member[1] = choice.Serialize() if isinstance(choice, MemberType) else choice
    return trussData


def GenerateRandomCubeTrusses(gridRange=(5, 5, 5), numCubeRange=(5, 5), numEachRange=(1, 10), lengthRange=(50, 150),
                              forceRange=[(-30000, 30000), (-30000, 30000), (-30000, 30000)], nForceRange=None,
                              method=GenerateMethod.Random, linkType=LinkType.Random, memberTypes=[[1., 1e7, 0.1]],
                              isAddPinSupport=True, isAllowParallel=False, isDoStructuralAnalysis=False, isPlotTruss=False,
                              isPrintMessage=True, saveFolder=None, augmenter=NoChange(), seed=None):
    if seed is not None:
        random.seed(seed)

    trussList, names = [], []
    for numCube in range(numCubeRange[0], numCubeRange[1] + 1):
        for case in range(numEachRange[0], numEachRange[1] + 1):
            while True:
                if isPrintMessage:
                    print(f"\rnumCube : {numCube :5d}, case : {case :5d}", end='')
                grid = CubeGrid(*gridRange)
                cubes = grid.RandomGenerateCubes(numCube, method)
                data = grid.CubesToTruss(cubes, [random.uniform(*lengthRange) for _ in range(3)], isAddPinSupport,
                                         isAllowParallel, linkType)
                _assign_random_forces(data, forceRange, nForceRange)
                _assign_random_member_types(data, memberTypes)
                truss = Truss(3).LoadFromJSON(data=augmenter(data))
                if truss.isStable:       # the counting rule Solve() applies first (truss.py:332-333)
                    break
                if isPrintMessage:
                    print("\nTruss is not stable. Re-genrating...\n")
            trussList.append(truss)
            names.append(f"cube-{numCube}_case_{case}")

    if isDoStructuralAnalysis:
        SolveBatch(trussList)            # one ragged GPU batch instead of a Solve() per truss
    if saveFolder is not None:
        for truss, name in zip(trussList, names):
            truss.DumpIntoJSON(os.path.join(saveFolder, name + ".json"))
    if isPlotTruss and isPrintMessage:
        print("\n[generate] isPlotTruss is not supported by the B200 build (plotting is outside the solve path).")
    return trussList


def GenerateAugmentedDataset(pool, nOut, moveToCentroid=False, translateRange=None, noiseMeans=None, noiseStds=None,
                             resetPin=None, seed=0, isDoStructuralAnalysis=True, asNumpy=True):
    """Bulk dataset generation on the GPU: ``pool`` (a list of Truss objects, e.g. from
    ``GenerateRandomCubeTrusses``) is expanded into ``nOut`` augmented trusses -- output truss ``o`` is pool truss
    ``o % len(pool)`` after MoveToCentroid, RandomTranslation(translateRange), AddJointNoise(noiseMeans, noiseStds) and
    RandomResetPin(*resetPin) (each optional, in this order: the order of the reference's recipe, example.py:239-267) --
    and solved in the same GPU pass (``tb_augment_ragged`` + ``tb_solve_ragged``).

    Returns the packed arrays (``joint_off``, ``member_off``, ``src``, ``xyz``, ``support``, ``conn``, ``aed``, ``force`` and,
    when solved, ``u``, ``ext``, ``axial``, ``weight``, ``info``; ``info[o] == -1`` marks a truss that fails the counting rule
    of truss.py:158-164, which the reference's generator would have re-drawn).  The host augmenter classes above follow
    Python's ``random`` stream; this path uses counter-based random numbers keyed by ``seed``."""
    from . import _lib
    from .batch import pack_ragged

    pool = list(pool)
    dim, jo, mo, xyz, sup, conn, aed, force = pack_ragged(pool)
    prm = _lib.TbAugmentParams()
    prm.move_to_centroid = 1 if moveToCentroid else 0
    if translateRange is not None:
        prm.random_translation, prm.translate_lo, prm.translate_hi = 1, float(translateRange[0]), float(translateRange[1])
    if noiseStds is not None or noiseMeans is not None:
        prm.joint_noise = 1
        means = [0., 0., 0.] if noiseMeans is None else list(noiseMeans)
        stds = [1., 1., 1.] if noiseStds is None else list(noiseStds)
        for i in range(dim):
            prm.noise_mean[i], prm.noise_std[i] = float(means[i]), float(stds[i])
    if resetPin is not None:
        minNumPin, maxNumPinRatio = resetPin
        if minNumPin < 3:
            from .utils import PinNotEnoughError
            raise PinNotEnoughError("Number of pins must >= 3.")
        prm.reset_pin, prm.min_pin = 1, int(minNumPin)
        prm.max_pin_ratio = 0.0 if maxNumPinRatio is None else float(maxNumPinRatio)
    prm.seed = int(seed)
    out = _lib.augment_and_solve_device(dim, {"joint_off": jo, "member_off": mo, "xyz": xyz, "support": sup, "conn": conn,
                                               "aed": aed, "force": force}, int(nOut), prm, solve=isDoStructuralAnalysis)
    out.pop("_keep", None)
    out["dim"] = dim
    if asNumpy:
        out = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in out.items()}
    return out


def GenerateRandomCubeTrussesOnDevice(nTruss, gridRange=(5, 5, 5), numCubeRange=(5, 5), lengthRange=(50, 150),
                                      forceRange=((-30000, 30000), (-30000, 30000), (-30000, 30000)), nForceRange=None,
                                      method=GenerateMethod.Random, linkType=LinkType.Random, memberTypes=((1., 1e7, 0.1),),
                                      isAddPinSupport=True, isAllowParallel=False, isDoStructuralAnalysis=False, seed=0,
                                      maxAttempts=64, asNumpy=True, export=False):
    """``GenerateRandomCubeTrusses`` as one GPU pass: ``nTruss`` random cube trusses -- the
    random walk over the grid, joint numbering, member linking, cell lengths, pin supports, random loads and member types
    and the stability re-draw of generate.py:152-336, 338-372 -- generated by ``tb_gencube`` (one thread per truss),
    compacted into the packed ragged layout and, if ``isDoStructuralAnalysis``, solved by ``tb_solve_ragged`` in the same
    pass.  ``numCubeRange`` is sampled uniformly per truss (the reference's generator loops over it with
    ``numEachRange`` trusses per value; call this once per value with ``numCubeRange=(k, k)`` for that).

    Returns the packed arrays (``joint_off``, ``member_off``, ``xyz``, ``support``, ``conn``, ``aed``, ``force``,
    ``gen_info`` [, ``u``, ``ext``, ``axial``, ``weight``, ``info``]) -- ``dataset.PackedDataset(3, arrays)`` gives the views
    into the reference's formats.  Random numbers are counter-based (keyed by ``seed``): reproducible, but not Python's
    ``random`` stream -- ``GenerateRandomCubeTrusses`` above keeps that."""
    import numpy as np

    from . import _lib

    prm = _lib.TbGencubeParams()
    for i in range(3):
        prm.grid[i] = int(gridRange[i])
        prm.force_lo[i], prm.force_hi[i] = float(forceRange[i][0]), float(forceRange[i][1])
    prm.ncube_lo, prm.ncube_hi = int(numCubeRange[0]), int(numCubeRange[1])
    prm.method, prm.link_type = int(method), int(linkType)
    prm.add_pin, prm.allow_parallel = int(bool(isAddPinSupport)), int(bool(isAllowParallel))
    prm.nforce_lo = -1 if nForceRange is None or nForceRange[0] is None else int(nForceRange[0])
    prm.nforce_hi = -1 if nForceRange is None or nForceRange[1] is None else int(nForceRange[1])
    table = np.array([t.Serialize() if isinstance(t, MemberType) else list(t) for t in memberTypes], dtype=np.float64).reshape(-1, 3)
    prm.n_type = int(table.shape[0])
    prm.max_attempts = int(maxAttempts)
    prm.length_lo, prm.length_hi = float(lengthRange[0]), float(lengthRange[1])
    prm.seed = int(seed)
    out = _lib.gencube_device(prm, nTruss, table, solve=isDoStructuralAnalysis, export=export)
    out["dim"] = 3
    if asNumpy:
        out = {k: (v.cpu().numpy() if hasattr(v, "cpu") else v) for k, v in out.items()}
    return out
