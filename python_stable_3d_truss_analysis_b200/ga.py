"""Member-type selection by genetic algorithm, with the population's fitness evaluated as ONE
batched GPU call.

Same constructor, hooks and return values as the reference's ``slientruss3d/ga.py`` (``GA :12-237``;
documented in ``detail/truss_optimization.md``).  The reference evaluates ``GetFitness`` gene by gene
(``Select :155-160``, ``GetBestFeasibleGene :110-123``), each call mutating the one shared ``Truss``
and running a full ``Solve()``; here ``Select`` / ``GetBestFeasibleGene`` send the whole gene matrix
through ``tb_fitness`` (``FitnessBatch``).  Fitness evaluation consumes no random numbers, so the
``random`` stream -- and with it the evolution trajectory -- is the one the reference would follow.
Subclasses that override ``GetFitness`` keep working: the batched path is only used while
``GetFitness`` is the stock one.
"""
from __future__ import annotations

import random

from .batch import FitnessBatch
from .truss import Truss
from .type import MemberType
from .utils import (EliteNumberTooMuchError, INF, InfinteLoop, MinDisplaceTooLargeError, MinStressTooLargeError,
                    OnlyOneMemberTypeError, ProbabilityGreaterThanOneError)


class GA:
    """Search the best combination of member types for the members of a truss."""

    def __init__(self, truss: Truss, memberTypeList: list[MemberType], allowStress: float = 30000.,
                 allowDisplace: float = 10., nIteration: int = None, nPatience: int = 50, nPop: int = 200,
                 nElite: int = 50, pCrossover: float = 0.7, pMutate: float = 0.1, pOrigin: float = 0.1,
                 isCheckWorst: bool = False):
        self.nPop, self.nElite = nPop, nElite
        self.pCrossover, self.pMutate, self.pOrigin = pCrossover, pMutate, pOrigin
        self.pRandomGene = 1. - pCrossover - pMutate - pOrigin
        self.nIteration, self.nPatience = nIteration, nPatience

        self.truss = truss
        self.allowStress, self.allowDisplace = allowStress, allowDisplace
        self.typeList = memberTypeList
        self.nMember = truss.nMember
        self.nType = len(memberTypeList)
        self.memberIDList = truss.GetMemberIDs()
        self.memberIDMap = dict(enumerate(self.memberIDList))

        self._lastFeasibleGene = [None] * self.nMember
        self._lastFeasibleFitness = None
        self.CheckRatioality(isCheckWorst)

    # ------------------------------------------------------------------ configuration checks
    @property
    def memberTypeWeightedInitProb(self):
        return [1.] * len(self.typeList)

    def CheckRatioality(self, isCheckWorst):
        if self.nElite > self.nPop:
            raise EliteNumberTooMuchError(f"Number of elites must <= number of population. Got [nElite] = {self.nElite}, [nPop] = {self.nPop}.")
        total = self.pCrossover + self.pMutate + self.pOrigin
        if total > 1.:
            raise ProbabilityGreaterThanOneError(f"[pCrossover] + [pMutate] + [pOrigin] must <= 1.0, but got [{total :.4f}].")
        if self.nType <= 1:
            raise OnlyOneMemberTypeError(f"Number of member types must >= 2, but got {self.nType}.")
        if not isCheckWorst:
            return
        # The strongest choices: largest area bounds the stress, largest E*A bounds the displacement.
        by_area = max(self.typeList, key=lambda t: t.a)
        by_ea = max(self.typeList, key=lambda t: t.e * t.a)
        for probe, check, limit, err in (
                (by_area, self.truss.IsInternalStressAllowed, self.allowStress,
                 MinStressTooLargeError("Minimum stress is too large. Need other member types which have more [A] value.")),
                (by_ea, self.truss.IsDisplacementAllowed, self.allowDisplace,
                 MinDisplaceTooLargeError("Minimum displacement is too large. Need other member types which have more [E*A] value."))):
            for memberID in self.memberIDList:
                self.truss.SetMemberType(memberID, probe)
            self.truss.Solve()
            if not check(limit)[0]:
                raise err

    # ------------------------------------------------------------------ genes
    def TranslateGene(self, gene):
        return {self.memberIDMap[i]: self.typeList[locus] for i, locus in enumerate(gene)}

    def GetRandomGene(self):
        return random.choices(range(self.nType), k=self.nMember)

    def SetMemberTypesByGene(self, gene, truss):
        for i, locus in enumerate(gene):
            truss.SetMemberType(self.memberIDMap[i], self.typeList[locus])
        return truss

    def Initialize(self):
        weights = self.memberTypeWeightedInitProb
        return [random.choices(range(self.nType), k=self.nMember, weights=weights) for _ in range(self.nPop)]

    # ------------------------------------------------------------------ fitness
    def _stock_fitness(self):
        return type(self).GetFitness is GA.GetFitness

    def GetFitnessBatch(self, pop):
        """[(fitness, isInternalAllowed, isDisplaceAllowed)] for every gene of ``pop`` -- one GPU call."""
        if not pop:
            return []
        out = FitnessBatch(self.truss, pop, self.typeList, self.allowStress, self.allowDisplace)
        bad = out["info"].nonzero()[0]
        if bad.size:
            from .truss import raise_for_info
            raise_for_info(int(out["info"][bad[0]]))
        return [(float(f), bool(s), bool(d)) for f, (s, d) in zip(out["fitness"], out["flags"])]

    def GetFitness(self, gene):
        """Single-gene fitness; like the reference it leaves ``self.truss`` set to and solved for the gene."""
        truss = self.SetMemberTypesByGene(gene, self.truss)
        out = FitnessBatch(truss, [gene], self.typeList, self.allowStress, self.allowDisplace, full=True)
        from .truss import raise_for_info
        raise_for_info(int(out["info"][0]))
        truss._set_dense_results(out["u"][0], out["ext"][0], out["axial"][0])
        return float(out["fitness"][0]), bool(out["flags"][0][0]), bool(out["flags"][0][1])

    def _evaluate(self, pop):
        if self._stock_fitness():
            return self.GetFitnessBatch(pop)
        return [self.GetFitness(gene) for gene in pop]

    def _RecordFeasible(self, evaluatedPop, isSorted=False):
        for gene, (fitness, okStress, okDisplace) in evaluatedPop:
            if okStress and okDisplace and (self._lastFeasibleFitness is None or fitness < self._lastFeasibleFitness):
                self._lastFeasibleGene[:], self._lastFeasibleFitness = gene, fitness
                if isSorted:
                    break

    def GetBestFeasibleGene(self, pop, isDirectlyReturnRecord=False):
        if isDirectlyReturnRecord and self._lastFeasibleFitness is not None:
            return self._lastFeasibleGene, (self._lastFeasibleFitness, True, True)
        best, bestGene = INF, None
        for gene, (fitness, okStress, okDisplace) in zip(pop, self._evaluate(pop)):
            if okStress and okDisplace and fitness < best:
                best, bestGene = fitness, gene
        if bestGene is None:
            if self._lastFeasibleFitness is not None:
                return self._lastFeasibleGene, (self._lastFeasibleFitness, True, True)
            return None, (INF, False, False)
        return bestGene, (best, True, True)

    # ------------------------------------------------------------------ evolution operators
    def Select(self, pop, isRecordFeasible=False):
        ranked = sorted(([gene, info] for gene, info in zip(pop, self._evaluate(pop))), key=lambda x: x[1][0])
        elitePop = [gene for gene, _ in ranked[:self.nElite]]
        if isRecordFeasible:
            self._RecordFeasible(ranked, isSorted=True)
        return elitePop, ranked[0][1]

    def Crossover(self, gene0, gene1):
        lo, hi = sorted(random.sample(range(self.nMember), k=2))
        return [gene0[i] if i < lo or i >= hi else gene1[i] for i in range(self.nMember)]

    def Mutate(self, gene):
        gene = gene.copy()
        at = random.randint(0, self.nMember - 1)
        gene[at] = random.choice([t for t in range(self.nType) if t != gene[at]])
        return gene

    def UpdatePop(self, pop, elitePop):
        edgeCross = self.pCrossover
        edgeMutate = edgeCross + self.pMutate
        edgeOrigin = edgeMutate + self.pOrigin
        newPop = list(elitePop) + [None] * (self.nPop - self.nElite)
        for j in range(self.nElite, self.nPop):
            p = random.random()
            if p <= edgeCross:
                newPop[j] = self.Crossover(*random.sample(elitePop, k=2))
            elif p <= edgeMutate:
                newPop[j] = self.Mutate(random.choice(elitePop))
            elif p <= edgeOrigin:
                newPop[j] = pop[j]
            else:
                newPop[j] = self.GetRandomGene()
        return newPop

    def Evolve(self, isPrintMessage=True):
        pop = self.Initialize()
        bestFitness, history, nWait, earlyStop = INF, [], 0, False
        for i in (range(self.nIteration) if self.nIteration is not None else InfinteLoop()):
            elitePop, (minFitness, okStress, okDisplace) = self.Select(pop, True)
            if minFitness < bestFitness:
                bestFitness, nWait = minFitness, 0
            else:
                nWait += 1
                if nWait >= self.nPatience:
                    earlyStop = True
                    break
            history.append(bestFitness)
            if isPrintMessage:
                print(f"\rIteration: {i :6d}, nWaitBestIter: {nWait :3d}, minFitness: {minFitness :12.4f}, "
                      f"isInternalAllowed: {str(okStress) :5s}, isDisplaceAllowed: {str(okDisplace) :5s}", end='')
            pop = self.UpdatePop(pop, elitePop)

        if isPrintMessage:
            print('...Early stoping !' if earlyStop else "")

        minGene, minGeneInfo = self.GetBestFeasibleGene(pop, earlyStop)
        if minGene is None:
            minGene = pop[0]
            minGeneInfo = self.GetFitness(minGene)
            if isPrintMessage:
                print('-' * 50 + '\n' + "Warning: Cannot find any feasible result, so only return the gene which has lowest fitness." + '\n' + '-' * 50)
        return minGene, minGeneInfo, pop, history

    # ------------------------------------------------------------------ the whole loop on the device
    def EvolveOnDevice(self, isPrintMessage=True, seed=None, device=None):
        """``Evolve`` with the population resident on the GPU: every generation is one ``tb_fitness`` launch on the gene
        matrix plus one ``tb_ga_step`` (ranking, elitism, crossover, mutation, re-seeding -- the operators of
        ``UpdatePop`` / ``Crossover`` / ``Mutate`` / ``Select``, ga.py:155-190, written as CUDA kernels); the host reads one
        40-byte report per generation for the early-stopping rule.  Same return value as ``Evolve``.

        The random numbers come from a counter-based generator on the device (reproducible from ``seed``), not from
        Python's ``random`` stream, so the trajectory differs from ``Evolve``'s while the operators and their
        probabilities are the same.  Needs the stock ``GetFitness`` (a Python override cannot run on the device)."""
        import ctypes as C

        import numpy as np
        import torch

        from . import _lib
        from .batch import type_table

        if not self._stock_fitness():
            raise TypeError("EvolveOnDevice needs the stock GetFitness; use Evolve() with an overridden fitness")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        seed = random.getrandbits(63) if seed is None else int(seed)
        nPop, M = self.nPop, self.nMember
        prm = _lib.TbGaParams(nPop, self.nElite, M, self.nType, self.pCrossover, self.pMutate, self.pOrigin, seed)
        xyz, support, conn, _, force = self.truss._pack()
        plan = self.truss._get_plan(support, conn)
        td = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
        d_xyz, d_force, d_tab = td(xyz), td(force), td(type_table(self.typeList))
        w = np.asarray(self.memberTypeWeightedInitProb, dtype=np.float64)
        d_cum = None if np.all(w == w[0]) else td(np.cumsum(w) / w.sum())
        genes = [torch.empty((nPop, M), dtype=torch.int32, device=dev) for _ in range(2)]
        out = {"fitness": torch.empty(nPop, dtype=torch.float64, device=dev),
               "flags": torch.empty((nPop, 2), dtype=torch.uint8, device=dev),
               "info": torch.empty(nPop, dtype=torch.int32, device=dev)}
        order = torch.empty(nPop, dtype=torch.int32, device=dev)
        rep_dev = torch.zeros(C.sizeof(_lib.TbGaReport), dtype=torch.uint8, device=dev)
        rep_host = torch.zeros(C.sizeof(_lib.TbGaReport), dtype=torch.uint8).pin_memory()

        def evaluate_and_rank(cur, nxt, generation):
            plan.fitness_device(nPop, d_xyz, d_force, genes[cur], d_tab, self.allowStress, self.allowDisplace, out)
            _lib.ga_step(prm, generation, out["fitness"], out["flags"], genes[cur], None if nxt is None else genes[nxt],
                         order, rep_dev)
            rep_host.copy_(rep_dev, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return _lib.TbGaReport.from_buffer_copy(rep_host.numpy().tobytes())

        def record_feasible(rep, cur):
            if rep.feasible_index >= 0 and (self._lastFeasibleFitness is None or rep.feasible_fitness < self._lastFeasibleFitness):
                self._lastFeasibleGene[:] = genes[cur][rep.feasible_index].tolist()
                self._lastFeasibleFitness = float(rep.feasible_fitness)

        _lib.ga_init(prm, genes[0], d_cum)
        cur = 0
        bestFitness, history, nWait, earlyStop = INF, [], 0, False
        for i in (range(self.nIteration) if self.nIteration is not None else InfinteLoop()):
            rep = evaluate_and_rank(cur, 1 - cur, i)          # Select + UpdatePop of generation i
            if i == 0 and bool(out["info"].any().item()):     # topology / input problems show in the first generation
                from .truss import raise_for_info             # (a gene that makes K singular ranks last: fitness inf)
                raise_for_info(int(out["info"][out["info"].nonzero()[0, 0]].item()))
            record_feasible(rep, cur)
            minFitness = float(rep.best_fitness)
            if minFitness < bestFitness:
                bestFitness, nWait = minFitness, 0
            else:
                nWait += 1
                if nWait >= self.nPatience:
                    earlyStop = True
                    break
            history.append(bestFitness)
            if isPrintMessage:
                print(f"\rIteration: {i :6d}, nWaitBestIter: {nWait :3d}, minFitness: {minFitness :12.4f}, "
                      f"isInternalAllowed: {str(bool(rep.best_stress_ok)) :5s}, isDisplaceAllowed: {str(bool(rep.best_displace_ok)) :5s}", end='')
            cur = 1 - cur                                      # the updated population becomes the current one
        if isPrintMessage:
            print('...Early stoping !' if earlyStop else "")

        # GetBestFeasibleGene (ga.py:110-123): the record when early-stopped, else the best feasible gene of the final population
        pop = genes[cur].cpu().numpy().tolist()
        if earlyStop and self._lastFeasibleFitness is not None:
            return self._lastFeasibleGene, (self._lastFeasibleFitness, True, True), pop, history
        rep = evaluate_and_rank(cur, None, 0)
        if rep.feasible_index >= 0:
            return pop[rep.feasible_index], (float(rep.feasible_fitness), True, True), pop, history
        if self._lastFeasibleFitness is not None:
            return self._lastFeasibleGene, (self._lastFeasibleFitness, True, True), pop, history
        if isPrintMessage:
            print('-' * 50 + '\n' + "Warning: Cannot find any feasible result, so only return the gene which has lowest fitness." + '\n' + '-' * 50)
        return pop[0], self.GetFitness(pop[0]), pop, history
