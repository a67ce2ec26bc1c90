"""GA generation step on the device (csrc/tb_ga.cu) against the reference's operator definitions (slientruss3d/ga.py:151-190).

The device path draws its random numbers from a counter-based generator, so it cannot replay Python's ``random`` stream
(the host ``GA.Evolve`` does, see test_host_api_cpu.py); what is checked here is what the operators guarantee whatever the
random numbers: the ranking is ``sorted(..., key=fitness)`` bit for bit, the elites come first in rank order, every other
child is explained by exactly the reference's four rules with the reference's probabilities, runs are reproducible from the
seed, and the loop's bookkeeping (best-fitness history, feasible record, return value) follows ``Evolve``."""
import ctypes as C
import random

import numpy as np
import pytest

from python_stable_3d_truss_analysis_b200 import _lib
from python_stable_3d_truss_analysis_b200.ga import GA
from python_stable_3d_truss_analysis_b200.truss import Truss
from python_stable_3d_truss_analysis_b200.type import MemberType
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _step(nPop, nElite, M, T, pc, pm, po, seed, generation, fitness, flags, genes):
    import torch
    dev = torch.device("cuda:0")
    prm = _lib.TbGaParams(nPop, nElite, M, T, pc, pm, po, seed)
    g_in = torch.from_numpy(genes.astype(np.int32)).to(dev)
    g_out = torch.empty_like(g_in)
    order = torch.empty(nPop, dtype=torch.int32, device=dev)
    rep = torch.zeros(C.sizeof(_lib.TbGaReport), dtype=torch.uint8, device=dev)
    _lib.ga_step(prm, generation, torch.from_numpy(fitness).to(dev), torch.from_numpy(flags).to(dev), g_in, g_out, order, rep)
    torch.cuda.synchronize()
    return g_out.cpu().numpy(), order.cpu().numpy(), _lib.TbGaReport.from_buffer_copy(rep.cpu().numpy().tobytes())


def _classify(child, j, genes, elites):
    """Which of the reference's rules explains child j (a child may satisfy several; first match in the order below)."""
    if np.array_equal(child, genes[j]):
        return "origin"
    d = (elites != child[None, :]).sum(axis=1)
    if (d == 1).any():
        return "mutate"
    M = child.shape[0]
    eq = elites == child[None, :]                              # [E, M]
    for a in np.nonzero(eq[:, 0] | eq[:, M - 1])[0]:           # gene0 provides both ends (cut0 > 0 or cut1 < M) or ...
        diff = np.nonzero(~eq[a])[0]
        if diff.size == 0:
            return "cross"
        lo, hi = diff[0], diff[-1] + 1                         # child differs from gene0 only inside [lo, hi)
        if (eq[:, lo:hi].all(axis=1)).any():
            return "cross"
    for b in range(elites.shape[0]):                           # ... the segment from gene1 reaches an end
        diff = np.nonzero(~eq[b])[0]
        if diff.size and (diff[0] > 0 and diff[-1] < M - 1):
            continue
        rest = np.nonzero(~eq[b])[0]
        if rest.size and (eq[:, rest].all(axis=1)).any() and (np.diff(rest) == 1).all():
            return "cross"
    return "random"


def test_rank_is_the_stable_sort_and_elites_come_first():
    rng = np.random.default_rng(0)
    nPop, nElite, M, T = 1000, 100, 24, 5
    fitness = np.round(rng.uniform(0, 50, size=nPop))          # many ties: stability matters
    fitness[rng.integers(0, nPop, size=20)] = np.inf           # failed systems sort last
    flags = rng.integers(0, 2, size=(nPop, 2)).astype(np.uint8)
    genes = rng.integers(0, T, size=(nPop, M))
    out, order, rep = _step(nPop, nElite, M, T, 0.7, 0.1, 0.1, 1234, 0, fitness, flags, genes)
    want = np.array(sorted(range(nPop), key=lambda i: fitness[i]))          # GA.Select, ga.py:157
    assert np.array_equal(order, want)
    assert np.array_equal(out[:nElite], genes[want[:nElite]])               # newPop[:nElite] = elitePop
    assert rep.best_index == want[0] and rep.best_fitness == fitness[want[0]]
    feas = [i for i in want if flags[i, 0] and flags[i, 1]]
    assert rep.feasible_index == feas[0] and rep.feasible_fitness == fitness[feas[0]]


def test_children_follow_the_reference_rules_with_the_reference_probabilities():
    rng = np.random.default_rng(1)
    nPop, nElite, M, T = 8192, 64, 72, 20
    pc, pm, po = 0.6, 0.15, 0.1
    fitness = rng.uniform(0, 1, size=nPop)
    flags = np.ones((nPop, 2), np.uint8)
    genes = rng.integers(0, T, size=(nPop, M))
    out, order, _ = _step(nPop, nElite, M, T, pc, pm, po, 99, 3, fitness, flags, genes)
    elites = genes[order[:nElite]]
    kinds = [_classify(out[j], j, genes, elites) for j in range(nElite, nPop)]
    n = len(kinds)
    frac = {k: kinds.count(k) / n for k in ("cross", "mutate", "origin", "random")}
    # a crossover whose segment changes nothing looks like "origin"-free copy of an elite and is counted as "cross" above;
    # binomial standard deviation at n = 8128 is < 0.6 %: 3 % is a 5-sigma band
    assert abs(frac["cross"] - pc) < 0.03 and abs(frac["mutate"] - pm) < 0.03, frac
    assert abs(frac["origin"] - po) < 0.03 and abs(frac["random"] - (1 - pc - pm - po)) < 0.03, frac
    assert out.min() >= 0 and out.max() < T
    # mutation always changes the type (ga.py:170) and crossover parents are distinct elites (random.sample k=2)
    again, _, _ = _step(nPop, nElite, M, T, pc, pm, po, 99, 3, fitness, flags, genes)
    other, _, _ = _step(nPop, nElite, M, T, pc, pm, po, 100, 3, fitness, flags, genes)
    assert np.array_equal(out, again) and not np.array_equal(out, other)    # reproducible from (seed, generation)


def test_evolve_on_device_bar72():
    random.seed(0)
    types = [MemberType(i, random.uniform(1e7, 3e7), random.uniform(0.1, 1.0)) for i in range(1, 21)]
    t = Truss(3).LoadFromJSON(f"{H.GOLDEN}/ref_data/bar-72_input_0.json")
    ga = GA(t, types, allowStress=30000., allowDisplace=10., nIteration=30, nPatience=50, nPop=2048, nElite=256)
    gene, info, pop, hist = ga.EvolveOnDevice(isPrintMessage=False, seed=7)
    assert len(pop) == 2048 and len(hist) == 30 and hist == sorted(hist, reverse=True)
    assert info[1] and info[2] and len(gene) == t.nMember
    fit, ok_s, ok_d = ga.GetFitness(gene)                       # the returned gene re-evaluated through the host path
    assert abs(fit - info[0]) <= 1e-9 * abs(fit) and ok_s and ok_d
    assert fit <= hist[0]                                       # 30 generations do not end worse than the first ranking
    ga2 = GA(t, types, allowStress=30000., allowDisplace=10., nIteration=30, nPatience=50, nPop=2048, nElite=256)
    gene2, info2, _, hist2 = ga2.EvolveOnDevice(isPrintMessage=False, seed=7)
    assert gene2 == gene and hist2 == hist                      # same seed, same run
    # early stopping returns the recorded feasible gene
    ga3 = GA(t, types, allowStress=30000., allowDisplace=10., nIteration=None, nPatience=3, nPop=512, nElite=64)
    gene3, info3, pop3, hist3 = ga3.EvolveOnDevice(isPrintMessage=False, seed=11)
    assert len(hist3) >= 1 and info3[1] and info3[2]
