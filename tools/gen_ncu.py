"""One pass of the device dataset generators and the GA step (for ncu captures).  usage: python tools/gen_ncu.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import copy, torch
from python_stable_3d_truss_analysis_b200 import generate as G
from python_stable_3d_truss_analysis_b200.truss import Truss
from tests import helpers as H
for _ in range(2):
    G.GenerateRandomCubeTrussesOnDevice(65536, (5, 5, 5), (7, 7), isDoStructuralAnalysis=True, seed=1, asNumpy=False)
pool = [Truss(3).LoadFromJSON(data=copy.deepcopy(d)) for _, _, d, _ in H.cube7_shipped()]
for _ in range(2):
    G.GenerateAugmentedDataset(pool, 65536, moveToCentroid=True, translateRange=(-5, 5), noiseStds=[1, 1, 1], resetPin=(3, 0.5), seed=3, asNumpy=False)
torch.cuda.synchronize()
print("done")
