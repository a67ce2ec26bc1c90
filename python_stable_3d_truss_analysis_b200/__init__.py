"""B200-native batched direct-stiffness truss solver behind the slientruss3d API.

    from python_stable_3d_truss_analysis_b200.truss import Truss, Member
    from python_stable_3d_truss_analysis_b200.type  import MemberType, SupportType
    from python_stable_3d_truss_analysis_b200.ga    import GA
    from python_stable_3d_truss_analysis_b200.generate import GenerateRandomCubeTrusses
    from python_stable_3d_truss_analysis_b200.batch import SolveBatch, SolveLoadCases, FitnessBatch

The arithmetic lives in ``csrc/libtruss_b200.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/truss_b200.h``); this package is the host-side mirror of the reference's Python API.
"""
__version__ = "0.1.0"
