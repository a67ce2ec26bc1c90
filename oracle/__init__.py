"""CPU oracle for the batched direct-stiffness hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker or the timed
CPU baseline.  The product path (``python_stable_3d_truss_analysis_b200``)
never imports this package and fails loudly when its CUDA library is missing.
"""
