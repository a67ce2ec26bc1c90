#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "shipped_goldens and (bar-10 or bar-120 or bar-6_)" > gpurun_out/sanitize.log 2>&1; echo "sanitize rc=$?"
grep -E "Invalid|ERROR SUMMARY|at .*\(|passed|failed" gpurun_out/sanitize.log | head -30
