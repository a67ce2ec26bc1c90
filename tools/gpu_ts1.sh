#!/bin/bash
# first GPU pass of the fused two-sided band kernel: memcheck on a tiny run, parity + timings, legacy timings, phases
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tail -1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "band and (bar-25 or bar-72 or bar-942)" --timeout 500 > gpurun_out/ts_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -25 gpurun_out/ts_memcheck.log
timeout 300 python tools/ts_dev.py 2>&1 | tee gpurun_out/ts_dev.log
TB_BAND_LEGACY=1 timeout 300 python tools/ts_dev.py 2>&1 | tee gpurun_out/ts_dev_legacy.log
timeout 300 python tools/ts_dev.py --phase 2>&1 | tee gpurun_out/ts_phase.log
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
