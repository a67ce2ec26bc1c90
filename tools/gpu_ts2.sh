#!/bin/bash
# iteration pass of the two-sided band kernel: memcheck on a small run, parity + timings, phases, one ncu capture
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "band and (bar-25 or bar-72 or bar-942 or bar-10)" --timeout 500 > gpurun_out/ts_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -5 gpurun_out/ts_memcheck.log
timeout 300 python tools/ts_dev.py 2>&1 | tee gpurun_out/ts_dev.log
timeout 300 python tools/ts_dev.py --phase 2>&1 | grep -v "^bar-942 x\(2048\|8192\)" | tee gpurun_out/ts_phase.log
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_band_ts -s 2 -c 1 -f -o gpurun_out/prof_ts_b1024 python tools/ts_ncu.py 1024 3 > gpurun_out/ncu_ts_b1024.log 2>&1; echo "ncu rc=$?"
