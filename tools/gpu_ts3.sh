#!/bin/bash
# experiment pass of the two-sided band kernel: band parity tests, parity + timings, phases, optional ncu capture ($1 = ncu)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "band or bar-942 or bitwise" --timeout 500 > gpurun_out/pytest_band.log 2>&1; echo "pytest band rc=$?"
tail -4 gpurun_out/pytest_band.log
timeout 300 python tools/ts_dev.py 2>&1 | grep -v "x\(2048\|8192\)" | tee gpurun_out/ts_dev.log
timeout 300 python tools/ts_dev.py --phase 2>&1 | grep -v "^bar-942 x\(2048\|8192\)" > gpurun_out/ts_phase.log; grep -A12 "x1024" gpurun_out/ts_phase.log | head -14
if [ "$1" == "ncu" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_band_ts -s 2 -c 1 -f -o gpurun_out/prof_ts_b1024 python tools/ts_ncu.py 1024 3 > gpurun_out/ncu_ts_b1024.log 2>&1; echo "ncu rc=$?"
fi
