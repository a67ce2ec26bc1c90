#!/bin/bash
mkdir -p gpurun_out
K=${1:-k_chol}
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_$K python tools/quick_time.py > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/ncu_full.log
