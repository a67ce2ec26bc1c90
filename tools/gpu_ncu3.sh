#!/bin/bash
# full ncu capture of one band kernel at B=1024 (quick_time's third bar-942 batch); $1 = kernel regex, $2 = TB_BAND_WARPS
mkdir -p gpurun_out
K=${1:-k_band2}
TB_BAND_WARPS=${2:-0} timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$K -s 14 -c 1 -f -o gpurun_out/prof_$K python tools/quick_time.py > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/ncu_full.log
