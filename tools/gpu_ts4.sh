#!/bin/bash
# A/B of the entry schedule: timings with the bank-aware order and with the scatter map's order
mkdir -p gpurun_out
timeout 300 python tools/ts_dev.py 2>&1 | grep -v "x\(2048\|8192\)" | tee gpurun_out/ts_dev.log
TB_TS_PLAIN_ORDER=1 timeout 300 python tools/ts_dev.py 2>&1 | grep -v "x\(2048\|8192\)" | tee gpurun_out/ts_dev_plain.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_prep -s 2 -c 1 -f -o gpurun_out/prof_prep_b1024 python tools/ts_ncu.py 1024 3 > gpurun_out/ncu_prep_b1024.log 2>&1; echo "ncu rc=$?"
