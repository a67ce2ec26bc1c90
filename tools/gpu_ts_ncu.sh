#!/bin/bash
# full ncu capture of the fused band kernel at two batch sizes (source-level stalls)
mkdir -p gpurun_out
for B in 148 1024; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_band_ts -s 2 -c 1 -f -o gpurun_out/prof_ts_b$B python tools/ts_ncu.py $B 3 > gpurun_out/ncu_ts_b$B.log 2>&1; echo "ncu B=$B rc=$?"
done
ls -la gpurun_out/*.ncu-rep
