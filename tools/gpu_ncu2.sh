#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_chol -s 2 -c 1 -f -o gpurun_out/prof_chol_b888 python tools/sweep_b.py 888 > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"
