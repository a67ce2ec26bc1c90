#!/bin/bash
# compute-sanitizer over the kernels added this round (small cases): memcheck, then racecheck on the band kernels
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
timeout 1500 $S --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_ga_device.py tests/test_gpu_generate_device.py -m gpu -q -x -k "rank_is or no_augmentation or centroid" > gpurun_out/san_mem1.log 2>&1; echo "memcheck ga/augment rc=$?"
tail -3 gpurun_out/san_mem1.log
timeout 1500 $S --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "load_cases_bar942 or every_band_width or long_tower" > gpurun_out/san_mem2.log 2>&1; echo "memcheck band/subst rc=$?"
tail -3 gpurun_out/san_mem2.log
timeout 1500 $S --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "load_cases_bar942" > gpurun_out/san_race.log 2>&1; echo "racecheck band rc=$?"
tail -5 gpurun_out/san_race.log
