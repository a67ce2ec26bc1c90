#!/bin/bash
# evidence run for profiles/: bench line, ncu launch list of the same command, full capture of the dominant kernel
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_band2 -s 4 -c 1 -f -o gpurun_out/prof_bench_k_band2 python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 1200 ncu --set full --clock-control none -k regex:'k_prep|k_recover' -s 8 -c 2 -f -o gpurun_out/prof_bench_stages python bench.py --steps 3 --warmup 3 > gpurun_out/ncu_full2.log 2>&1; echo "ncu stages rc=$?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
