#!/bin/bash
# evidence run for profiles/: bench line, reference arm, ncu launch list of the same command, full captures of the
# kernels of the step (k_band2 at the bench batch, k_prep, k_recover), of k_band3 (one wave) and of k_band_subst
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --headline-only > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_band2 -s 4 -c 1 -f -o gpurun_out/prof_bench_k_band2 python bench.py --steps 3 --warmup 3 --headline-only > gpurun_out/ncu_full.log 2>&1; echo "ncu k_band2 rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_prep|k_recover' -s 8 -c 2 -f -o gpurun_out/prof_bench_stages python bench.py --steps 3 --warmup 3 --headline-only > gpurun_out/ncu_full2.log 2>&1; echo "ncu stages rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_band3 -s 9 -c 1 -f -o gpurun_out/prof_k_band3_b148 python tools/quick_time.py > gpurun_out/ncu_full3.log 2>&1; echo "ncu k_band3 rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_band_subst -s 3 -c 1 -f -o gpurun_out/prof_k_band_subst python tools/quick_time.py > gpurun_out/ncu_full4.log 2>&1; echo "ncu subst rc=$?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 600 python tools/quick_time.py > gpurun_out/quick_time.log 2>&1; echo "quick rc=$?"
ls -la gpurun_out
