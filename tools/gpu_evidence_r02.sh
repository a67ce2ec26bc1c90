#!/bin/bash
# evidence run of round 2 for profiles/: smoke, GPU tests, bench line, reference arm, ncu launch list of the headline
# step, full captures of the band kernel (bench batch), the stage kernels, k_dense16 (configs 3 and 4), k_chol (config 5)
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --headline-only --sustain-s 0 > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_band_ts -s 2 -c 1 -f -o gpurun_out/prof_ts_b1024 python tools/ts_ncu.py 1024 3 > gpurun_out/ncu_ts_b1024.log 2>&1; echo "ncu ts rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_prep|k_recover' -s 4 -c 2 -f -o gpurun_out/prof_stages_b1024 python tools/ts_ncu.py 1024 3 > gpurun_out/ncu_stages.log 2>&1; echo "ncu stages rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dense16 -s 3 -c 1 -f -o gpurun_out/prof_d16_cfg3 python tools/configs_time.py 3 > gpurun_out/ncu_cfg3.log 2>&1; echo "ncu cfg3 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dense16 -s 2 -c 1 -f -o gpurun_out/prof_d16_cfg4 python tools/configs_time.py 4 > gpurun_out/ncu_cfg4.log 2>&1; echo "ncu cfg4 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_chol -s 1 -c 1 -f -o gpurun_out/prof_chol_cfg5 python tools/configs_time.py 5 > gpurun_out/ncu_cfg5.log 2>&1; echo "ncu cfg5 rc=$?"
timeout 600 python tools/quick_time.py > gpurun_out/quick_time.log 2>&1; echo "quick rc=$?"
python tools/bench_brief.py gpurun_out/bench_n1.json 2>/dev/null | head -40
