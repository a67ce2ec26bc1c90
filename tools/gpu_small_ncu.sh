#!/bin/bash
# ncu captures of the small-system kernel (configs 3 and 4), the tiled kernel (config 5, B = 8) and the band kernel (full batch)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dense16 -s 3 -c 1 -f -o gpurun_out/prof_d16_cfg3 python tools/configs_time.py 3 > gpurun_out/ncu_cfg3.log 2>&1; echo "ncu cfg3 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_dense16 -s 2 -c 1 -f -o gpurun_out/prof_d16_cfg4 python tools/configs_time.py 4 > gpurun_out/ncu_cfg4.log 2>&1; echo "ncu cfg4 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_band_ts -s 2 -c 1 -f -o gpurun_out/prof_ts_b1024 python tools/ts_ncu.py 1024 3 > gpurun_out/ncu_ts_b1024.log 2>&1; echo "ncu ts rc=$?"
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read())
print(json.dumps({k:d.get(k) for k in ('value','ms_per_step','e2e','sustained','kernels')})[:1500])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --headline-only --sustain-s 0 > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
