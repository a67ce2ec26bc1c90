#!/bin/bash
mkdir -p gpurun_out
for s in 36 38 40 42 44; do
  TB_TS_SPLIT=$s timeout 200 python tools/ts_dev.py 2>&1 | grep -E "program|x1024|x1:" | sed "s/^/split $s: /" | cut -c1-260
done | tee gpurun_out/ts_split_sweep.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -5 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read())
print(json.dumps({k:d.get(k) for k in ('value','ms_per_step','e2e','roofline','kernels','gpu_launches','wall_s_timed_region')}, indent=1)[:3000])
for k,v in d.get('configs',{}).items(): print(k, json.dumps(v)[:1500])
PY
