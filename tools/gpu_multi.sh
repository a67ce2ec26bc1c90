#!/bin/bash
# multi-GPU bench lines: usage gpu_multi.sh N [extra env]
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"
tail -4 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read())
print(json.dumps({k:d.get(k) for k in ('value','ms_per_step','n_gpus','e2e','kernels')}, indent=None)[:1200])
print(d.get('run'))
for k,v in d.get('configs',{}).items(): print(k, {kk:v.get(kk) for kk in ('value','ms_per_step','error','multi_gpu')})
PY
