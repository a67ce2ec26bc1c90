#!/bin/bash
# bench only (headline + configs), brief print
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "pipelined" --timeout 300 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 3 $@ > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "import sys,json; d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1]); print(json.dumps({k:d.get(k) for k in ('value','ms_per_step','e2e','kernels','sustained')}, indent=1)); print({k:(v.get('value'),v.get('ms_per_step'),v.get('e2e',{}).get('value'), v.get('roofline',{}).get('frac')) for k,v in d.get('configs',{}).items()})"
tail -3 gpurun_out/bench.err
