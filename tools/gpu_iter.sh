#!/bin/bash
# iteration loop: parity tests for the blocked path + timings (+ optional bench)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python tools/quick_time.py 2>&1 | tee gpurun_out/quick_time.log
if [ "$1" == "bench" ]; then
  timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
  python -c "import sys,json; d=json.loads(open('gpurun_out/bench.json').read()); print(json.dumps({k:d[k] for k in ('value','ms_per_step','e2e','roofline','kernels','clocks','gpu_launches')}, indent=1))"
  tail -3 gpurun_out/bench.err
fi
