#!/bin/bash
# iteration loop: parity tests + quick timings (+ optional bench / ncu of one kernel via $1, $2)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python tools/quick_time.py 2>&1 | tee gpurun_out/quick_time.log
if [ "$1" == "bench" ]; then
  timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
  python -c "import sys,json; d=json.loads(open('gpurun_out/bench.json').read()); print(json.dumps({k:d.get(k) for k in ('value','ms_per_step','e2e','roofline','roofline_stages','kernels','clocks','gpu_launches')}, indent=1))"
  tail -3 gpurun_out/bench.err
fi
if [ -n "$2" ]; then
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$2 -s 6 -c 1 -f -o gpurun_out/prof_$2 python tools/quick_time.py > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
