#!/bin/bash
# iteration loop for the band kernels: parity tests, quick timings for each kernel variant, phase cycles
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
for w in 0 3 2; do
  echo "== TB_BAND_WARPS=$w"
  TB_BAND_DEBUG=1 TB_BAND_WARPS=$w timeout 600 python tools/quick_time.py 2>&1 | grep -E "bar-942|launch_band" | tee -a gpurun_out/quick_time_w$w.log
done
timeout 900 python tools/band_phase.py 1 1024 2>&1 | tee gpurun_out/band_phase.log
