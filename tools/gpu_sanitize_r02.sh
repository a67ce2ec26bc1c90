#!/bin/bash
# compute-sanitizer over the kernels added / changed in round 2: memcheck (k_band_ts, k_chol with the TMA pipeline,
# k_compact_out, k_gencube / k_gencube_pack, the pipelined host path) and racecheck (shared-memory hazards of k_band_ts, k_chol)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gencube_device.py -q -x --timeout 1400 \
  -k "compact or pipelined or fixed_member or (band and bar-942) or (tiled and bar-72) or (replay and True) or solved_in_one_pass" > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -4 gpurun_out/sanitize_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x --timeout 1400 \
  -k "(band and bar-942) or (tiled and bar-72) or (tiled and bar-120)" > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -4 gpurun_out/sanitize_racecheck.log
grep -c "ERROR SUMMARY: 0 errors" gpurun_out/sanitize_memcheck.log gpurun_out/sanitize_racecheck.log
