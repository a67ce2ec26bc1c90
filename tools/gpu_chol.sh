#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "tiled or cube or large or path" --timeout 600 2>&1 | tail -4
timeout 600 python tools/configs_time.py 5 2>&1 | tee gpurun_out/config5_time.log
