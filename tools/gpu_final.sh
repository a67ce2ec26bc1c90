#!/bin/bash
# final lines of a round: smoke, GPU tests, bench (N = number of visible GPUs via $1), reference arm
mkdir -p gpurun_out
N=${1:-1}
if [ "$N" == "1" ]; then
  timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
  timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
  timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
  timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
  timeout 600 python tools/quick_time.py > gpurun_out/quick_time.log 2>&1; echo "quick rc=$?"
else
  for n in 2 4 8; do
    if [ "$n" -le "$N" ]; then
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; echo "bench n=$n rc=$?"
    fi
  done
  timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/pytest_multi.log 2>&1; echo "multi rc=$?"; tail -2 gpurun_out/pytest_multi.log
fi
