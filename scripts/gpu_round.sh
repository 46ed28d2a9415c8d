#!/bin/bash
# Developer helper run on the GPU box: the gpu test files given as arguments (all if none), then a
# short bench; logs into gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
files=("$@")
if [ ${#files[@]} -eq 0 ]; then files=(test_gpu_gemm.py test_gpu_elementwise.py test_gpu_omc.py test_gpu_retrieval.py test_gpu_api.py); fi
for f in "${files[@]}"; do
  timeout 900 python -m pytest tests/$f -q -m gpu -x --timeout 600 > gpurun_out/${f%.py}.log 2>&1
  echo "exit $?" >> gpurun_out/${f%.py}.log
done
timeout 600 python bench.py --steps ${BENCH_STEPS:-100} --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
tail -n 6 gpurun_out/*.log
tail -c 3000 gpurun_out/bench.json
tail -n 5 gpurun_out/bench.err
