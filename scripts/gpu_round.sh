#!/bin/bash
# Developer helper run on the GPU box: diag + the gpu test files, logs into gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
timeout 300 python scripts/gemm_diag.py > gpurun_out/diag.log 2>&1; echo "diag exit $?" >> gpurun_out/diag.log
for f in "$@"; do
  timeout 900 python -m pytest tests/$f -q -m gpu -x --timeout 600 > gpurun_out/${f%.py}.log 2>&1
  echo "exit $?" >> gpurun_out/${f%.py}.log
done
tail -5 gpurun_out/*.log
