#!/bin/bash
# Developer helper (GPU box): retrieval tests + per-rank shapes of cfg5 at 8 / 4 / 2 / 1 GPUs on one GPU.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_retrieval.py tests/test_gpu_gemm.py -q -m gpu --timeout 600 -x > gpurun_out/test_ret.log 2>&1
echo "exit $?" >> gpurun_out/test_ret.log
tail -n 5 gpurun_out/test_ret.log
for n in 12500 25000 50000 100000; do NK=100000 python scripts/prof_retrieval.py $n 512 16 3; done
