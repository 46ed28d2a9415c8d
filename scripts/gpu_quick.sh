#!/bin/bash
# Developer helper (GPU box): OMC + API tests, then A/B bench lines (VARIANTS as in gpu_ab2.sh).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_omc.py tests/test_gpu_api.py -q -m gpu --timeout 600 -x > gpurun_out/test_quick.log 2>&1
echo "exit $?" >> gpurun_out/test_quick.log
tail -n ${TAILN:-6} gpurun_out/test_quick.log
bash scripts/gpu_ab2.sh
