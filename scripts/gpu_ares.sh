#!/bin/bash
# Developer helper (GPU box): the A-stationary top-k mainloop (VAST_TOPK_ARES=1) -- retrieval tests, then timings.
set -u
mkdir -p gpurun_out
VAST_TOPK_ARES=1 timeout 200 python -m pytest tests/test_gpu_retrieval.py -q -m gpu --timeout 120 -x > gpurun_out/test_ares.log 2>&1
echo "exit $?" >> gpurun_out/test_ares.log
tail -n 6 gpurun_out/test_ares.log
for a in 0 1; do
  for n in 100000 12500; do VAST_TOPK_ARES=$a NK=100000 timeout 60 python scripts/prof_retrieval.py $n 512 16 3 | sed "s/^/ares=$a /"; done
done
