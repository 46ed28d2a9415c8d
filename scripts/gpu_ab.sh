#!/bin/bash
# Developer helper (GPU box): OMC tests, then the full gpu suite, then A/B bench lines of the contrastive step:
# default (stats fused into the dQ epilogue, pack fused into prep) vs each fusion switched off.
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_omc.py -q -m gpu --timeout 600 > gpurun_out/test_gpu_omc.log 2>&1
echo "exit $?" >> gpurun_out/test_gpu_omc.log
tail -n 25 gpurun_out/test_gpu_omc.log
timeout 1200 python -m pytest tests -q -m gpu --timeout 600 --deselect tests/test_gpu_omc.py > gpurun_out/test_gpu_rest.log 2>&1
echo "exit $?" >> gpurun_out/test_gpu_rest.log
tail -n 12 gpurun_out/test_gpu_rest.log
S=${BENCH_STEPS:-300}
timeout 300 python bench.py --steps $S --warmup 10 --no-retrieval --no-cpu > gpurun_out/ab_default.json 2> gpurun_out/ab_default.err
VAST_OMC_SEPARATE_ROW_STATS=1 timeout 300 python bench.py --steps $S --warmup 10 --no-retrieval --no-cpu > gpurun_out/ab_sep_stats.json 2> gpurun_out/ab_sep_stats.err
timeout 300 python bench.py --steps $S --warmup 10 --no-retrieval --no-cpu --separate-pack > gpurun_out/ab_sep_pack.json 2> gpurun_out/ab_sep_pack.err
VAST_OMC_SEPARATE_ROW_STATS=1 timeout 300 python bench.py --steps $S --warmup 10 --no-retrieval --no-cpu --separate-pack > gpurun_out/ab_old.json 2> gpurun_out/ab_old.err
for f in ab_default ab_sep_stats ab_sep_pack ab_old; do
  echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(d["ms_per_step"]*1e3, "us/step", d["value"], d["roofline"]["kernels_us"], "e2e", d["e2e"]["ms_per_step"]*1e3)
except Exception as e:
    print("failed", e); print(open(f"gpurun_out/{sys.argv[1]}.err").read()[-1500:])
PY
done
