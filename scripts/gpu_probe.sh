#!/bin/bash
# Developer helper (GPU box): in-graph cost of the gated fallback launches + ncu capture of selected kernels.
set -u
mkdir -p gpurun_out
bash scripts/gpu_ab2.sh
echo "== assume_in_range"
VAST_OMC_ASSUME_IN_RANGE=1 timeout 300 python bench.py --steps 300 --warmup 10 --no-retrieval --no-cpu > gpurun_out/ab_inrange.json 2> gpurun_out/ab_inrange.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/ab_inrange.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"]*1e3,2), "us/step", d["roofline"]["kernels_us"], d["config"]["loss_last_step"])
PY
K=${NCU_KERNELS:-omc_pack_prep}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -c ${NCU_COUNT:-2} -f -o gpurun_out/probe \
  python bench.py --steps 4 --warmup 3 --no-retrieval --no-cpu --no-graph > gpurun_out/ncu_probe.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_probe.log
