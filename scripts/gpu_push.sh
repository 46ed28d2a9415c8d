#!/bin/bash
# Developer helper (N GPUs): the fused pack + gather with multimem stores vs unicast peer stores.
set -u
N=${1:-2}
mkdir -p gpurun_out
for mc in 1 0; do
  VAST_PEER_MULTICAST=$mc timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$mc bench.py --gpus $N --steps 300 --warmup 10 --no-retrieval --no-cpu > gpurun_out/push_mc$mc.json 2> gpurun_out/push_mc$mc.err
  python - $mc <<'PY'
import json,sys
mc=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/push_mc{mc}.json").read().splitlines() if l.startswith("{")][-1])
    print("multicast" if mc=="1" else "unicast", round(d["ms_per_step"]*1e3,2),"us/step",d["roofline"]["kernels_us"],"e2e",round(d["e2e"]["value"]/1e6,2), d["config"]["workload"][-70:])
except Exception as e:
    print("failed",e); print(open(f"gpurun_out/push_mc{mc}.err").read()[-1500:])
PY
done
