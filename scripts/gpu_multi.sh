#!/bin/bash
# Developer helper (GPU box with N GPUs): NCCL correctness check, then the bench line at N GPUs.
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > gpurun_out/multi_check_n$N.log 2>&1
echo "check exit $?"; grep -v "^W\|^\[" gpurun_out/multi_check_n$N.log | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 300 --warmup 10 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench exit $?"
python - $N <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/bench_n{n}.json").read().splitlines() if l.startswith("{")][-1])
    print(round(d["ms_per_step"]*1e3,2),"us/step",d["value"],d["roofline"]["kernels_us"],"e2e",d["e2e"]["value"],"ret",d["retrieval"]["value"], d["retrieval"].get("col_sharded",{}).get("value"))
except Exception as e:
    print("failed",e); print(open(f"gpurun_out/bench_n{n}.err").read()[-2000:])
PY
