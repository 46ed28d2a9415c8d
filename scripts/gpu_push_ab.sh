#!/bin/bash
# A/B of the all-gather's completion protocol at N GPUs: arrival flags (default) vs the symmetric-memory barrier.
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > gpurun_out/multi_check_n$N.log 2>&1
echo "check exit $?"; grep -v "^W\|^\[" gpurun_out/multi_check_n$N.log | tail -4
for s in 1 0 1 0; do
  VAST_PEER_SIGNAL=$s timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$s bench.py --gpus $N --steps 300 --warmup 10 --no-retrieval --no-headroom --no-full-path --no-cpu > gpurun_out/bench_sig${s}_n$N.json 2> gpurun_out/bench_sig${s}_n$N.err
  python - $N $s <<'PY'
import json,sys
n,s=sys.argv[1],sys.argv[2]
try:
    d=json.loads([l for l in open(f"gpurun_out/bench_sig{s}_n{n}.json").read().splitlines() if l.startswith("{")][-1])
    print("signal",s,round(d["ms_per_step"]*1e3,2),"us/step parity",d["parity_ok"],d["roofline"]["kernels_us"],"e2e us",round(d["e2e"]["ms_per_step"]*1e3,1))
except Exception as e:
    print("failed",e); print(open(f"gpurun_out/bench_sig{s}_n{n}.err").read()[-1500:])
PY
done
