#!/bin/bash
# Round helper (GPU box, 1 GPU): full gpu test suite, the default bench line, then -- each only after its own
# command exited 0 without ncu -- the ncu launch list of the same bench command and `ncu --set full` captures of
# the contrastive step's kernels and of the streaming top-k GEMM.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/test_gpu.log 2>&1
echo "exit $?" >> gpurun_out/test_gpu.log
tail -n 6 gpurun_out/test_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
rc=$?; echo "bench exit $rc"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "reference arm exit $?"
if [ $rc -eq 0 ] && [ "${SKIP_NCU:-0}" != "1" ]; then
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_short.json 2>&1 && \
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
  echo "launch list exit $?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc_kernel|omc_' -s 30 -c 9 -f -o gpurun_out/prof_omc \
    python bench.py --steps 6 --warmup 3 --no-cpu --no-retrieval --no-graph > gpurun_out/ncu_omc.log 2>&1
  echo "ncu omc exit $?"
  python scripts/prof_retrieval.py 100000 512 16 1 > gpurun_out/prof_ret_plain.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 1 -f -o gpurun_out/prof_ret \
    python scripts/prof_retrieval.py 100000 512 16 1 > gpurun_out/ncu_ret.log 2>&1
  echo "ncu ret exit $?"
fi
python - <<'PY'
import json
for f in ("bench", "bench_reference"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json").read().splitlines() if l.startswith("{")][-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "impl")}, (d.get("roofline") or {}).get("step_frac"),
              (d.get("e2e") or {}).get("value"), (d.get("retrieval") or {}).get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
