#!/bin/bash
# Round helper (1 GPU, no ncu): gpu tests, smoke(), the default bench line and the reference arm.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/test_gpu.log 2>&1
echo "exit $?" >> gpurun_out/test_gpu.log
tail -n 4 gpurun_out/test_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?"; tail -n 2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "reference arm exit $?"
python - <<'PY'
import json
for f in ("bench", "bench_reference"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json").read().splitlines() if l.startswith("{")][-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "impl")}, (d.get("roofline") or {}).get("step_frac"),
              (d.get("e2e") or {}).get("value"), (d.get("retrieval") or {}).get("value"), ((d.get("retrieval") or {}).get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "unreadable", e)
PY
