#!/bin/bash
# Developer helper (GPU box): A/B bench lines only.  VARIANTS="default sep_pack sep_stats old"
set -u
mkdir -p gpurun_out
S=${BENCH_STEPS:-300}
for v in ${VARIANTS:-default sep_pack}; do
  case $v in
    default) e=0; a="";;
    sep_pack) e=0; a="--separate-pack";;
    sep_stats) e=1; a="";;
    old) e=1; a="--separate-pack";;
  esac
  VAST_OMC_SEPARATE_ROW_STATS=$e timeout 300 python bench.py --steps $S --warmup 10 --no-retrieval --no-cpu $a > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  echo "== $v"; python - "ab_$v" <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(round(d["ms_per_step"]*1e3,2), "us/step", d["roofline"]["kernels_us"], "e2e", round(d["e2e"]["ms_per_step"]*1e3,1))
except Exception as e:
    print("failed", e); print(open(f"gpurun_out/{sys.argv[1]}.err").read()[-1500:])
PY
done
