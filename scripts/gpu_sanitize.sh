#!/bin/bash
# compute-sanitizer over scripts/sanitizer_cases.py, ONE tool per gpurun call (B200_PROFILING.md):
#   gpurun -- 'bash scripts/gpu_sanitize.sh memcheck'      (then racecheck, synccheck in separate calls)
# The same command must have exited 0 without the tool first; the log lands in gpurun_out/sanitize_<tool>.log.
set -u
tool=${1:-memcheck}
mkdir -p gpurun_out
timeout 300 python scripts/sanitizer_cases.py > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitize_plain.log; exit 1; }
timeout 1500 compute-sanitizer --tool "$tool" --print-limit 20 python scripts/sanitizer_cases.py > gpurun_out/sanitize_$tool.log 2>&1
echo "compute-sanitizer $tool exit $?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error:|Hazard|sanitizer cases done" gpurun_out/sanitize_$tool.log | sort | uniq -c | head -20
