#!/bin/bash
# compute-sanitizer over a small run through every kernel.  usage: gpu_sanitize.sh <tag>
set -u
OUT=gpurun_out/${1:-sanitize}; mkdir -p $OUT
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/probes/sanitize_workload.py > $OUT/$tool.log 2>&1
  echo "== $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|^ok" $OUT/$tool.log | tail -16
done
