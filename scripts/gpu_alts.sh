#!/bin/bash
# Tuning experiment: same bench with alternative builds of the library (RVH_LIB).
set -u
OUT=gpurun_out/${1:-alts}; mkdir -p $OUT; shift
for lib in librvh.so "$@"; do
  for w in ns_full; do
    RVH_LIB=$lib timeout 300 python bench.py --workload $w --no-cpu-baseline --no-e2e > $OUT/bench_${w}_$lib.json 2>$OUT/bench_${w}_$lib.err
    python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_${w}_$lib.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("$lib $w ms/step %.4f"%d["ms_per_step"], {k:round(x,4) for k,x in r["per_kernel_ms"].items() if x})
except Exception as e: print("ERR",e, open("$OUT/bench_${w}_$lib.err").read()[-500:])
PY
  done
done
