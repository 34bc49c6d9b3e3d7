#!/bin/bash
# A/B the splat kernels (RVH_SPLAT_VARIANT) on the grid workloads; parity tests for both.
set -u
OUT=gpurun_out/${1:-ab}; mkdir -p $OUT
for v in 0 1; do
  RVH_SPLAT_VARIANT=$v timeout 900 python -m pytest tests -m gpu -x -q -k "grid or synthetic or c1 or ragged or border or full_size" > $OUT/pytest_v$v.log 2>&1; tail -1 $OUT/pytest_v$v.log
  for w in ns_full c3 c4; do
    RVH_SPLAT_VARIANT=$v timeout 300 python bench.py --workload $w --no-cpu-baseline --no-e2e > $OUT/bench_${w}_v$v.json 2>$OUT/bench_${w}_v$v.err
    python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_${w}_v$v.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("variant $v $w ms/step %.4f"%d["ms_per_step"], {k:round(x,4) for k,x in r["per_kernel_ms"].items() if x})
except Exception as e: print("ERR",e)
PY
  done
done
