#!/bin/bash
# Experiment helper: bench lines for a list of "name|ENV=.. ENV=..|bench args" specs.  usage: exp_bench.sh <tag> spec...
set -u
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
for spec in "$@"; do
  name=${spec%%|*}; rest=${spec#*|}; envs=${rest%%|*}; args=${rest#*|}
  env $envs timeout 600 python bench.py --no-cpu-baseline --no-e2e $args > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_$name.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("$name value %.3e ms/step %.4f frac %.3f step_frac %.3f"%(d["value"],d["ms_per_step"],r["frac"],r["step_frac"]), {k:round(v,4) for k,v in r["per_kernel_ms"].items() if v}, (d.get("clocks") or {}).get("sm_mhz"))
except Exception as e:
    print("$name ERR", e); print(open("$OUT/bench_$name.err").read()[-1500:])
PY
done
