#!/bin/bash
# Generic GPU-box pass.  usage: gpu_pass.sh <tag> [tests] [bench workloads...] ; env NCU_K=<kernel regex> for a full capture
set -u
TAG=${1:-pass}; shift || true
OUT=gpurun_out/$TAG; mkdir -p $OUT
if [ "${RUN_TESTS:-1}" = "1" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
  tail -5 $OUT/pytest_gpu.log
fi
for w in "$@"; do
  timeout 600 python bench.py --workload $w ${BENCH_ARGS:-} > $OUT/bench_$w.json 2> $OUT/bench_$w.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_$w.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("$w value %.3e ms/step %.4f frac %.3f step_frac %.3f"%(d["value"],d["ms_per_step"],r["frac"],r["step_frac"]), {k:round(v,4) for k,v in r["per_kernel_ms"].items()}, d.get("clocks"))
except Exception as e:
    print("$w ERR", e); print(open("$OUT/bench_$w.err").read()[-1500:])
PY
done
if [ -n "${NCU_K:-}" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$NCU_K" -s ${NCU_SKIP:-8} -c ${NCU_COUNT:-4} -o $OUT/prof_${NCU_W:-ns_full} -f python bench.py --workload ${NCU_W:-ns_full} --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full.log 2>&1
  tail -2 $OUT/ncu_full.log | cut -c1-300
fi
if [ -n "${NCU_LIST:-}" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_${NCU_W:-ns_full}.csv python bench.py --workload ${NCU_W:-ns_full} --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_launches.log 2>&1
fi
