#!/bin/bash
# Last pass of a round: GPU tests, smoke, both bench arms, launch list and ncu --set full of the default workload.  usage: gpu_final.sh <tag>
set -u
OUT=gpurun_out/${1:-final}; mkdir -p $OUT
B="--no-cpu-baseline --no-e2e"
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; tail -2 $OUT/pytest_gpu.log
python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err
for w in c1 c3 c4; do timeout 300 python bench.py --workload $w $B > $OUT/bench_$w.json 2> $OUT/bench_$w.err; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_ns_full.csv python bench.py --steps 5 --warmup 3 $B > $OUT/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_grid_splat|k_ftl_step|k_grid_finalize' -s 3 -c 3 -o $OUT/prof_ns_full -f python bench.py --steps 3 --warmup 3 $B > $OUT/ncu_full.log 2>&1
python - <<PY
import json, glob, os
for f in sorted(glob.glob("$OUT/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d.get("roofline") or {}
        print(os.path.basename(f), "%.3e" % d["value"], "ms/step %.4f" % d["ms_per_step"], "frac %.3f" % r.get("frac", 0), {k: round(v, 4) for k, v in (r.get("per_kernel_ms") or {}).items() if v},
              "e2e %.3e" % d["e2e"]["value"] if d.get("e2e") else "")
    except Exception as e:
        print(os.path.basename(f), "ERR", e)
PY
