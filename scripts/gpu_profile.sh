#!/bin/bash
# Round profile pass: both bench arms, every workload, ncu launch list, ncu --set full of the step kernels and of the
# extension kernels.  usage: gpu_profile.sh <tag>
set -u
OUT=gpurun_out/${1:-prof}; mkdir -p $OUT
B="--no-cpu-baseline --no-e2e"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err
for w in ns_nogrid c2 c3 c4 c5 c3_sdf ns_sdf; do timeout 600 python bench.py --workload $w --no-cpu-baseline > $OUT/bench_$w.json 2> $OUT/bench_$w.err; done
timeout 600 python bench.py --workload ns_sdf --flags grid+windB+sdf+sdftma $B > $OUT/bench_ns_sdf_tma.json 2> $OUT/bench_ns_sdf_tma.err
timeout 600 python bench.py --workload ns_full --flags grid+windB+rep $B > $OUT/bench_ns_full_rep.json 2> $OUT/bench_ns_full_rep.err
timeout 600 python bench.py --workload c3 --expand $B > $OUT/bench_c3_expand.json 2> $OUT/bench_c3_expand.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_ns_full.csv python bench.py --steps 5 --warmup 3 $B > $OUT/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_grid_splat|k_ftl_step|k_grid_finalize' -s 3 -c 3 -o $OUT/prof_ns_full -f python bench.py --steps 3 --warmup 3 $B > $OUT/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_ftl_step' -s 1 -c 1 -o $OUT/prof_ns_nogrid -f python bench.py --workload ns_nogrid --steps 3 --warmup 3 $B > $OUT/ncu_full_nogrid.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_ftl_step' -s 2 -c 1 -o $OUT/prof_ns_sdf -f python bench.py --workload ns_sdf --steps 3 --warmup 3 $B > $OUT/ncu_full_sdf.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_ftl_step' -s 2 -c 1 -o $OUT/prof_ns_sdf_tma -f python bench.py --workload ns_sdf --flags grid+windB+sdf+sdftma --steps 3 --warmup 3 $B > $OUT/ncu_full_sdf_tma.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_expand_strands' -s 1 -c 1 -o $OUT/prof_c3_expand -f python bench.py --workload c3 --expand --steps 3 --warmup 3 $B > $OUT/ncu_full_expand.log 2>&1
python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
python - <<PY
import json, glob, os
for f in sorted(glob.glob("$OUT/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d.get("roofline") or {}
        print(os.path.basename(f), "%.3e" % d["value"], "ms/step %.4f" % d["ms_per_step"], "frac %.3f" % r.get("frac", 0), {k: round(v, 4) for k, v in (r.get("per_kernel_ms") or {}).items() if v},
              "e2e %.3e" % d["e2e"]["value"] if d.get("e2e") else "", d.get("expand", ""))
    except Exception as e:
        print(os.path.basename(f), "ERR", e)
PY
