#!/bin/bash
# Round profile pass: default bench (both arms), ncu launch list, ncu --set full of the step kernels.  usage: gpu_profile.sh <tag>
set -u
OUT=gpurun_out/${1:-prof}; mkdir -p $OUT
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err
for w in ns_nogrid c2 c3 c4; do timeout 600 python bench.py --workload $w --no-cpu-baseline > $OUT/bench_$w.json 2> $OUT/bench_$w.err; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_ns_full.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_grid_splat|k_ftl_step|k_grid_finalize' -s 3 -c 3 -o $OUT/prof_ns_full -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_ftl_step' -s 1 -c 1 -o $OUT/prof_ns_nogrid -f python bench.py --workload ns_nogrid --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_full_nogrid.log 2>&1
python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
cat $OUT/bench_default.json | cut -c1-3000
cat $OUT/bench_reference.json | cut -c1-600
