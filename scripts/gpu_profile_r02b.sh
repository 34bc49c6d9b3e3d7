#!/bin/bash
# Late round-2 profile pass: GPU tests, smoke, both bench arms, per-config lines, ncu launch list, ncu --set full of the step kernels
# and of the two small-scene kernels (k_scene_step on C1, k_ftl_wave on C2).   usage: gpu_profile_r02b.sh <tag>
set -u
OUT=gpurun_out/${1:-prof_r02c}; mkdir -p $OUT
B="--no-cpu-baseline --no-e2e --no-configs --no-checksum"
timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; tail -2 $OUT/pytest_gpu.log
python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
timeout 900 python bench.py --steps 300 --warmup 5 > $OUT/bench_default.json 2> $OUT/bench_default.err
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_default_driver_args.json 2> $OUT/bench_default_driver_args.err
timeout 300 python bench.py --workload c1 --steps 400 --no-configs --no-checksum > $OUT/bench_c1.json 2> $OUT/bench_c1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_ns_full.csv python bench.py --steps 5 --warmup 3 $B > $OUT/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_grid_splat|k_ftl_step|k_grid_finalize' -s 12 -c 3 -o $OUT/prof_ns_full -f python bench.py --steps 3 --warmup 3 $B > $OUT/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_scene_step' -s 1 -c 1 -o $OUT/prof_c1_scene -f python bench.py --workload c1 --steps 64 --warmup 3 $B > $OUT/ncu_full_c1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_ftl_wave' -s 1 -c 1 -o $OUT/prof_c2_wave -f python bench.py --workload c2 --steps 64 --warmup 3 $B > $OUT/ncu_full_c2.log 2>&1
python - <<PY
import json, glob, os
for f in sorted(glob.glob("$OUT/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d.get("roofline") or {}
        print(os.path.basename(f), "%.3e" % d["value"], "ms/step %.4f" % d["ms_per_step"], "frac %.3f" % r.get("frac", 0), {k: round(v, 4) for k, v in (r.get("per_kernel_ms") or {}).items() if v},
              "e2e %.3e" % d["e2e"]["value"] if d.get("e2e") else "", "resident %.3e" % d["e2e_resident"]["value"] if d.get("e2e_resident") else "")
        for c in d.get("configs") or []:
            print("   ", c["workload"], "%.3e" % c["value"], "ms/step %.4f" % c["ms_per_step"], "step_frac %.3f" % c["step_frac"], c["step_n_fast_path"][:60])
    except Exception as e:
        print(os.path.basename(f), "ERR", e)
PY
