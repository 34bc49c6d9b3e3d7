#!/bin/bash
# One GPU-box pass: parity tests, bench lines, ncu launch list and full captures of the hot kernels.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
ls /usr/share/vulkan/icd.d /etc/vulkan/icd.d > gpurun_out/vulkan_probe.txt 2>&1; find / -name 'libvulkan*' -not -path '/proc/*' 2>/dev/null | head >> gpurun_out/vulkan_probe.txt
nproc > gpurun_out/nproc.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_ns_full.json 2> gpurun_out/bench_ns_full.err
for w in ns_nogrid c2 c3 c4; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_ns_full.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_ftl_step|k_grid_splat|k_grid_gather|k_grid_finalize' -s 8 -c 4 -o gpurun_out/prof_ns_full -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_ns_full.json
