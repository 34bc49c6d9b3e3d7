#!/bin/bash
# Round-2: step wavefront k_ftl_wave on grid-less small scenes (C2) -- parity, then bench lines with it on and off
set -u
TAG=${1:-wave1}
mkdir -p gpurun_out/$TAG
timeout 900 python -m pytest tests/test_parity_full_gpu.py -m gpu -x -q -k "wavefront or step_n" > gpurun_out/$TAG/pytest.log 2>&1; tail -15 gpurun_out/$TAG/pytest.log
A='--no-configs --no-checksum --steps 640'
bash scripts/exp_bench.sh $TAG "c2_wave||$A --workload c2" "c2_multi|RVH_WAVE_STEPS=0|$A --workload c2"
