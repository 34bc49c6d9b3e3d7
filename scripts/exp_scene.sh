#!/bin/bash
# Round-2: persistent small-scene kernel (k_scene_step) -- parity tests, then C1 bench lines with the kernel on and off.
set -u
TAG=${1:-scene1}
mkdir -p gpurun_out/$TAG
timeout 900 python -m pytest tests/test_parity_full_gpu.py -m gpu -x -q -k "scene or graph or step_n or rvh_step" > gpurun_out/$TAG/pytest.log 2>&1; tail -15 gpurun_out/$TAG/pytest.log
A='--no-configs --no-checksum --steps 400'
bash scripts/exp_bench.sh $TAG "c1_scene|RVH_SCENE_CTAS=2|$A --workload c1" "c1_graph|RVH_SCENE_CTAS=0|$A --workload c1"
