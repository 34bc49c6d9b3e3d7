#!/bin/bash
# Scaling pass on one multi-GPU box: bench at N = 8, 4 (and 2) for the given workloads.  usage: gpu_scale.sh <tag> "<Ns>" workloads...
set -u
TAG=${1:-scale}; NS=${2:-"8 4"}; shift; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
for N in $NS; do
 for w in "$@"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --workload $w --no-e2e --steps 50 > $OUT/bench_${w}_n$N.json 2> $OUT/bench_${w}_n$N.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_${w}_n$N.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("N=$N $w value %.3e ms/step %.4f"%(d["value"],d["ms_per_step"]), {k:round(x,4) for k,x in r["per_kernel_ms"].items() if x})
except Exception as e: print("ERR",e); print(open("$OUT/bench_${w}_n$N.err").read()[-1500:])
PY
 done
done
