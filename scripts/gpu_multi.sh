#!/bin/bash
# Multi-GPU pass: 2-GPU parity test + torchrun bench at N ranks.   usage: gpu_multi.sh <tag> <N> [workloads...]
set -u
TAG=${1:-multi}; N=${2:-2}; shift; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > $OUT/pytest_multi.log 2>&1; tail -3 $OUT/pytest_multi.log
for w in "$@"; do
  NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload $w --no-e2e > $OUT/bench_${w}_n$N.json 2> $OUT/bench_${w}_n$N.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_${w}_n$N.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("N=$N $w value %.3e ms/step %.4f"%(d["value"],d["ms_per_step"]), {k:round(x,4) for k,x in r["per_kernel_ms"].items() if x})
except Exception as e: print("ERR",e); print(open("$OUT/bench_${w}_n$N.err").read()[-2000:])
PY
done
