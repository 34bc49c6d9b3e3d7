#!/bin/bash
# Round-2: C5 strong scaling, spatial shards (Morton order of the roots) vs contiguous id ranges.  usage: exp_strong.sh <tag> <N>
set -u
TAG=${1:-strong}; N=${2:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for mode in spatial ids; do
  RVH_BENCH_STRONG_SHARDS=$mode NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --workload c5 --scaling strong --steps 40 --no-e2e --no-configs --no-checksum > $OUT/bench_c5_strong_${mode}_n$N.json 2> $OUT/bench_c5_strong_${mode}_n$N.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_c5_strong_${mode}_n$N.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("N=$N c5 strong $mode value %.4e ms/step %.4f"%(d["value"],d["ms_per_step"]), {k:round(x,4) for k,x in r["per_kernel_ms"].items() if x}, d["implementation"]["parallelism"][:60])
except Exception as e: print("ERR",e); print(open("$OUT/bench_c5_strong_${mode}_n$N.err").read()[-2000:])
PY
done
