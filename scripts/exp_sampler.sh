#!/bin/bash
# What the measurement itself costs a sharded run: kernel events around k_ftl_step (every step / every 4th / none).  usage: exp_sampler.sh <tag> <N> [steps]
set -u
TAG=${1:-sampler}; N=${2:-2}; K=${3:-100}; OUT=gpurun_out/$TAG; mkdir -p $OUT
run() { name=$1; shift
  env "$@" NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps $K --warmup 5 --no-e2e --no-configs --no-checksum ${EXTRA:-} > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_$name.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("$name N=$N steps $K ms/step %.4f  value %.4e  plain-rerun %.4f  ftl %.4f"%(d["ms_per_step"], d["value"], d["ms_per_step"]-r["sampler_overhead_ms_per_step"], r["per_kernel_ms"]["ftl_step"]))
except Exception as e: print("ERR",e); print(open("$OUT/bench_$name.err").read()[-1500:])
PY
}
run events_every_4th X=1
EXTRA=--no-kernel-events run noevents X=1
