#!/bin/bash
# One line of the weak-scaling table: bench.py at N ranks (ns_full weak + appended c4 weak / c5 strong + e2e + checksum).  usage: gpu_scale_r02.sh <N>
set -u
N=${1:-2}; OUT=gpurun_out/scale_r02; mkdir -p $OUT
nvidia-smi -L | wc -l
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --steps 100 --no-cpu-baseline > $OUT/bench_n1.json 2> $OUT/bench_n1.err
else
  NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 100 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
fi
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_n$N.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("N=$N value %.4e ms/step %.4f e2e %.3e"%(d["value"],d["ms_per_step"],d["e2e"]["value"]), {k:round(x,4) for k,x in r["per_kernel_ms"].items() if x})
    print(d["checksum"]["grid_sum_u64"], d["checksum"]["state_xor_u64"], d["checksum"]["exchange"])
    for c in d["configs"] or []: print("  ", c["workload"], c["scaling"], "%.4e"%c["value"], "ms %.4f"%c["ms_per_step"], {k:round(x,4) for k,x in c["per_kernel_ms"].items() if x})
except Exception as e: print("ERR",e); print(open("$OUT/bench_n$N.err").read()[-2500:])
PY
