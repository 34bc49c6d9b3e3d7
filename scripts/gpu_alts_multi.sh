#!/bin/bash
# Tuning experiment over several workloads: gpu_alts_multi.sh "<libs>" "<workloads>"
for lib in ${1:-librvh.so}; do for w in ${2:-ns_full}; do RVH_LIB=$lib timeout 200 python bench.py --workload $w --no-cpu-baseline --no-e2e --steps 60 | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib $w ms/step %.4f (sampler +%.4f)'%(d['ms_per_step'], d['roofline']['sampler_overhead_ms_per_step']), {k:round(v,4) for k,v in d['roofline']['per_kernel_ms'].items() if v})"; done; done
