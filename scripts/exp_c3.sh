#!/bin/bash
# Round-2: latency-bound mid-size scene C3 (100K x 64): one strand per thread (twice the warps) vs two
set -u
TAG=${1:-c3b}
A='--no-configs --no-checksum --steps 100'
bash scripts/exp_bench.sh $TAG "c3_spt1||$A --workload c3 --spt 1" "c3_spt2||$A --workload c3 --spt 2"
