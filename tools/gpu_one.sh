#!/bin/bash
# one ncu --set full capture: tools/gpu_one.sh <workload> <kernel regex> <out name> [ENV=VAL ...]
wl=$1; k=$2; out=$3; shift; shift; shift
mkdir -p gpurun_out
env "$@" timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 3 -c 1 -f \
   -o gpurun_out/$out python bench.py --workload $wl --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/$out.log 2>&1
tail -2 gpurun_out/$out.log
