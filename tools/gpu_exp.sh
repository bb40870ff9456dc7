#!/bin/bash
# Experiment pass on one GPU box.  outputs: gpurun_out/exp_*.txt
mkdir -p gpurun_out
summ='import sys,json
for l in sys.stdin:
    l=l.strip()
    if not l.startswith("{"): continue
    j=json.loads(l); r=j["roofline"]
    print("value",round(j["value"],1),"ms/step",round(j["ms_per_step"],4),"kernel_ms",round(r["kernel_ms"],4),"frac",round(r["frac"],4),"launches",j["gpu_launches"])'
run() {  # workload, env assignments...
  wl=$1; shift
  echo "== $wl $*"
  env "$@" timeout 600 python bench.py --workload $wl --no-e2e --no-cpu --steps 10 2>&1 | tail -1 | python -c "$summ"
}
{
timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
run mttkrp X=0
run mttkrp_fibers X=0
run mttkrp TACO_B200_MTTKRP_VARIANT=4
run mttkrp_fibers TACO_B200_MTTKRP_VARIANT=4
run mttkrp_fibers TACO_B200_MTTKRP_VARIANT=1
run ttm X=0
run spmv X=0
} > gpurun_out/exp_r2_5.txt 2>&1
cat gpurun_out/exp_r2_5.txt
