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
timeout 2400 python -m pytest tests/test_parity_gpu.py tests/test_multi_gpu_gpu.py -m gpu -q -x -k "spmm or mttkrp or ttm or dcsr or csf or shards" 2>&1 | tail -4
run spmm X=0
run spmm TACO_B200_SPMM_LONGCTAS=3
run spmm TACO_B200_SPMM_LONGCTAS=4
run spmm TACO_B200_SPMM_LONGVAR=1
run spmm TACO_B200_SPMM_LONGVAR=3
run spmm TACO_B200_SPMM_LONG=64
run spmm TACO_B200_SPMM_LONG=256
run spmm TACO_B200_SPMM_OVERLAP=0
run mttkrp X=0
run mttkrp TACO_B200_MTTKRP_NOHOIST=1
run mttkrp TACO_B200_MTTKRP_VARIANT=3
run mttkrp_fibers X=0
run mttkrp_fibers TACO_B200_MTTKRP_NOHOIST=1
run mttkrp_fibers TACO_B200_MTTKRP_VARIANT=3
} > gpurun_out/exp_r2_6.txt 2>&1
cat gpurun_out/exp_r2_6.txt
