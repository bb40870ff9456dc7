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
timeout 2400 python -m pytest tests/test_parity_gpu.py tests/test_multi_gpu_gpu.py tests/test_pytaco_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x -k "spmv or ttv or shards or pytaco or c1 or c2 or mttkrp or nccl or csf" 2>&1 | tail -6
run spmv X=0
run spmv TACO_B200_SPMV_KERNEL=0
run spmv TACO_B200_SPMV_KERNEL=2
run spmv TACO_B200_SPMV_KERNEL=3
run spmv TACO_B200_SPMV_KERNEL=4
run ttv X=0
run ttv TACO_B200_SPMV_KERNEL=0
run mttkrp X=0
run mttkrp TACO_B200_MTTKRP_NOTICKET=1
run mttkrp_fibers X=0
run spmm X=0
} > gpurun_out/exp_r2_7.txt 2>&1
cat gpurun_out/exp_r2_7.txt
