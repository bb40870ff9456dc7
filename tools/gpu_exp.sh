#!/bin/bash
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
run ttm X=0
run ttm TACO_B200_LIB=$PWD/tools/ab/libtaco_b200_r1.so
run spmm X=0
run spmm TACO_B200_LIB=$PWD/tools/ab/libtaco_b200_r1.so
run mttkrp X=0
run mttkrp TACO_B200_LIB=$PWD/tools/ab/libtaco_b200_r1.so
run sddmm X=0
run sddmm TACO_B200_LIB=$PWD/tools/ab/libtaco_b200_r1.so
run spmv X=0
run spmv TACO_B200_LIB=$PWD/tools/ab/libtaco_b200_r1.so
} > gpurun_out/exp_r2_11.txt 2>&1
cat gpurun_out/exp_r2_11.txt
