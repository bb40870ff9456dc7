#!/bin/bash
# Experiment pass on one GPU box: parity tests, then A/B runs of launch variants selected by environment variables.
# usage: tools/gpu_exp.sh   (edit the EXPS list)   outputs: gpurun_out/exp_*.txt
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/smi.txt 2>&1; nproc >> gpurun_out/smi.txt
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
ncuq() {  # workload, kernel regex, env...
  wl=$1; k=$2; shift; shift
  echo "== ncu $wl $k $*"
  env "$@" timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct \
     --clock-control none -k regex:$k --launch-skip ${SKIP:-3} -c ${CNT:-1} python bench.py --workload $wl --steps 1 --warmup 3 --no-e2e --no-cpu 2>&1 \
     | grep -E "dram__|gpu__time|hit_rate|void " 
}
{
timeout 1500 python -m pytest tests/test_parity_gpu.py tests/test_multi_gpu_gpu.py tests/test_bench_contract_gpu.py -m gpu -q -x 2>&1 | tail -8
run spmm X=0
SKIP=24 CNT=8 ncuq spmm "spmm_" X=0
timeout 900 python bench.py --steps 10 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 600 gpurun_out/bench_default.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_ref.err
} > gpurun_out/exp_r2_3.txt 2>&1
cat gpurun_out/exp_r2_3.txt
