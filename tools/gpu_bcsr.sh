#!/bin/bash
# BCSR pass on one GPU box: tcgen05 layout probe, parity, full bench line (e2e + reference CPU), ncu capture + launch list.
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/tc_probe tools/tc_probe.cu && /tmp/tc_probe > gpurun_out/tc_probe.txt 2>&1
grep -E "hyp (0|10|11|12) " gpurun_out/tc_probe.txt
timeout 300 python -m pytest tests -m gpu -q -k "bspm" 2>&1 | tail -3
( time timeout 900 python bench.py --workload bspmm ) > gpurun_out/bench_bspmm.json 2> gpurun_out/bench_bspmm.err
( time timeout 900 python bench.py --workload bspmm --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref_bspmm.json 2> gpurun_out/bench_ref_bspmm.err
OURS='regex:^(bspm|bcsr)' 
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bspmm_tc --launch-skip 3 -c 1 -f \
   -o gpurun_out/ncu_bspmm python bench.py --workload bspmm --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_bspmm.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:bspm -c 400 --csv \
   --log-file gpurun_out/launches_bspmm.csv python bench.py --workload bspmm --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncul_bspmm.log 2>&1
cat gpurun_out/bench_bspmm.json gpurun_out/bench_ref_bspmm.json
