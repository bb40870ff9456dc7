#!/bin/bash
# BCSR pass on one GPU box: parity under every kernel variant, block-shape sweep, full bench line, ncu capture + launch list
mkdir -p gpurun_out
{
for v in 0 4 13; do
echo "== parity, TACO_B200_BSPMM_VARIANT=$v"
TACO_B200_BSPMM_VARIANT=$v timeout 300 python -m pytest tests -m gpu -q -k "bspm" 2>&1 | tail -4
done
} > gpurun_out/bcsr_5.txt 2>&1
cat gpurun_out/bcsr_5.txt
( time timeout 900 python bench.py --workload bspmm ) > gpurun_out/bench_bspmm.json 2> gpurun_out/bench_bspmm.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bspmm_t --launch-skip 3 -c 1 -f \
   -o gpurun_out/ncu_bspmm python bench.py --workload bspmm --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_bspmm.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:bspm -c 400 --csv \
   --log-file gpurun_out/launches_bspmm.csv python bench.py --workload bspmm --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncul_bspmm.log 2>&1
cat gpurun_out/bench_bspmm.json
