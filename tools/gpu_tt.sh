#!/bin/bash
# TTV / TTM pass: full bench lines (e2e + reference CPU), ncu capture + launch list
mkdir -p gpurun_out
for wl in ttv ttm; do
  ( time timeout 900 python bench.py --workload $wl ) > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
done
bash tools/gpu_prof.sh "ttv:spmv_csr ttm:spmm_csr" > /dev/null 2>&1
cat gpurun_out/bench_ttv.json gpurun_out/bench_ttm.json
