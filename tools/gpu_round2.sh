#!/bin/bash
# end-of-round pass on one GPU: all GPU tests, default bench (both arms), ncu captures for profiles/
mkdir -p gpurun_out
{
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 900 python bench.py --steps 20 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 300 gpurun_out/bench_default.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 300 gpurun_out/bench_ref.err
for wl in sddmm spadd spgemm ttv ttm mttkrp_fibers bspmm bspmv pack; do
  timeout 600 python bench.py --workload $wl --steps 10 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
done
bash tools/gpu_prof2.sh "spmm spmv mttkrp mttkrp_fibers sddmm ttv ttm" > gpurun_out/prof2.log 2>&1
tail -5 gpurun_out/prof2.log
} > gpurun_out/exp_round2.txt 2>&1
cat gpurun_out/exp_round2.txt
