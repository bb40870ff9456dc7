#!/bin/bash
# end-of-round pass: all tests, smoke, every bench line (ours + reference arm for the default workload), ncu captures
bash tools/gpu_round.sh > gpurun_out/round.log 2>&1
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref_spmm.json 2> gpurun_out/bench_ref_spmm.err
bash tools/gpu_prof.sh "spmm:spmm_csr_kernel spmv:spmv_csr_kernel sddmm:sddmm_csr mttkrp:mttkrp_csf spadd:spadd_union spgemm:spgemm_fill_warp bspmm:bspmm_tma bspmv:bspmv_warp ttv:spmv_csr ttm:spmm_csr" > gpurun_out/prof.log 2>&1
rm -f gpurun_out/*.ncu-rep
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; du -sh gpurun_out
