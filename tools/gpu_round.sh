#!/bin/bash
# One GPU-box pass: tests, smoke, per-workload bench lines, ncu launch lists. Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1; nproc >> gpurun_out/smi.txt; free -g >> gpurun_out/smi.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
( time timeout 300 python __graft_entry__.py smoke ) > gpurun_out/smoke.log 2>&1
for wl in ${WORKLOADS:-spmm spmv sddmm mttkrp spadd spgemm bspmm bspmv ttv ttm pack}; do
  ( time timeout 900 python bench.py --workload $wl ) > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
done
OURS='regex:^(spmm|spmv|sddmm|csf3|spadd|spgemm|slot_first|scan_|partition|mttkrp|csr_|csf_|bspm|dcsr)'
for wl in ${NCU_WORKLOADS:-spmm}; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 400 --csv \
     --log-file gpurun_out/launches_$wl.csv python bench.py --workload $wl --steps 2 --warmup 3 --no-e2e --no-cpu \
     > gpurun_out/ncu_$wl.log 2>&1
done
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_*.json
