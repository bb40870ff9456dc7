#!/bin/bash
# Experiment pass on one GPU box.  outputs: gpurun_out/exp_*.txt
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/smi.txt 2>&1; nproc >> gpurun_out/smi.txt
full() {  # name, workload, kernel regex, skip
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$3 --launch-skip $4 -c 1 -o gpurun_out/ncu_$1 -f \
      python bench.py --workload $2 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_$1.log 2>&1
  ncu -i gpurun_out/ncu_$1.ncu-rep --page raw --csv > gpurun_out/ncu_$1.raw.csv 2>/dev/null
  ncu -i gpurun_out/ncu_$1.ncu-rep --page source --csv > gpurun_out/ncu_$1.source.csv 2>/dev/null
  rm -f gpurun_out/ncu_$1.ncu-rep
}
{
timeout 1500 python -m pytest tests/test_dropin_patch.py -m gpu -q -x 2>&1 | tail -30
full spmm_long spmm spmm_long_kernel 3
full spmm_csr spmm spmm_csr_kernel 3
full spmv spmv spmv_csr_kernel 3
full mttkrp mttkrp mttkrp_csf_kernel 3
} > gpurun_out/exp_r2_4.txt 2>&1
cat gpurun_out/exp_r2_4.txt
