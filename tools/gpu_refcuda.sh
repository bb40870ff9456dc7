#!/bin/bash
# the reference's own generated CUDA (recompiled for sm_100a) on the bench operands: whole-call times + ncu kernel times
# usage: tools/gpu_refcuda.sh "spmv spmm" | "mttkrp"
mkdir -p gpurun_out
KS=${1:-"spmv spmm"}
{
if [ "$KS" = "spmv spmm" ]; then timeout 300 python tools/ref_cuda_inputs.py /dev/shm; else timeout 600 python tools/ref_cuda_inputs.py /dev/shm $KS; fi
for k in $KS; do
  echo "== reference-generated CUDA $k: compute() wall time per call"
  timeout 600 oracle/_ref/ref_cuda_$k /dev/shm/ref_$k.tbin 2
  echo "== reference-generated CUDA $k: kernel times (ncu)"
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 8 oracle/_ref/ref_cuda_$k /dev/shm/ref_$k.tbin 1 2>&1 \
     | grep -E "computeDeviceKernel|taco_binarySearch|gpu__time|dram__bytes" | head -40
done
} > gpurun_out/refcuda_$(echo $KS | tr ' ' '_').txt 2>&1
cat gpurun_out/refcuda_*.txt | tail -40
