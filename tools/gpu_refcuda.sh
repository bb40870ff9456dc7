#!/bin/bash
# the reference's own generated CUDA (recompiled for sm_100a) on the bench operands: whole-call times + ncu kernel times
mkdir -p gpurun_out
{
timeout 300 python tools/ref_cuda_inputs.py /dev/shm
for k in spmv spmm; do
  echo "== reference-generated CUDA $k: compute() wall time per call"
  timeout 300 oracle/_ref/ref_cuda_$k /dev/shm/ref_$k.tbin 3
  echo "== reference-generated CUDA $k: kernel times (ncu)"
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6 oracle/_ref/ref_cuda_$k /dev/shm/ref_$k.tbin 1 2>&1 \
     | grep -E "computeDeviceKernel|taco_binarySearch|gpu__time|dram__bytes" | head -40
done
} > gpurun_out/refcuda.txt 2>&1
cat gpurun_out/refcuda.txt
