#!/bin/bash
# BCSR bring-up on one GPU box: blocked-kernel parity tests (bounded by timeouts), then the remaining gpu tests.
mkdir -p gpurun_out
{
echo "== CUDA-core blocked kernels only (TACO_B200_BSPMM_TC=0)"
TACO_B200_BSPMM_TC=0 timeout 300 python -m pytest tests -m gpu -x -q -k "bspm" 2>&1 | tail -15
echo "== tensor-core path"
timeout 300 python -m pytest tests -m gpu -x -q -k "bspm" 2>&1 | tail -40
} > gpurun_out/bcsr_1.txt 2>&1
cat gpurun_out/bcsr_1.txt
