#!/bin/bash
# compute-sanitizer memcheck over the parity tests of the newest kernels (bounded by timeouts)
mkdir -p gpurun_out
SEL=${1:-"bspm or pack or dcsr or ttv or ttm or mttkrp_host"}
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "$SEL" > gpurun_out/sanitize.log 2>&1
echo "exit $?" >> gpurun_out/sanitize.log
grep -E "ERROR SUMMARY|passed|failed|exit|Invalid|out of bounds" gpurun_out/sanitize.log | head -30
