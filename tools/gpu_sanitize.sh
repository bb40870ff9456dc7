#!/bin/bash
# compute-sanitizer over the parity tests of the newest kernels (bounded by timeouts)
#   usage: tools/gpu_sanitize.sh "<pytest -k selection>" [memcheck|racecheck] [test files...]
mkdir -p gpurun_out
SEL=${1:-"spmm or ttm or ttv or mttkrp or spmv"}
TOOL=${2:-memcheck}
shift; shift
FILES=${@:-tests/test_parity_gpu.py}
timeout ${SAN_TIMEOUT:-1500} compute-sanitizer --tool $TOOL --error-exitcode 7 --print-limit 20 python -m pytest $FILES -m gpu -q -x -k "$SEL" > gpurun_out/sanitize_$TOOL.log 2>&1
echo "exit $?" >> gpurun_out/sanitize_$TOOL.log
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit|Invalid|out of bounds|hazard" gpurun_out/sanitize_$TOOL.log | head -30
