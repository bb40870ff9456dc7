SAN_TIMEOUT=900 bash tools/gpu_sanitize.sh "not (spmm or ttm or ttv or mttkrp or spmv)" memcheck; cp gpurun_out/sanitize_memcheck.log gpurun_out/sanitize_memcheck_rest.log
SAN_TIMEOUT=900 bash tools/gpu_sanitize.sh "deterministic or long_rows or hub or golden_spmm or golden_spmv or mode" racecheck; cp gpurun_out/sanitize_racecheck.log gpurun_out/sanitize_racecheck_sel.log
SAN_TIMEOUT=600 bash tools/gpu_sanitize.sh "pytaco or from_ or zero" memcheck tests/test_pytaco_gpu.py; cp gpurun_out/sanitize_memcheck.log gpurun_out/sanitize_memcheck_pytaco.log
