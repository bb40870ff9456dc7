#!/bin/bash
# ncu captures of the dominant kernel of each workload (one launch, --set full) + launch list of our kernels only.
# usage: tools/gpu_prof.sh "spmm:spmm_csr_kernel spmv:spmv_csr_kernel ..."   outputs: gpurun_out/ncu_<wl>.ncu-rep, launches_<wl>.csv
mkdir -p gpurun_out
OURS='regex:^(spmm|spmv|sddmm|csf3|spadd|spgemm|slot_first|scan_|partition|mttkrp|csr_|csf_|bspm|dcsr)'
for item in $1; do
  wl=${item%%:*}; kern=${item##*:}
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kern --launch-skip 3 -c 1 -f \
     -o gpurun_out/ncu_$wl python bench.py --workload $wl --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_$wl.log 2>&1
  # keep the box's gpurun_out small (64 MiB pull limit): the raw page as CSV is what tools/make_profiles.py reads
  ncu -i gpurun_out/ncu_$wl.ncu-rep --page raw --csv > gpurun_out/ncu_$wl.raw.csv 2>/dev/null && [ "${KEEP_REP:-0}" = 1 ] || rm -f gpurun_out/ncu_$wl.ncu-rep
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 400 --csv \
     --log-file gpurun_out/launches_$wl.csv python bench.py --workload $wl --steps 2 --warmup 3 --no-e2e --no-cpu \
     > gpurun_out/ncul_$wl.log 2>&1
done
ls -la gpurun_out
