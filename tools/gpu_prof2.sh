#!/bin/bash
# Round-2 ncu captures (one GPU).  For every workload in $1 ("spmm spmv mttkrp ..."):
#   gpurun_out/step_<wl>.csv      every kernel of ONE warmed-up step of `bench.py --workload <wl>` with its duration, DRAM bytes and
#                                 cache hit rates (limited metric set: a few replays per kernel) -> roofline.traffic = sum over the step
#   gpurun_out/ncu_<wl>.raw.csv   `--set full` raw page of the dominant kernel (regex in $KERN_<wl>, default: the longest kernel family)
#   gpurun_out/launches_<wl>.csv  launch list of two timed steps (gpu__time_duration only): the kernel's SHARE of the step
# tools/make_profiles2.py turns these into profiles/r02_<wl>.md and profiles/traffic.json.
mkdir -p gpurun_out
OURS='regex:^(spmm|spmv|sddmm|csf3|spadd|spgemm|slot_first|scan_|total_i64|partition|mttkrp|csr_|csf_|bspm|dcsr|pack_|ing_)'
declare -A KERN=( [spmm]=spmm_long_kernel [spmv]=spmv_csr_kernel [mttkrp]=mttkrp_csf_kernel [mttkrp_fibers]=mttkrp_csf_kernel [sddmm]=sddmm_csr_chunk_kernel [ttv]=spmv_warp_kernel [ttm]=spmm_csr_kernel [spadd]=spadd_union_kernel [spgemm]=spgemm_fill_warp_kernel )
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct
for wl in $1; do
  # launch list first: it tells how many of our kernels one step launches (3 warm-up + 2 timed steps + untimed continuation)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 600 --csv \
     --log-file gpurun_out/launches_$wl.csv python bench.py --workload $wl --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncul_$wl.log 2>&1
  per=$(python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_$wl.csv")) if len(r)>10 and r[0].isdigit()]
names=[r[4].split("(")[0] for r in rows]
# period of the launch sequence = kernels per step
for p in range(1,len(names)//3+1):
    if all(names[i]==names[i+p] for i in range(len(names)-2*p, len(names)-p)):
        print(p); break
else:
    print(1)
PY
)
  skip=$((per*4))
  timeout 900 ncu --metrics $M --clock-control none -k "$OURS" --launch-skip $skip -c $per --csv \
     --log-file gpurun_out/step_$wl.csv python bench.py --workload $wl --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncus_$wl.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KERN[$wl]} --launch-skip 3 -c 1 -f \
     -o gpurun_out/ncu_$wl python bench.py --workload $wl --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_$wl.log 2>&1
  ncu -i gpurun_out/ncu_$wl.ncu-rep --page raw --csv > gpurun_out/ncu_$wl.raw.csv 2>/dev/null
  rm -f gpurun_out/ncu_$wl.ncu-rep
  echo "$wl: $per kernels per step"
done
git rev-parse --short HEAD 2>/dev/null > gpurun_out/prof_commit.txt || true
ls -la gpurun_out | head -40
