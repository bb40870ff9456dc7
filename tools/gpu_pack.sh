#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python bench.py --workload pack ) > gpurun_out/bench_pack.json 2> gpurun_out/bench_pack.err
tail -3 gpurun_out/bench_pack.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv \
   --log-file gpurun_out/launches_pack.csv python bench.py --workload pack --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncul_pack.log 2>&1
cat gpurun_out/bench_pack.json
