#!/bin/bash
# one-workload pass: full bench line (e2e + reference CPU), ncu capture + launch list.  usage: tools/gpu_tt.sh "wl:kernelregex ..."
mkdir -p gpurun_out
for item in $1; do
  wl=${item%%:*}
  ( time timeout 900 python bench.py --workload $wl ) > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
done
bash tools/gpu_prof.sh "$1" > /dev/null 2>&1
rm -f gpurun_out/*.ncu-rep
for item in $1; do cat gpurun_out/bench_${item%%:*}.json; done
