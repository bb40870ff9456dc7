#!/bin/bash
# usage: tools/gpu_variants.sh <workload> <ENVVAR> "<values>" [pytest -k expr]
mkdir -p gpurun_out
wl=$1; var=$2; vals=$3; kexpr=$4
if [ -n "$kexpr" ]; then timeout 900 python -m pytest tests -m gpu -x -q -k "$kexpr" 2>&1 | tail -5; fi
for v in $vals; do
  echo "== $var=$v"
  env $var=$v timeout 600 python bench.py --workload $wl --no-e2e --no-cpu --steps 10 2>&1 | tail -1 | python -c "
import sys,json
j=json.loads(sys.stdin.read()); r=j['roofline']
print('value',round(j['value'],1),'ms/step',round(j['ms_per_step'],3),'kernel_ms',round(r['kernel_ms'],3),'frac',round(r['frac'],4))"
done
