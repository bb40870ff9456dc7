#!/bin/bash
mkdir -p gpurun_out
run() { wl=$1; shift; echo "== $wl $*"; env "$@" timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: r=json.loads(l)
    except Exception: continue
    print('value',round(r['value'],1),'ms/step',round(r['ms_per_step'],4),'kernel_ms',round(r['roofline'].get('kernel_ms',0),4),'frac',r['roofline']['frac'],'launches',r['gpu_launches'])
"; }
{
run spmm X=0
run spmm TACO_B200_SPMM_PANELS=8
run spmm TACO_B200_SPMM_PANELS=2
run spmm TACO_B200_SPMM_OVERLAP=0
run spmm X=0
} > gpurun_out/exp_r2_22.txt 2>&1
cat gpurun_out/exp_r2_22.txt
