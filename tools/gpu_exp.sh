#!/bin/bash
# Experiment pass on one GPU box.  outputs: gpurun_out/exp_*.txt
mkdir -p gpurun_out
summ='import sys,json
for l in sys.stdin:
    l=l.strip()
    if not l.startswith("{"): continue
    j=json.loads(l); r=j["roofline"]
    print("value",round(j["value"],1),"ms/step",round(j["ms_per_step"],4),"kernel_ms",round(r["kernel_ms"],4),"frac",round(r["frac"],4),"launches",j["gpu_launches"])'
run() {  # workload, env assignments...
  wl=$1; shift
  echo "== $wl $*"
  env "$@" timeout 600 python bench.py --workload $wl --no-e2e --no-cpu --steps 10 2>&1 | tail -1 | python -c "$summ"
}
{
timeout 2400 python -m pytest tests/test_parity_gpu.py tests/test_multi_gpu_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x -k "spmm or ttm or dcsr or shards or c2 or spadd or sparse" 2>&1 | tail -4
run spmm X=0
run spmm TACO_B200_SPMM_PANELS=16
run spmm TACO_B200_SPMM_PANELS=24
run spmm TACO_B200_SPMM_LONG=96
run spmm TACO_B200_SPMM_LONG=192
run spmm TACO_B200_SPMM_CAP=512
run spmm TACO_B200_SPMM_OVERLAP=1
run ttm X=0
run spadd X=0
run spadd X=0
run spgemm X=0
python - <<'PY'
import sys, time
sys.path[:0]=['.','tests','oracle']
import torch, gpu_util as G, synth, taco_b200 as tb
tb.use_torch_stream(); tb.set_result_space("device")
w = synth.make("spadd", "cuda")
k, ts = G.build("spadd", w)
for _ in range(5): k(*ts)
torch.cuda.synchronize()
for rep in range(3):
    t0=time.perf_counter()
    for _ in range(20): k(*ts)
    torch.cuda.synchronize()
    print("spadd wall per call ms", (time.perf_counter()-t0)/20*1e3)
PY
} > gpurun_out/exp_r2_8.txt 2>&1
cat gpurun_out/exp_r2_8.txt
