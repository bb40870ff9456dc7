#!/bin/bash
# strong-scaling bench at N GPUs of one box (usage: tools/gpu_scale.sh N [steps]); output gpurun_out/bench_n<N>.json
N=${1:-8}; STEPS=${2:-10}
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -$N
nvidia-smi topo -m 2>/dev/null | head -14
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps $STEPS --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"
grep -v "^\*\|OMP_NUM" gpurun_out/bench_n$N.err | tail -20
} > gpurun_out/exp_n$N.txt 2>&1
cat gpurun_out/exp_n$N.txt
