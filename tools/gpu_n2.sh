#!/bin/bash
# two-GPU pass: NCCL world-2 parity test, then the strong-scaling bench at N=2
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=index,name --format=csv
timeout 900 python -m pytest tests/test_multi_gpu_gpu.py -m gpu -q -x -k nccl 2>&1 | tail -5
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -c 1500 gpurun_out/bench_n2.err
} > gpurun_out/exp_n2.txt 2>&1
cat gpurun_out/exp_n2.txt
