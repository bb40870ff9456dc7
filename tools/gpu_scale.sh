#!/bin/bash
# multi-GPU bench exactly as the driver launches it: tools/gpu_scale.sh N
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 \
   > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 2 --warmup 1 \
   > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err
for wl in ${EXTRA_WL:-spmv mttkrp}; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 10 --warmup 3 --workload $wl --no-cpu \
   > gpurun_out/bench_${wl}_n$N.json 2> gpurun_out/bench_${wl}_n$N.err
done
tail -2 gpurun_out/bench_n$N.err; cat gpurun_out/bench_*_n$N.json gpurun_out/bench_n$N.json
