#!/bin/bash
# torchrun bench at N GPUs (default 8), as the driver launches it
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_n${N}_smi.txt
( timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 3 --warmup 3 2> gpurun_out/r02_bench_n${N}.err | tail -1 ) > gpurun_out/r02_bench_n${N}.json
cut -c1-300 gpurun_out/r02_bench_n${N}.json; tail -5 gpurun_out/r02_bench_n${N}.err
