#!/bin/bash
# 2-GPU validation: multi-device context test, torchrun bench at N=2, bar() end to end through the shim library
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_two_smi.txt
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "every_visible_device or mixed_shapes or flower_submit" 2>&1 | tail -6 ) > gpurun_out/r02_two_pytest.txt
( timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 2> gpurun_out/r02_bench_n2.err | tail -1 ) > gpurun_out/r02_bench_n2.json
( timeout 1200 python scripts/bar_e2e.py 240 2>&1 | tail -22 ) > gpurun_out/r02_bar_e2e.txt
tail -3 gpurun_out/r02_two_pytest.txt; cut -c1-300 gpurun_out/r02_bench_n2.json; tail -4 gpurun_out/r02_bench_n2.err; cat gpurun_out/r02_bar_e2e.txt
