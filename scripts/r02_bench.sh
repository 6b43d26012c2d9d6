#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python bench.py --steps 3 --warmup 3 2> gpurun_out/r02_bench.err | tail -1 ) > gpurun_out/r02_bench.json
( timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/r02_bench_ref.err | tail -1 ) > gpurun_out/r02_bench_ref.json
cut -c1-300 gpurun_out/r02_bench.json; tail -5 gpurun_out/r02_bench.err; cut -c1-300 gpurun_out/r02_bench_ref.json
