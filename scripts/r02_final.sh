#!/bin/bash
# final N=1 bench lines (both arms)
mkdir -p gpurun_out
( timeout 1500 python bench.py --steps 5 --warmup 3 2> gpurun_out/r02_bench.err | tail -1 ) > gpurun_out/r02_bench.json
( timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/r02_bench_ref.err | tail -1 ) > gpurun_out/r02_bench_ref.json
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) > gpurun_out/r02_smoke.txt
cut -c1-260 gpurun_out/r02_bench.json; tail -3 gpurun_out/r02_bench.err; cut -c1-200 gpurun_out/r02_bench_ref.json; cat gpurun_out/r02_smoke.txt
