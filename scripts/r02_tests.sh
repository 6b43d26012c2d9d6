#!/bin/bash
# GPU suite (verbose on failure) + default bench line
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 ) > gpurun_out/r02_pytest.txt
( timeout 900 python bench.py --steps 3 --warmup 3 --pecan-pairs-per-step 0 2> gpurun_out/r02_bench.err | tail -1 ) > gpurun_out/r02_bench.json
tail -5 gpurun_out/r02_pytest.txt; cut -c1-400 gpurun_out/r02_bench.json; tail -5 gpurun_out/r02_bench.err
