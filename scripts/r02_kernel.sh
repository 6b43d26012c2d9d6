#!/bin/bash
# GPU suite + default bench line (+ optional probe with phase clocks)
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r02_pytest.txt
( timeout 900 python bench.py --steps 3 --warmup 3 --pecan-pairs-per-step 0 2> gpurun_out/r02_bench.err | tail -1 ) > gpurun_out/r02_bench.json
( timeout 600 python scripts/gpu_probe.py 2368 2>&1 | tail -12 ) > gpurun_out/r02_probe.txt
tail -3 gpurun_out/r02_pytest.txt; cat gpurun_out/r02_bench.json | cut -c1-600; tail -5 gpurun_out/r02_probe.txt
