#!/bin/bash
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/r02_pytest.txt
( timeout 600 python scripts/gpu_probe.py 2368 2>&1 | tail -8 ) > gpurun_out/r02_probe.txt
( timeout 900 python scripts/shapes_probe.py 2>&1 | tail -12 ) > gpurun_out/r02_shapes.txt
tail -4 gpurun_out/r02_pytest.txt; tail -3 gpurun_out/r02_probe.txt | cut -c1-500; cat gpurun_out/r02_shapes.txt | cut -c1-400
