#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -X faulthandler -m pytest tests/test_gpu_flowers.py -x -v 2>&1 | tail -80 ) > gpurun_out/r02_dbg.txt
( BARB200_LANES=1 timeout 900 python -X faulthandler -m pytest tests/test_gpu_flowers.py -x -v 2>&1 | tail -30 ) > gpurun_out/r02_dbg_lane1.txt
( timeout 1500 python -X faulthandler -m pytest tests -m gpu -q --deselect tests/test_gpu_flowers.py 2>&1 | tail -40 ) > gpurun_out/r02_dbg_rest.txt
tail -40 gpurun_out/r02_dbg.txt
