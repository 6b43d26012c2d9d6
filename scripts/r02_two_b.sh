#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python scripts/inproc_probe.py 2368 2>&1 | tail -24 ) > gpurun_out/r02_inproc.txt
bash scripts/r02_ab.sh > /dev/null 2>&1
cat gpurun_out/r02_inproc.txt; cat gpurun_out/r02_ab.txt
