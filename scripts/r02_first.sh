#!/bin/bash
# round-2 first GPU call: GPU suite sanity + full-set ncu capture of the PRODUCTION grid (592 CTAs = 4 per SM) of the POA kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r02_smi.txt
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/r02_pytest_first.txt
( timeout 1200 ncu --set full --clock-control none --import-source on -k regex:poa_msa_kernel_t128 -s 1 -c 1 -f -o gpurun_out/r02_poa_full592 python scripts/prof_run.py 592 2 8 2000 2>&1 | tail -20 ) > gpurun_out/r02_ncu_full592.log
ls -la gpurun_out/
