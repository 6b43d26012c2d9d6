#!/bin/bash
# A/B: CTAs per SM of the 128-thread POA kernel (register budget 128 / 96 / 80 per thread), device Gcell/s of the benchmark stage
mkdir -p gpurun_out
{
echo "== default (4 CTAs/SM, 128 regs, 48 KB scratch)"; python scripts/prof_run.py 2368 3 8 2000 2>&1 | tail -2
echo "== default, 40 KB scratch"; BARB200_SCRATCH_KB=40 python scripts/prof_run.py 2368 3 8 2000 2>&1 | tail -2
echo "== minb5 (5 CTAs/SM, <= 96 regs), 40 KB scratch"; BARB200_LIB=cactus_b200/_variants/libbarb200_minb5.so BARB200_SCRATCH_KB=40 python scripts/prof_run.py 2368 3 8 2000 2>&1 | tail -2
echo "== minb6 (6 CTAs/SM, <= 80 regs), 36 KB scratch"; BARB200_LIB=cactus_b200/_variants/libbarb200_minb6.so BARB200_SCRATCH_KB=36 python scripts/prof_run.py 2368 3 8 2000 2>&1 | tail -2
} > gpurun_out/r02_ab.txt 2>&1
cat gpurun_out/r02_ab.txt
