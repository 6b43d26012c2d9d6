#!/bin/bash
# device guide tree: GPU suite, full-set ncu capture of the production grid, probe with phase clocks, bench
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/r02_pytest.txt
( timeout 1200 ncu --set full --clock-control none --import-source on -k regex:poa_msa_kernel_t128 -s 1 -c 1 -f -o gpurun_out/r02_poa_full592_v2 python scripts/prof_run.py 592 2 8 2000 2>&1 | tail -8 ) > gpurun_out/r02_ncu_v2.log
( timeout 600 python scripts/gpu_probe.py 2368 2>&1 | tail -8 ) > gpurun_out/r02_probe.txt
( timeout 1500 python bench.py --steps 3 --warmup 3 2> gpurun_out/r02_bench.err | tail -1 ) > gpurun_out/r02_bench.json
tail -4 gpurun_out/r02_pytest.txt; tail -3 gpurun_out/r02_probe.txt; cut -c1-250 gpurun_out/r02_bench.json; tail -3 gpurun_out/r02_bench.err
