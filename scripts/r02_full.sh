#!/bin/bash
# full single-GPU evidence run: bench line, reference arm, launch list of the bench command, full-set ncu of the POA production grid
# and of the guide-tree kernel, CPU-arm scaling, pair-HMM block-shape sweep
mkdir -p gpurun_out
( timeout 1500 python bench.py --steps 5 --warmup 3 2> gpurun_out/r02_bench.err | tail -1 ) > gpurun_out/r02_bench.json
( timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/r02_bench_ref.err | tail -1 ) > gpurun_out/r02_bench_ref.json
( timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-extras --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1 )
( timeout 1200 ncu --set full --clock-control none --import-source on -k regex:poa_msa_kernel_t128 -s 1 -c 1 -f -o gpurun_out/r02_poa_full592_v4 python scripts/prof_run.py 592 2 8 2000 2>&1 | tail -5 ) > gpurun_out/r02_ncu_v3.log
( timeout 600 ncu --set full --clock-control none -k regex:guide_tree_kernel -s 1 -c 1 -f -o gpurun_out/r02_gt_full python scripts/prof_run.py 2368 2 8 2000 2>&1 | tail -5 ) > gpurun_out/r02_ncu_gt.log
( timeout 900 python scripts/cpu_arm_probe.py 2>&1 | tail -8 ) > gpurun_out/r02_cpu_arm.txt
( timeout 900 bash scripts/pecan_sweep.sh 4736 default "96:32:24:96,1:128:4:320" "96:32:24:96,1:64:12:192" "96:32:24:96,1:128:6:384" 2>&1 | tail -12 ) > gpurun_out/r02_pecan_sweep.txt
cut -c1-300 gpurun_out/r02_bench.json; tail -3 gpurun_out/r02_bench.err; cut -c1-200 gpurun_out/r02_bench_ref.json; cat gpurun_out/r02_cpu_arm.txt; cat gpurun_out/r02_pecan_sweep.txt
