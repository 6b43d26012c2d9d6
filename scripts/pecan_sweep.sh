#!/bin/bash
# tuning aid: the pair-HMM benchmark under different block-shape classes (max_w:threads:blocks_per_sm,...)
N=${1:-2368}; shift
for cfg in "$@"; do
  echo "== classes: ${cfg}"
  if [ "$cfg" != "default" ]; then export BARB200_PECAN_CLASSES="$cfg"; else unset BARB200_PECAN_CLASSES; fi
  timeout 100 python scripts/prof_pecan.py $N 2 2>&1 | tail -1
done
