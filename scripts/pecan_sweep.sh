#!/bin/bash
# tuning aid: the pair-HMM benchmark under different block-shape classes (max_w:threads:blocks_per_sm,...)
for cfg in "" "96:32:24,320:128:6,608:128:4,1280:128:2" "96:32:24,320:128:6,608:256:3,1280:256:2" "96:32:24,320:64:6,608:64:4,1280:128:2" "96:32:24,320:128:6,608:128:4,800:128:2,1280:256:2" "96:32:24,448:128:5,704:256:3,1280:256:2"; do
  echo "== classes: ${cfg:-default}"
  if [ -n "$cfg" ]; then export BARB200_PECAN_CLASSES="$cfg"; else unset BARB200_PECAN_CLASSES; fi
  timeout 100 python scripts/prof_pecan.py ${1:-2368} 2 2>&1 | tail -1
done
