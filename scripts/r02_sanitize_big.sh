#!/bin/bash
# memcheck + racecheck with the long classes (t640, t1024) added
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 2400 compute-sanitizer --tool $tool --print-limit 100000 python scripts/sanitize_run.py big > /tmp/san_$tool.log 2>&1
  {
    echo "== compute-sanitizer --tool $tool python scripts/sanitize_run.py big"
    grep -E "^poa:|^windowed|^pair-HMM|SUMMARY" /tmp/san_$tool.log
    grep -E "Uninitialized|Race reported|Invalid|Barrier error|and (Write|Read) access" /tmp/san_$tool.log | sed -E 's/0x[0-9a-f]+/ADDR/g; s/\+ADDR//; s/\[[0-9]+ hazards\]//' | sort | uniq -c | sort -rn | head -30
  } > gpurun_out/r02_sanitizer_big_$tool.txt
  cat gpurun_out/r02_sanitizer_big_$tool.txt | cut -c1-260
done
