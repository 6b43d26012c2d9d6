#!/bin/bash
# compute-sanitizer over a small run of every kernel (scripts/sanitize_run.py); summaries -> gpurun_out/r02_sanitizer_*.txt
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck initcheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 100000 python scripts/sanitize_run.py > /tmp/san_$tool.log 2>&1
  {
    echo "== compute-sanitizer --tool $tool python scripts/sanitize_run.py"
    grep -E "^poa:|^windowed|^pair-HMM|SUMMARY" /tmp/san_$tool.log
    echo "-- reports by kind / place:"
    grep -E "Uninitialized|Race reported|Invalid|Barrier error|and (Write|Read) access" /tmp/san_$tool.log | sed -E 's/0x[0-9a-f]+/ADDR/g; s/\+ADDR//; s/\[[0-9]+ hazards\]//' | sort | uniq -c | sort -rn | head -30
    echo "-- device-side reports (kernel names):"
    grep -E "^=========     at " /tmp/san_$tool.log | sed -E 's/0x[0-9a-f]+/ADDR/g' | sort | uniq -c | sort -rn | head -20
  } > gpurun_out/r02_sanitizer_$tool.txt
  cat gpurun_out/r02_sanitizer_$tool.txt | cut -c1-260
done
