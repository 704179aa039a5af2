#!/bin/bash
set -u
TAG=${1:-sanitize}
OUT=gpurun_out/$TAG; mkdir -p "$OUT"; : > "$OUT/summary.txt"
for tool in memcheck racecheck initcheck; do
  echo "== compute-sanitizer --tool $tool (384x320 grid, 30+30 sweeps, 2 steps; then slabs R=3)" | tee -a "$OUT/summary.txt"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python scripts/sanitize_case.py > "$OUT/$tool.log" 2>&1
  echo "exit $?" | tee -a "$OUT/summary.txt"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard" "$OUT/$tool.log" | head -8 | tee -a "$OUT/summary.txt"
done
