#!/bin/bash
set -u
TAG=${1:-exp}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"; : > "$OUT/summary.txt"
timeout 1500 python -m pytest tests -x -q -m gpu > "$OUT/pytest.log" 2>&1
echo "pytest exit $?" | tee -a "$OUT/summary.txt"; tail -6 "$OUT/pytest.log" | tee -a "$OUT/summary.txt"
timeout 600 python scripts/adaptive_report.py > "$OUT/adaptive.json" 2> "$OUT/adaptive.err"; echo "adaptive exit $?" | tee -a "$OUT/summary.txt"
timeout 600 python scripts/frame_mode_timing.py 1024 24 > "$OUT/frames_1024.json" 2> "$OUT/frames.err"; echo "frames exit $?" | tee -a "$OUT/summary.txt"
timeout 600 python scripts/frame_mode_timing.py 4096 8 > "$OUT/frames_4096.json" 2>> "$OUT/frames.err"; echo "frames4096 exit $?" | tee -a "$OUT/summary.txt"
nproc | tee -a "$OUT/summary.txt"
