#!/bin/bash
set -u
TAG=${1:-exp}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"; : > "$OUT/summary.txt"
timeout 1500 python -m pytest tests -x -q -m gpu > "$OUT/pytest.log" 2>&1
echo "pytest exit $?" | tee -a "$OUT/summary.txt"; tail -6 "$OUT/pytest.log" | tee -a "$OUT/summary.txt"
timeout 600 python scripts/adaptive_report.py > "$OUT/adaptive.json" 2> "$OUT/adaptive.err"; echo "adaptive exit $?" | tee -a "$OUT/summary.txt"
timeout 600 python bench.py --no-e2e --no-cpu --width 16384 --height 2048 --steps 20 --warmup 3 > "$OUT/bench_16384x2048.json" 2> "$OUT/bench_16384x2048.err"; echo "bench slab-shaped exit $?" | tee -a "$OUT/summary.txt"
timeout 900 python bench.py > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"; echo "bench default exit $?" | tee -a "$OUT/summary.txt"
