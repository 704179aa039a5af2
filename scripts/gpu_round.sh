#!/bin/bash
# One self-contained GPU job: parity tests, smoke, bench, ncu launch list.  Writes to gpurun_out/.
# Usage (from the repo root, on the GPU box):  bash scripts/gpu_round.sh [tag]
set -u
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm,power.limit --format=csv > "$OUT/gpu.csv" 2>&1
echo "== pytest -m gpu" | tee "$OUT/summary.txt"
timeout 1500 python -m pytest tests -x -q -m gpu > "$OUT/pytest_gpu.log" 2>&1
echo "pytest exit $?" | tee -a "$OUT/summary.txt"
tail -5 "$OUT/pytest_gpu.log" | tee -a "$OUT/summary.txt"
echo "== smoke" | tee -a "$OUT/summary.txt"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1
echo "smoke exit $?" | tee -a "$OUT/summary.txt"
tail -3 "$OUT/smoke.log" | tee -a "$OUT/summary.txt"
echo "== bench" | tee -a "$OUT/summary.txt"
timeout 900 python bench.py --steps 10 --warmup 3 > "$OUT/bench.json" 2> "$OUT/bench.err"
echo "bench exit $?" | tee -a "$OUT/summary.txt"
cat "$OUT/bench.json" | tee -a "$OUT/summary.txt"
tail -5 "$OUT/bench.err" | tee -a "$OUT/summary.txt"
