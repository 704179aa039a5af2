#!/bin/bash
set -u
TAG=${1:-exp}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"; : > "$OUT/summary.txt"
echo "== new tests" | tee -a "$OUT/summary.txt"
timeout 900 python -m pytest tests/test_gpu_operators.py tests/test_gpu_driver.py -x -q -m gpu > "$OUT/pytest.log" 2>&1
echo "pytest exit $?" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/pytest.log" | tee -a "$OUT/summary.txt"
echo "== ncu launch list of bench.py (2 timed steps, eager launches)" | tee -a "$OUT/summary.txt"
PFS_STEP_GRAPH=0 timeout 900 ncu --clock-control none --metrics gpu__time_duration.sum -c 600 --csv --log-file "$OUT/bench_launches.csv" python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > "$OUT/bench_under_ncu.json" 2> "$OUT/bench_under_ncu.err"
echo "ncu exit $?" | tee -a "$OUT/summary.txt"; wc -l "$OUT/bench_launches.csv" | tee -a "$OUT/summary.txt"
echo "== 16384^2 on one GPU" | tee -a "$OUT/summary.txt"
timeout 900 python scripts/strong_16384.py 16384 100 4 > "$OUT/strong_16384.json" 2> "$OUT/strong_16384.err"
echo "exit $?" | tee -a "$OUT/summary.txt"; cat "$OUT/strong_16384.json" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/strong_16384.err" | tee -a "$OUT/summary.txt"
