#!/bin/bash
set -u
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
echo "== pytest -m gpu" | tee "$OUT/summary.txt"
timeout 1500 python -m pytest tests -x -q -m gpu > "$OUT/pytest_gpu.log" 2>&1
echo "pytest exit $?" | tee -a "$OUT/summary.txt"
tail -6 "$OUT/pytest_gpu.log" | tee -a "$OUT/summary.txt"
echo "== smoke" | tee -a "$OUT/summary.txt"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1
echo "smoke exit $?" | tee -a "$OUT/summary.txt"; tail -2 "$OUT/smoke.log" | tee -a "$OUT/summary.txt"
echo "== bench (default flags)" | tee -a "$OUT/summary.txt"
timeout 900 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"
echo "bench exit $?" | tee -a "$OUT/summary.txt"; cat "$OUT/bench.json" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/bench.err" | tee -a "$OUT/summary.txt"
echo "== bench --impl reference" | tee -a "$OUT/summary.txt"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"
cat "$OUT/bench_reference.json" | cut -c1-400 | tee -a "$OUT/summary.txt"
echo "== config 5 (stochastic, 4096^2, 1000 steps, 30+30 sweeps)" | tee -a "$OUT/summary.txt"
timeout 900 python scripts/stochastic_run.py 4096 1000 0.01 30 > "$OUT/stochastic_4096.jsonl" 2> "$OUT/stochastic.err"
tail -1 "$OUT/stochastic_4096.jsonl" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/stochastic.err" | tee -a "$OUT/summary.txt"
echo "== bench config 2 (1024^2, 50 sweeps)" | tee -a "$OUT/summary.txt"
timeout 600 python bench.py --width 1024 --height 1024 --iters 50 --steps 200 --warmup 10 --no-e2e > "$OUT/bench_cfg2.json" 2> "$OUT/bench_cfg2.err"
cut -c1-300 "$OUT/bench_cfg2.json" | tee -a "$OUT/summary.txt"
