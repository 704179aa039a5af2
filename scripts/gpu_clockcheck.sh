#!/bin/bash
# Short check of bench.py's NVML clock sampler: the default line and a 20-step line (50 ms timed region).
set -u
TAG=${1:-clk}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 600 python bench.py > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"; echo "exit $?" | tee "$OUT/summary.txt"
python -c "import json;d=json.load(open('$OUT/bench_default.json'));print(d['ms_per_step'], d['clocks'], d['e2e']['ms_per_step'])" | tee -a "$OUT/summary.txt"
timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > "$OUT/bench_k20.json" 2> "$OUT/bench_k20.err"; echo "exit $?" | tee -a "$OUT/summary.txt"
python -c "import json;d=json.load(open('$OUT/bench_k20.json'));print(d['ms_per_step'], d['clocks'])" | tee -a "$OUT/summary.txt"
tail -3 "$OUT/bench_default.err" | tee -a "$OUT/summary.txt"
