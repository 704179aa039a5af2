#!/bin/bash
set -u
TAG=${1:-quick}
OUT=gpurun_out/$TAG; mkdir -p "$OUT"; : > "$OUT/summary.txt"
echo "== parity (operators + slabs + stochastic)" | tee -a "$OUT/summary.txt"
timeout 900 python -m pytest tests/test_gpu_operators.py tests/test_gpu_slabs.py tests/test_stochastic.py -x -q -m gpu > "$OUT/pytest.log" 2>&1
echo "pytest exit $?" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/pytest.log" | tee -a "$OUT/summary.txt"
run() {
  name=$1; extra=$2
  echo "== $name" | tee -a "$OUT/summary.txt"
  timeout 600 python bench.py --no-e2e --no-cpu $extra > "$OUT/bench_$name.json" 2> "$OUT/bench_$name.err"
  python -c "import json;d=json.load(open('$OUT/bench_$name.json'));print('ms/step %.4f'%d['ms_per_step'], 'value %.4g'%d['value'], {k: round(v,3) for k,v in d['phases_ms'].items()}, 'launches', d['gpu_launches'])" | tee -a "$OUT/summary.txt"
  tail -2 "$OUT/bench_$name.err" | tee -a "$OUT/summary.txt"
}
run headline "--steps 50 --warmup 5"
run headline_graph "--steps 50 --warmup 5 --graph"
run cfg2 "--width 1024 --height 1024 --iters 50 --steps 400 --warmup 20"
run cfg2_graph "--width 1024 --height 1024 --iters 50 --steps 400 --warmup 20 --graph"
run n30_2048 "--width 2048 --height 2048 --iters 30 --steps 200 --warmup 10"
run n30_2048_graph "--width 2048 --height 2048 --iters 30 --steps 200 --warmup 10 --graph"
