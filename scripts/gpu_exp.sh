#!/bin/bash
set -u
TAG=${1:-exp}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"; : > "$OUT/summary.txt"
timeout 900 python -m pytest tests -x -q -m gpu > "$OUT/pytest.log" 2>&1
echo "pytest exit $?" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/pytest.log" | tee -a "$OUT/summary.txt"
run() {
  name=$1; shift
  echo "== $name" | tee -a "$OUT/summary.txt"
  env "$@" timeout 600 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu > "$OUT/bench_$name.json" 2> "$OUT/bench_$name.err"
  python -c "import json;d=json.load(open('$OUT/bench_$name.json'));print('ms/step %.4f'%d['ms_per_step'], {k: round(v,4) for k,v in d['phases_ms'].items()})" | tee -a "$OUT/summary.txt"
  tail -2 "$OUT/bench_$name.err" | tee -a "$OUT/summary.txt"
}
run default X=1
