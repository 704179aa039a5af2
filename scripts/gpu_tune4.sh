#!/bin/bash
set -u
TAG=${1:-tune4}
OUT=gpurun_out/$TAG; mkdir -p "$OUT"; : > "$OUT/summary.txt"
echo "== parity (operators)" | tee -a "$OUT/summary.txt"
timeout 900 python -m pytest tests/test_gpu_operators.py -x -q -m gpu > "$OUT/pytest.log" 2>&1
echo "pytest exit $?" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/pytest.log" | tee -a "$OUT/summary.txt"
run() {
  name=$1; shift
  echo "== $name" | tee -a "$OUT/summary.txt"
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > "$OUT/bench_$name.json" 2> "$OUT/bench_$name.err"
  python -c "import json;d=json.load(open('$OUT/bench_$name.json'));print('ms/step %.3f'%d['ms_per_step'], {k: round(v,3) for k,v in d['phases_ms'].items()})" | tee -a "$OUT/summary.txt"
  tail -2 "$OUT/bench_$name.err" | tee -a "$OUT/summary.txt"
}
run d8 X=1
run d6 PFS_DIFFUSE_DEPTH=6
run d6m3 PFS_DIFFUSE_DEPTH=6 PFS_DIFFUSE_MINB=3
run d5 PFS_DIFFUSE_DEPTH=5
run d5m3 PFS_DIFFUSE_DEPTH=5 PFS_DIFFUSE_MINB=3
run d4 PFS_DIFFUSE_DEPTH=4
run d7 PFS_DIFFUSE_DEPTH=7
