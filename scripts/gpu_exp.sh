#!/bin/bash
set -u
TAG=${1:-exp}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"; : > "$OUT/summary.txt"
PFS_DIFFUSE_CELLS=2 timeout 900 python -m pytest tests/test_gpu_operators.py tests/test_gpu_golden.py -x -q -m gpu > "$OUT/pytest_cells2.log" 2>&1
echo "pytest (cells=2) exit $?" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/pytest_cells2.log" | tee -a "$OUT/summary.txt"
run() {
  name=$1; shift
  echo "== $name" | tee -a "$OUT/summary.txt"
  env "$@" timeout 600 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu > "$OUT/bench_$name.json" 2> "$OUT/bench_$name.err"
  python -c "import json;d=json.load(open('$OUT/bench_$name.json'));print('ms/step %.4f'%d['ms_per_step'], {k: round(v,4) for k,v in d['phases_ms'].items()})" | tee -a "$OUT/summary.txt"
  tail -2 "$OUT/bench_$name.err" | tee -a "$OUT/summary.txt"
}
run c4_d6 X=1
run c2_d6 PFS_DIFFUSE_CELLS=2
run c2_d4 PFS_DIFFUSE_CELLS=2 PFS_DIFFUSE_DEPTH=4
run c2_d5 PFS_DIFFUSE_CELLS=2 PFS_DIFFUSE_DEPTH=5
run c2_d7 PFS_DIFFUSE_CELLS=2 PFS_DIFFUSE_DEPTH=7
run c2_d8 PFS_DIFFUSE_CELLS=2 PFS_DIFFUSE_DEPTH=8
run c2_d6_w12 PFS_DIFFUSE_CELLS=2 PFS_DIFFUSE_WARPS_PER_SM=12
