#!/bin/bash
set -u
TAG=${1:-exp}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"; : > "$OUT/summary.txt"
run() {
  name=$1; extra=$2; shift; shift
  echo "== $name" | tee -a "$OUT/summary.txt"
  env "$@" timeout 600 python bench.py --no-e2e --no-cpu $extra > "$OUT/bench_$name.json" 2> "$OUT/bench_$name.err"
  python -c "import json;d=json.load(open('$OUT/bench_$name.json'));print('ms/step %.4f'%d['ms_per_step'], 'eager %.4f'%d['phase_region']['ms_per_step_eager_with_phase_events'], {k: round(v,4) for k,v in d['phases_ms'].items()}, d['gpu_launches'])" | tee -a "$OUT/summary.txt"
  tail -2 "$OUT/bench_$name.err" | tee -a "$OUT/summary.txt"
}
timeout 600 python -m pytest tests/test_gpu_operators.py tests/test_gpu_fullsize.py -x -q -m gpu > "$OUT/pytest.log" 2>&1; echo "pytest exit $?" | tee -a "$OUT/summary.txt"; tail -2 "$OUT/pytest.log" | tee -a "$OUT/summary.txt"
run hints "--steps 50 --warmup 5" X=1
run nohints "--steps 50 --warmup 5" PFS_PRESSURE_L2_HINTS=0
run hints_2048 "--width 2048 --height 2048 --iters 100 --steps 100 --warmup 10" X=1
run nohints_2048 "--width 2048 --height 2048 --iters 100 --steps 100 --warmup 10" PFS_PRESSURE_L2_HINTS=0
