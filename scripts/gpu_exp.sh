#!/bin/bash
set -u
TAG=${1:-exp}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"; : > "$OUT/summary.txt"
timeout 900 python -m pytest tests/test_gpu_operators.py tests/test_gpu_fullsize.py tests/test_gpu_golden.py tests/test_gpu_adaptive_pressure.py -x -q -m gpu > "$OUT/pytest.log" 2>&1
echo "pytest exit $?" | tee -a "$OUT/summary.txt"; tail -4 "$OUT/pytest.log" | tee -a "$OUT/summary.txt"
run() {
  name=$1; extra=$2; shift; shift
  echo "== $name" | tee -a "$OUT/summary.txt"
  env "$@" timeout 600 python bench.py --no-e2e --no-cpu $extra > "$OUT/bench_$name.json" 2> "$OUT/bench_$name.err"
  python -c "import json;d=json.load(open('$OUT/bench_$name.json'));print('ms/step %.4f'%d['ms_per_step'], 'eager %.4f'%d['phase_region']['ms_per_step_eager_with_phase_events'], {k: round(v,4) for k,v in d['phases_ms'].items()}, d['gpu_launches'])" | tee -a "$OUT/summary.txt"
  tail -2 "$OUT/bench_$name.err" | tee -a "$OUT/summary.txt"
}
run default "--steps 50 --warmup 5" X=1
run p7 "--steps 50 --warmup 5" PFS_FUSE_DEPTH=7 PFS_DIFFUSE_DEPTH=6
run p6 "--steps 50 --warmup 5" PFS_FUSE_DEPTH=6 PFS_DIFFUSE_DEPTH=6
run cfg2 "--width 1024 --height 1024 --iters 50 --steps 400 --warmup 20" X=1
