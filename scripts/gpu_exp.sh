#!/bin/bash
set -u
TAG=${1:-exp}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"; : > "$OUT/summary.txt"
LIB=probabilistic_fluid_simulation_b200/lib/libpfs_b200.so
run() {
  name=$1; extra=$2; shift; shift
  echo "== $name" | tee -a "$OUT/summary.txt"
  env "$@" timeout 600 python bench.py --no-e2e --no-cpu $extra > "$OUT/bench_$name.json" 2> "$OUT/bench_$name.err"
  python -c "import json;d=json.load(open('$OUT/bench_$name.json'));print('ms/step %.4f'%d['ms_per_step'], 'eager %.4f'%d['phase_region']['ms_per_step_eager_with_phase_events'], {k: round(v,4) for k,v in d['phases_ms'].items()}, d['gpu_launches'])" | tee -a "$OUT/summary.txt"
  tail -2 "$OUT/bench_$name.err" | tee -a "$OUT/summary.txt"
}
cp scratch_libs/libpfs_r8.so $LIB
timeout 600 python -m pytest tests/test_gpu_operators.py tests/test_gpu_fullsize.py -x -q -m gpu > "$OUT/pytest_r8.log" 2>&1; echo "pytest r8 exit $?" | tee -a "$OUT/summary.txt"; tail -2 "$OUT/pytest_r8.log" | tee -a "$OUT/summary.txt"
run r8 "--steps 50 --warmup 5" X=1
run r8_cfg2 "--width 1024 --height 1024 --iters 50 --steps 400 --warmup 20" X=1
cp scratch_libs/libpfs_r4.so $LIB
run r4 "--steps 50 --warmup 5" X=1
