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
cp scratch_libs/libpfs_c5.so $LIB
run c5_d5 "--steps 50 --warmup 5" PFS_DIFFUSE_DEPTH=5
run c5_d6 "--steps 50 --warmup 5" PFS_DIFFUSE_DEPTH=6
cp scratch_libs/libpfs_c6.so $LIB
timeout 600 python -m pytest tests/test_gpu_operators.py tests/test_gpu_fullsize.py -x -q -m gpu > "$OUT/pytest_c6.log" 2>&1; echo "pytest c6 exit $?" | tee -a "$OUT/summary.txt"; tail -2 "$OUT/pytest_c6.log" | tee -a "$OUT/summary.txt"
run c6_d6 "--steps 50 --warmup 5" PFS_DIFFUSE_DEPTH=6
run c6_d5 "--steps 50 --warmup 5" PFS_DIFFUSE_DEPTH=5
cp scratch_libs/libpfs_c7.so $LIB
run c7_d7 "--steps 50 --warmup 5" PFS_DIFFUSE_DEPTH=7
cp scratch_libs/libpfs_c5.so $LIB
