#!/bin/bash
set -u
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
echo "== pytest -m gpu" | tee "$OUT/summary.txt"
timeout 1500 python -m pytest tests -x -q -m gpu > "$OUT/pytest_gpu.log" 2>&1
echo "pytest exit $?" | tee -a "$OUT/summary.txt"
tail -6 "$OUT/pytest_gpu.log" | tee -a "$OUT/summary.txt"
echo "== pytest -m gpu with PFS_STEP_GRAPH=0 (eager launches)" | tee -a "$OUT/summary.txt"
PFS_STEP_GRAPH=0 timeout 1500 python -m pytest tests/test_gpu_golden.py tests/test_gpu_operators.py -x -q -m gpu > "$OUT/pytest_gpu_eager.log" 2>&1
echo "pytest exit $?" | tee -a "$OUT/summary.txt"; tail -2 "$OUT/pytest_gpu_eager.log" | tee -a "$OUT/summary.txt"
echo "== smoke" | tee -a "$OUT/summary.txt"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1
echo "smoke exit $?" | tee -a "$OUT/summary.txt"; tail -2 "$OUT/smoke.log" | tee -a "$OUT/summary.txt"
run() {
  name=$1; extra=$2; shift; shift
  echo "== bench $name" | tee -a "$OUT/summary.txt"
  env "$@" timeout 900 python bench.py $extra > "$OUT/bench_$name.json" 2> "$OUT/bench_$name.err"
  python -c "import json;d=json.load(open('$OUT/bench_$name.json'));print('ms/step %.4f'%d['ms_per_step'], 'value %.4g'%d['value'], 'eager %.4f'%d['phase_region']['ms_per_step_eager_with_phase_events'], {k: round(v,3) for k,v in d['phases_ms'].items()}, 'launches', d['gpu_launches'], 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],2))" | tee -a "$OUT/summary.txt"
  tail -2 "$OUT/bench_$name.err" | tee -a "$OUT/summary.txt"
}
run default "" X=1
run default_nograph "--no-e2e --no-cpu" PFS_STEP_GRAPH=0
run cfg2 "--width 1024 --height 1024 --iters 50 --steps 400 --warmup 20 --no-e2e --no-cpu" X=1
run n30_2048 "--width 2048 --height 2048 --iters 30 --steps 200 --warmup 10 --no-e2e --no-cpu" X=1
echo "== bench --impl reference" | tee -a "$OUT/summary.txt"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > "$OUT/bench_reference_arm.json" 2> "$OUT/bench_reference_arm.err"; echo "exit $?" | tee -a "$OUT/summary.txt"; cut -c1-300 "$OUT/bench_reference_arm.json" | tee -a "$OUT/summary.txt"
