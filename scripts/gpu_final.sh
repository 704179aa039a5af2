#!/bin/bash
# Round-end check on one B200: the whole GPU suite, smoke(), the default bench line, the reference arm,
# and a re-check of the diffusion fuse depth on the final kernels.
set -u
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
echo "== pytest -m gpu" | tee "$OUT/summary.txt"
timeout 900 python -m pytest tests -x -q -m gpu > "$OUT/pytest_gpu.log" 2>&1
echo "pytest exit $?" | tee -a "$OUT/summary.txt"
tail -4 "$OUT/pytest_gpu.log" | tee -a "$OUT/summary.txt"
echo "== smoke" | tee -a "$OUT/summary.txt"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1
echo "smoke exit $?" | tee -a "$OUT/summary.txt"; tail -2 "$OUT/smoke.log" | tee -a "$OUT/summary.txt"
run() {
  name=$1; extra=$2; shift; shift
  echo "== bench $name" | tee -a "$OUT/summary.txt"
  env "$@" timeout 600 python bench.py $extra > "$OUT/bench_$name.json" 2> "$OUT/bench_$name.err"
  python -c "import json;d=json.load(open('$OUT/bench_$name.json'));print('ms/step %.4f'%d['ms_per_step'], 'value %.4g'%d['value'], {k: round(v,3) for k,v in d['phases_ms'].items()}, 'launches', d['gpu_launches'], 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],2))" | tee -a "$OUT/summary.txt"
  tail -2 "$OUT/bench_$name.err" | tee -a "$OUT/summary.txt"
}
run default "" X=1
run depth4 "--steps 40 --warmup 5 --no-e2e --no-cpu" PFS_DIFFUSE_DEPTH=4
run depth5 "--steps 40 --warmup 5 --no-e2e --no-cpu" PFS_DIFFUSE_DEPTH=5
echo "== bench --impl reference" | tee -a "$OUT/summary.txt"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > "$OUT/bench_reference_arm.json" 2> "$OUT/bench_reference_arm.err"; echo "exit $?" | tee -a "$OUT/summary.txt"; cut -c1-300 "$OUT/bench_reference_arm.json" | tee -a "$OUT/summary.txt"
