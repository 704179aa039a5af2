#!/bin/bash
# Parity + tuning of the packed diffusion kernel.  Usage: bash scripts/gpu_tune2.sh <tag>
set -u
TAG=${1:-tune2}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
echo "== pytest" | tee "$OUT/summary.txt"
timeout 1200 python -m pytest tests -x -q -m gpu > "$OUT/pytest_gpu.log" 2>&1
echo "pytest exit $?" | tee -a "$OUT/summary.txt"
tail -15 "$OUT/pytest_gpu.log" | tee -a "$OUT/summary.txt"
run() {  # name, env...
  name=$1; shift
  echo "== bench $name" | tee -a "$OUT/summary.txt"
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > "$OUT/bench_$name.json" 2> "$OUT/bench_$name.err"
  python - "$OUT/bench_$name.json" <<'PY' | tee -a "$OUT/summary.txt"
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("ms/step %.3f  phases %s  frac_step %.2f" % (d["ms_per_step"], {k: round(v,3) for k,v in d["phases_ms"].items()}, d["whole_step_roofline"]["frac"]))
except Exception as e:
    print("bench failed", e)
PY
  tail -2 "$OUT/bench_$name.err" | tee -a "$OUT/summary.txt"
}
run default X=1
run d4 PFS_DIFFUSE_DEPTH=4
run d5 PFS_DIFFUSE_DEPTH=5
run d6w12 PFS_DIFFUSE_DEPTH=6 PFS_DIFFUSE_WARPS_PER_SM=12
run d8 PFS_DIFFUSE_DEPTH=8
run d4w8 PFS_DIFFUSE_DEPTH=4 PFS_DIFFUSE_WARPS_PER_SM=8
run d4r64 PFS_DIFFUSE_DEPTH=4 PFS_DIFFUSE_ROWS=64
run d6r128 PFS_DIFFUSE_DEPTH=6 PFS_DIFFUSE_ROWS=128
