#!/bin/bash
# Parity of the fused sweeps + a small tuning sweep.  Usage: bash scripts/gpu_tune.sh <tag>
set -u
TAG=${1:-tune}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
echo "== pytest (fused-sensitive tests)" | tee "$OUT/summary.txt"
timeout 1200 python -m pytest tests -x -q -m gpu > "$OUT/pytest_gpu.log" 2>&1
echo "pytest exit $?" | tee -a "$OUT/summary.txt"
tail -15 "$OUT/pytest_gpu.log" | tee -a "$OUT/summary.txt"
for cfg in "0 0" "8 64" "8 128" "8 256" "6 128" "4 128" "1 0"; do
  set -- $cfg
  echo "== bench depth=$1 rows=$2" | tee -a "$OUT/summary.txt"
  PFS_CHUNK_ROWS=$2 timeout 600 python bench.py --steps 10 --warmup 3 --fuse-depth $1 --no-e2e --no-cpu > "$OUT/bench_d$1_r$2.json" 2> "$OUT/bench_d$1_r$2.err"
  python - "$OUT/bench_d$1_r$2.json" <<'PY' | tee -a "$OUT/summary.txt"
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("ms/step %.3f  phases %s  frac_step %.2f" % (d["ms_per_step"], {k: round(v,3) for k,v in d["phases_ms"].items()}, d["whole_step_roofline"]["frac"]))
except Exception as e:
    print("bench failed", e)
PY
  tail -2 "$OUT/bench_d$1_r$2.err" | tee -a "$OUT/summary.txt"
done
