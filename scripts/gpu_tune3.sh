#!/bin/bash
set -u
TAG=${1:-tune3}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
echo "== pytest" | tee "$OUT/summary.txt"
timeout 1200 python -m pytest tests -x -q -m gpu > "$OUT/pytest_gpu.log" 2>&1
echo "pytest exit $?" | tee -a "$OUT/summary.txt"
tail -6 "$OUT/pytest_gpu.log" | tee -a "$OUT/summary.txt"
run() {
  name=$1; shift
  echo "== bench $name" | tee -a "$OUT/summary.txt"
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu $EXTRA > "$OUT/bench_$name.json" 2> "$OUT/bench_$name.err"
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
EXTRA=""
run p6 X=1
run p4 PFS_PRESSURE_DEPTH=4
run p8 PFS_PRESSURE_DEPTH=8
run p6w12 PFS_PRESSURE_WARPS_PER_SM=12
run p6r128 PFS_PRESSURE_ROWS=128
run p6r64 PFS_PRESSURE_ROWS=64
run scalar PFS_PRESSURE_KERNEL=scalar
EXTRA="--width 1024 --height 1024 --iters 50"
run cfg2_1024 X=1
EXTRA="--width 2048 --height 2048 --iters 30"
run 2048_n30 X=1
