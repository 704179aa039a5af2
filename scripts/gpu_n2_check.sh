#!/bin/bash
# Two GPUs: the slab tests (incl. the two-process ones) and one N=2 line of the scaling bench.
set -u
TAG=${1:-n2c}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 400 python -m pytest tests/test_gpu_slabs.py -x -q -m gpu > "$OUT/pytest_slabs.log" 2>&1; echo "pytest exit $?" | tee "$OUT/summary.txt"
tail -3 "$OUT/pytest_slabs.log" | tee -a "$OUT/summary.txt"
bash scripts/gpu_n2.sh "$TAG/n2" > /dev/null 2>&1
cat "$OUT/n2/summary.txt" | tee -a "$OUT/summary.txt"
