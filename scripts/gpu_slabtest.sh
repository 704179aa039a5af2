#!/bin/bash
set -u
TAG=${1:-slabtest}; N=${2:-2}
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
echo "== slab tests" | tee "$OUT/summary.txt"
timeout 900 python -m pytest tests/test_gpu_slabs.py -x -q -m gpu > "$OUT/pytest_slabs.log" 2>&1
echo "pytest exit $?" | tee -a "$OUT/summary.txt"; tail -8 "$OUT/pytest_slabs.log" | tee -a "$OUT/summary.txt"
for halo in 32 16 8; do
for n in 2 4 8; do
  [ $n -le $N ] || continue
  echo "== bench $n GPUs halo=$halo" | tee -a "$OUT/summary.txt"
  PFS_SLAB_HALO=$halo timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 --no-e2e > "$OUT/bench_n${n}_h$halo.json" 2> "$OUT/bench_n${n}_h$halo.err"
  echo "exit $?" | tee -a "$OUT/summary.txt"
  python -c "import json;d=json.load(open('$OUT/bench_n${n}_h$halo.json'));print('ms/step',d['ms_per_step'],'value',d['value'], d['phases_ms_rank0'])" | tee -a "$OUT/summary.txt"
  grep -v "OMP_NUM_THREADS\|^\*\*\*" "$OUT/bench_n${n}_h$halo.err" | tail -3 | tee -a "$OUT/summary.txt"
done; done
