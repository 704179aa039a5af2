#!/bin/bash
# Scaling run the way the driver does it: N = 1, 2, 4, 8 back to back.  Usage: gpu_scale.sh <tag> <maxN>
set -u
TAG=${1:-scale}; N=${2:-8}
OUT=gpurun_out/$TAG; mkdir -p "$OUT"
nvidia-smi topo -m > "$OUT/topo.txt" 2>&1
echo "== N=1 headline (4096^2)" | tee "$OUT/summary.txt"
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"
python -c "import json;d=json.load(open('$OUT/bench_n1.json'));print('ms/step',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['ms_per_step'])" | tee -a "$OUT/summary.txt"
echo "== N=1 slab-shaped workload 16384x2048 (regular single-GPU path)" | tee -a "$OUT/summary.txt"
timeout 900 python bench.py --gpus 1 --width 16384 --height 2048 --steps 10 --warmup 3 --no-e2e --no-cpu > "$OUT/bench_n1_slabshape.json" 2> "$OUT/bench_n1_slabshape.err"
python -c "import json;d=json.load(open('$OUT/bench_n1_slabshape.json'));print('ms/step',d['ms_per_step'],'value',d['value'], d['phases_ms'])" | tee -a "$OUT/summary.txt"
for n in 2 4 8; do
  [ $n -le $N ] || continue
  echo "== N=$n" | tee -a "$OUT/summary.txt"
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > "$OUT/bench_n$n.json" 2> "$OUT/bench_n$n.err"
  echo "exit $?" | tee -a "$OUT/summary.txt"
  python -c "import json;d=json.load(open('$OUT/bench_n$n.json'));print('ms/step',d['ms_per_step'],'value',d['value'],'e2e',d['e2e'] and d['e2e']['ms_per_step'], d['phases_ms_rank0'], d['clocks'])" | tee -a "$OUT/summary.txt"
  grep -v "OMP_NUM_THREADS\|^\*\*\*\|NCCL version" "$OUT/bench_n$n.err" | tail -3 | tee -a "$OUT/summary.txt"
done
