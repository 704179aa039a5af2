#!/bin/bash
# Multi-GPU checks: NCCL ring parity + scaling bench.  Usage: bash scripts/gpu_multi.sh <tag> <ngpus>
set -u
TAG=${1:-multi}
N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=index,name --format=csv > "$OUT/gpus.csv" 2>&1
nvidia-smi topo -m > "$OUT/topo.txt" 2>&1
echo "== nccl ring test" | tee "$OUT/summary.txt"
timeout 900 python -m pytest tests/test_gpu_slabs.py -x -q -m gpu -k "nccl" > "$OUT/pytest_nccl.log" 2>&1
echo "pytest exit $?" | tee -a "$OUT/summary.txt"
tail -5 "$OUT/pytest_nccl.log" | tee -a "$OUT/summary.txt"
echo "== bench 1 GPU, slab-sized workload 16384x2048 (regular path)" | tee -a "$OUT/summary.txt"
timeout 900 python bench.py --gpus 1 --width 16384 --height 2048 --steps 10 --warmup 3 --no-e2e --no-cpu > "$OUT/bench_n1_slabshape.json" 2> "$OUT/bench_n1_slabshape.err"
python -c "import json;d=json.load(open('$OUT/bench_n1_slabshape.json'));print('ms/step',d['ms_per_step'],'value',d['value'], d['phases_ms'])" | tee -a "$OUT/summary.txt"
tail -2 "$OUT/bench_n1_slabshape.err" | tee -a "$OUT/summary.txt"
for n in $(seq 2 $N); do
  case $n in 2|4|8) ;; *) continue;; esac
  echo "== bench $n GPUs" | tee -a "$OUT/summary.txt"
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > "$OUT/bench_n$n.json" 2> "$OUT/bench_n$n.err"
  echo "exit $?" | tee -a "$OUT/summary.txt"
  python -c "import json;d=json.load(open('$OUT/bench_n$n.json'));print('ms/step',d['ms_per_step'],'value',d['value'],'e2e',d['e2e'] and d['e2e']['ms_per_step'], d['phases_ms_rank0'])" | tee -a "$OUT/summary.txt"
  tail -3 "$OUT/bench_n$n.err" | tee -a "$OUT/summary.txt"
done
