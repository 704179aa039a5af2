#!/bin/bash
# N GPUs: the bench line (peer transport), as the driver launches it
set -u
TAG=${1:-exp}; n=${2:-8}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"; : > "$OUT/summary.txt"
nvidia-smi topo -m > "$OUT/topo.txt" 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > "$OUT/bench_n$n.json" 2> "$OUT/bench_n$n.err"
echo "bench exit $?" | tee -a "$OUT/summary.txt"
python -c "import json;d=json.load(open('$OUT/bench_n$n.json'));print('ms/step',d['ms_per_step'],'value',d['value'],'e2e',d['e2e'] and d['e2e']['ms_per_step'], d['phases_ms_rank0'], d['transport'][:40], d['clocks'])" | tee -a "$OUT/summary.txt"
grep -v "OMP_NUM_THREADS\|^\*\*\*\|NCCL version" "$OUT/bench_n$n.err" | tail -3 | tee -a "$OUT/summary.txt"
