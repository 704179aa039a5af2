#!/bin/bash
# One N=2 line of the scaling bench (16384 x 4096 over two slabs), launched the way the driver does.
set -u
TAG=${1:-n2}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > "$OUT/bench_n2.json" 2> "$OUT/bench_n2.err"
echo "exit $?" | tee "$OUT/summary.txt"
python -c "import json;d=json.load(open('$OUT/bench_n2.json'));print('ms/step',d['ms_per_step'],'value',d['value'],'e2e',d['e2e'] and d['e2e']['ms_per_step'], d['phases_ms_rank0'], d['clocks'])" | tee -a "$OUT/summary.txt"
grep -v "OMP_NUM_THREADS\|^\*\*\*\|NCCL version" "$OUT/bench_n2.err" | tail -3 | tee -a "$OUT/summary.txt"
