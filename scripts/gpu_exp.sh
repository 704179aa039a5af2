#!/bin/bash
# two GPUs: the whole GPU suite (incl. the two-process transports), then the N=2 bench line
set -u
TAG=${1:-exp}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"; : > "$OUT/summary.txt"
timeout 1500 python -m pytest tests -x -q -m gpu > "$OUT/pytest.log" 2>&1
echo "pytest exit $?" | tee -a "$OUT/summary.txt"; tail -5 "$OUT/pytest.log" | tee -a "$OUT/summary.txt"
n=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > "$OUT/bench_n$n.json" 2> "$OUT/bench_n$n.err"
echo "bench exit $?" | tee -a "$OUT/summary.txt"
python -c "import json;d=json.load(open('$OUT/bench_n$n.json'));print('ms/step',d['ms_per_step'],'value',d['value'],'e2e',d['e2e'] and d['e2e']['ms_per_step'], d['phases_ms_rank0'], d['transport'][:30])" | tee -a "$OUT/summary.txt"
