#!/bin/bash
# N=2: slab tests (both transports) + the N=2 bench line with each transport
set -u
TAG=${1:-exp}; OUT=gpurun_out/$TAG; mkdir -p "$OUT"; : > "$OUT/summary.txt"
timeout 900 python -m pytest tests/test_gpu_slabs.py -x -q -m gpu > "$OUT/pytest_slabs.log" 2>&1
echo "pytest slabs exit $?" | tee -a "$OUT/summary.txt"; tail -30 "$OUT/pytest_slabs.log" | cut -c1-400 | tee -a "$OUT/summary.txt"
n=2
for tr in p2p nccl; do
  if [ $tr = nccl ]; then export PFS_SLAB_TRANSPORT=nccl; else unset PFS_SLAB_TRANSPORT; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 --no-e2e > "$OUT/bench_n${n}_$tr.json" 2> "$OUT/bench_n${n}_$tr.err"
  echo "bench $tr exit $?" | tee -a "$OUT/summary.txt"
  python -c "import json;d=json.load(open('$OUT/bench_n${n}_$tr.json'));print('ms/step',d['ms_per_step'],'value',d['value'], d['phases_ms_rank0'], d['transport'][:40])" | tee -a "$OUT/summary.txt"
  grep -v "OMP_NUM_THREADS\|^\*\*\*\|NCCL version" "$OUT/bench_n${n}_$tr.err" | tail -3 | tee -a "$OUT/summary.txt"
done
