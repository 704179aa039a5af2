#!/bin/bash
set -u
TAG=${1:-nccltune}
OUT=gpurun_out/$TAG; mkdir -p "$OUT"; : > "$OUT/summary.txt"
run() {
  name=$1; shift
  echo "== $name" | tee -a "$OUT/summary.txt"
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > "$OUT/bench_$name.json" 2> "$OUT/bench_$name.err"
  python -c "import json;d=json.load(open('$OUT/bench_$name.json'));print('ms/step',d['ms_per_step'], d['phases_ms_rank0'])" | tee -a "$OUT/summary.txt"
  grep -v "OMP_NUM_THREADS\|^\*\*\*\|NCCL version" "$OUT/bench_$name.err" | tail -2 | tee -a "$OUT/summary.txt"
}
run base X=1
run p2p16 NCCL_MIN_P2P_NCHANNELS=16 NCCL_MAX_P2P_NCHANNELS=32
run p2p32 NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32
run halo8 PFS_SLAB_HALO=8
run halo8_p2p16 PFS_SLAB_HALO=8 NCCL_MIN_P2P_NCHANNELS=16 NCCL_MAX_P2P_NCHANNELS=32
