#!/bin/bash
set -u
TAG=${1:-chunk}
OUT=gpurun_out/$TAG; mkdir -p "$OUT"; : > "$OUT/summary.txt"
echo "== quick parity (operators + slabs)" | tee -a "$OUT/summary.txt"
timeout 900 python -m pytest tests/test_gpu_operators.py tests/test_gpu_slabs.py tests/test_gpu_fullsize.py -x -q -m gpu > "$OUT/pytest.log" 2>&1
echo "pytest exit $?" | tee -a "$OUT/summary.txt"; tail -3 "$OUT/pytest.log" | tee -a "$OUT/summary.txt"
run1() {
  name=$1; extra=$2; shift; shift
  echo "== $name" | tee -a "$OUT/summary.txt"
  env "$@" timeout 600 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu $extra > "$OUT/bench_$name.json" 2> "$OUT/bench_$name.err"
  python -c "import json;d=json.load(open('$OUT/bench_$name.json'));print('ms/step %.3f'%d['ms_per_step'], {k: round(v,3) for k,v in d['phases_ms'].items()})" | tee -a "$OUT/summary.txt"
  tail -2 "$OUT/bench_$name.err" | tee -a "$OUT/summary.txt"
}
run1 n1_4096_default "" X=1
run1 n1_4096_rows128 "" PFS_CHUNK_ROWS=128
run1 n1_slabshape "--width 16384 --height 2048" X=1
echo "== N=2" | tee -a "$OUT/summary.txt"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > "$OUT/bench_n2.json" 2> "$OUT/bench_n2.err"
python -c "import json;d=json.load(open('$OUT/bench_n2.json'));print('ms/step',d['ms_per_step'], d['phases_ms_rank0'])" | tee -a "$OUT/summary.txt"
