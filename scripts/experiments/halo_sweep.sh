mkdir -p gpurun_out/r2v
python bench.py --width 16384 --height 2048 --iters 100 --steps 10 --warmup 3 --no-unit 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('unit 16384x2048', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['phases_ms'].items()}, 'eager', d['phase_region'])"
for hh in 32 16 8; do
PFS_SLAB_HALO=$hh python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$((hh/8)) bench.py --gpus 2 --steps 10 --warmup 3 --no-parity --no-unit 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('N=2 halo $hh', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['phases_ms_rank0'].items()}, d['gpu_launches'])"
done
