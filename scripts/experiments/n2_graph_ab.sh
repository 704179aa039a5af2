mkdir -p gpurun_out/r2x
true
for g in 1 0; do for pdl in 1; do
PFS_SLAB_GRAPH=$g PFS_PDL=$pdl python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$g bench.py --gpus 2 --steps 10 --warmup 3 --no-unit > gpurun_out/r2x/bench2_g${g}_p${pdl}.json 2> gpurun_out/r2x/bench2_g${g}_p${pdl}.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2x/bench2_g${g}_p${pdl}.json') if l.startswith('{')][-1])
print('graph=$g pdl=$pdl', round(d['ms_per_step'],4), 'parity', d['parity']['bit_identical'], {k:round(v,4) for k,v in d['phases_ms_rank0'].items()}, d['gpu_launches'], d['phase_region'], 'e2e', round(d['e2e']['ms_per_step'],2))"
done; done
