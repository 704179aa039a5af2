mkdir -p gpurun_out/r2w
python -m pytest tests -m gpu -x -q > gpurun_out/r2w/pytest.log 2>&1; tail -3 gpurun_out/r2w/pytest.log
for pdl in 1 0 1 0; do
PFS_PDL=$pdl python bench.py --steps 20 --warmup 3 --no-unit 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('PDL=$pdl 4096', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['phases_ms'].items()}, 'stateless', round(d['stateless_entry_points']['ms_per_step'],4), 'eager', round(d['phase_region']['ms_per_step_eager_with_phase_events'],4))"
done
for pdl in 1 0; do
PFS_PDL=$pdl python bench.py --width 1024 --height 1024 --iters 50 --steps 50 --warmup 5 --no-unit 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('PDL=$pdl 1024', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['phases_ms'].items()})"
done
