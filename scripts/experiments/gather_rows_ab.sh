mkdir -p gpurun_out/r2h
python -m pytest tests -m gpu -x -q > gpurun_out/r2h/pytest.log 2>&1; tail -2 gpurun_out/r2h/pytest.log
for r in 2 1 2 1; do
PFS_GATHER_ROWS=$r python bench.py --steps 20 --warmup 3 --no-unit --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('rows=$r', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['phases_ms'].items()}, 'stateless', round(d['stateless_entry_points']['ms_per_step'],4))"
done
