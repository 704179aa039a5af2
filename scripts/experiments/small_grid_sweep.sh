mkdir -p gpurun_out/r2s
B="python bench.py --width 1024 --height 1024 --iters 50 --steps 50 --warmup 5 --no-unit"
run() { tag=$1; shift; env "$@" $B 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$tag', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['phases_ms'].items()}, d['gpu_launches'])"; }
run default A=1
for dd in 2 3 4 5; do for r in 8 16 32; do run "depth$dd rows$r" PFS_DIFFUSE_DEPTH=$dd PFS_DIFFUSE_ROWS=$r; done; done
for fd in 2 3 4 6 8; do for r in 8 16 32; do run "pdepth$fd prows$r" PFS_FUSE_DEPTH=$fd PFS_CHUNK_ROWS=$r; done; done
