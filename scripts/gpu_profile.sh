#!/bin/bash
# ncu evidence for profiles/: launch list of one 4096^2 step and a full-set capture of every kernel of the step.
# Usage (GPU box, repo root): bash scripts/gpu_profile.sh <tag>      -> gpurun_out/<tag>/*.ncu-rep, launches.csv
set -u
TAG=${1:-prof}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
export PFS_STEP_GRAPH=0          # ncu profiles kernel launches; keep them eager
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -s 0 -c 400 --csv --log-file "$OUT/launches.csv" python scripts/profile_step.py 0 2 > "$OUT/launches.log" 2>&1
$NCU --set full --import-source on --kernel-name-base mangled -k regex:diffuse_packed_kernelILi6ELb0 -s 3 -c 1 -o "$OUT/diffuse_packed_t6" -f python scripts/profile_step.py 0 1 > "$OUT/diffuse.log" 2>&1
$NCU --set full --import-source on --kernel-name-base mangled -k regex:fused_sweeps_kernelILi0ELi8 -s 3 -c 1 -o "$OUT/pressure_fused_t8" -f python scripts/profile_step.py 0 1 > "$OUT/pressure.log" 2>&1
$NCU --set full --import-source on --kernel-name-base mangled -k regex:"advect_kernel|divergence_kernel|project_pack_kernel|advect_color_kernel|sweep_kernel" -s 0 -c 6 -o "$OUT/streaming" -f python scripts/profile_step.py 0 1 > "$OUT/streaming.log" 2>&1
ls -la "$OUT"
# launch list of the bench command itself (eager launches; numbers printed by this run are not bench values)
PFS_STEP_GRAPH=0 $NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file "$OUT/bench_py_launches.csv" python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > "$OUT/bench_under_ncu.log" 2>&1
