#!/bin/bash
# ncu captures: launch list of one step + full-set capture of the fused kernels.  Usage: gpu_profile.sh <tag>
set -u
TAG=${1:-prof}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
NCU="ncu --clock-control none"
# launch list (every launch of 1 step after 1 warm step), depth 8
$NCU --metrics gpu__time_duration.sum -s 0 -c 400 --csv --log-file "$OUT/launches_d8.csv" python scripts/profile_step.py 8 1 > "$OUT/launches_d8.log" 2>&1
# full capture: pressure T=8 (2 launches), diffuse T=8 (2), diffuse T=4 (2)
$NCU --set full --import-source on --kernel-name-base mangled -k regex:fused_sweeps_kernelILi0ELi8 -s 2 -c 2 -o "$OUT/press_t8" -f python scripts/profile_step.py 8 1 > "$OUT/press_t8.log" 2>&1
$NCU --set full --import-source on --kernel-name-base mangled -k regex:fused_sweeps_kernelILi1ELi8 -s 2 -c 2 -o "$OUT/diff_t8" -f python scripts/profile_step.py 8 1 > "$OUT/diff_t8.log" 2>&1
$NCU --set full --import-source on --kernel-name-base mangled -k regex:fused_sweeps_kernelILi1ELi4 -s 2 -c 2 -o "$OUT/diff_t4" -f python scripts/profile_step.py 4 1 > "$OUT/diff_t4.log" 2>&1
$NCU --set full --import-source on --kernel-name-base mangled -k regex:"advect_kernel|divergence_kernel|project_pack_kernel|advect_color_kernel|sweep_kernel" -s 0 -c 8 -o "$OUT/others" -f python scripts/profile_step.py 8 1 > "$OUT/others.log" 2>&1
ls -la "$OUT"
