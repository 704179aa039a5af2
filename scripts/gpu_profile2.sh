#!/bin/bash
set -u
TAG=${1:-prof2}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
NCU="ncu --clock-control none"
$NCU --set full --import-source on --kernel-name-base mangled -k regex:diffuse_packed_kernelILi6ELb0 -s 2 -c 1 -o "$OUT/dpk_t6" -f python scripts/profile_step.py 0 1 > "$OUT/dpk_t6.log" 2>&1
PFS_DIFFUSE_DEPTH=4 $NCU --set full --import-source on --kernel-name-base mangled -k regex:diffuse_packed_kernelILi4ELb0 -s 2 -c 1 -o "$OUT/dpk_t4" -f python scripts/profile_step.py 0 1 > "$OUT/dpk_t4.log" 2>&1
ls -la "$OUT"
