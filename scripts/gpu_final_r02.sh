#!/bin/bash
# Round-2 end check on one B200: GPU suite, smoke(), default bench line + reference arm, the 1024^2 config, stateless phase
# times, the dispatch-cost microbenchmark, and the ncu evidence for profiles/ (launch list of the bench command, DRAM bytes
# per launch -> traffic.json, full-set captures of the step's kernels summarised to text).
# Usage (GPU box, repo root): bash scripts/gpu_final_r02.sh <tag>     -> gpurun_out/<tag>/
set -u
TAG=${1:-final2}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
echo "== pytest -m gpu" | tee "$OUT/summary.txt"
timeout 1200 python -m pytest tests -x -q -m gpu > "$OUT/pytest_gpu.log" 2>&1
echo "pytest exit $?" | tee -a "$OUT/summary.txt"
tail -3 "$OUT/pytest_gpu.log" | tee -a "$OUT/summary.txt"
echo "== smoke" | tee -a "$OUT/summary.txt"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1
echo "smoke exit $?" | tee -a "$OUT/summary.txt"; tail -2 "$OUT/smoke.log" | tee -a "$OUT/summary.txt"
show() { python -c "
import json
d=json.loads([l for l in open('$1') if l.startswith('{')][-1])
print('ms/step %.4f'%d['ms_per_step'], 'value %.4g'%d['value'], {k: round(v,4) for k,v in d['phases_ms'].items()}, 'launches', d['gpu_launches'], 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],2), 'stateless', d.get('stateless_entry_points') and round(d['stateless_entry_points']['ms_per_step'],4), 'roofline', d['roofline'] and round(d['roofline']['frac'],3))" | tee -a "$OUT/summary.txt"; }
echo "== bench default" | tee -a "$OUT/summary.txt"
timeout 900 python bench.py > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"; show "$OUT/bench_default.json"
echo "== bench 1024^2 50+50 (configs[1])" | tee -a "$OUT/summary.txt"
timeout 600 python bench.py --width 1024 --height 1024 --iters 50 --steps 100 --warmup 20 > "$OUT/bench_cfg2_1024_n50.json" 2> "$OUT/bench_cfg2.err"; show "$OUT/bench_cfg2_1024_n50.json"
echo "== bench --stateless (caller-owned interleaved buffers as the timed path)" | tee -a "$OUT/summary.txt"
timeout 600 python bench.py --stateless --steps 20 --warmup 3 --no-unit > "$OUT/bench_stateless.json" 2> "$OUT/bench_stateless.err"; show "$OUT/bench_stateless.json"
echo "== bench --impl reference" | tee -a "$OUT/summary.txt"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > "$OUT/bench_reference_arm.json" 2> "$OUT/bench_reference_arm.err"; echo "exit $?" | tee -a "$OUT/summary.txt"; cut -c1-300 "$OUT/bench_reference_arm.json" | tee -a "$OUT/summary.txt"
echo "== dispatch_cost microbenchmark" | tee -a "$OUT/summary.txt"
timeout 120 scripts/ubench/dispatch_cost > "$OUT/dispatch_cost.txt" 2>&1; tail -5 "$OUT/dispatch_cost.txt" | tee -a "$OUT/summary.txt"
echo "== ncu" | tee -a "$OUT/summary.txt"
NCU="ncu --clock-control none"
PFS_STEP_GRAPH=0 timeout 900 $NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file "$OUT/bench_py_launches_ncu.csv" python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-unit > "$OUT/bench_under_ncu.log" 2>&1
timeout 900 python scripts/make_traffic.py --out "$OUT/traffic.json" --tag r02 > "$OUT/traffic.log" 2>&1; tail -2 "$OUT/traffic.log" | tee -a "$OUT/summary.txt"
export PFS_STEP_GRAPH=0
timeout 600 $NCU --set full --import-source on --kernel-name-base mangled -k regex:diffuse_packed_kernelILi5ELi2ELb1ELb0 -s 3 -c 1 -o "$OUT/diffuse_t5" -f python scripts/profile_step.py 0 1 > "$OUT/ncu_diffuse.log" 2>&1
timeout 600 $NCU --set full --import-source on --kernel-name-base mangled -k regex:fused_sweeps_kernelILi8ELi3ELb0 -s 3 -c 1 -o "$OUT/pressure_t8" -f python scripts/profile_step.py 0 1 > "$OUT/ncu_pressure.log" 2>&1
timeout 600 $NCU --set full --import-source on --kernel-name-base mangled -k regex:"advect_kernel|divergence_kernel|project_uv_kernel|advect_color_kernel" -s 0 -c 4 -o "$OUT/streaming" -f python scripts/profile_step.py 0 1 > "$OUT/ncu_streaming.log" 2>&1
python scripts/summarize_ncu.py "$OUT/ncu_fused_kernels.txt" "$OUT/diffuse_t5.ncu-rep" "$OUT/pressure_t8.ncu-rep" > /dev/null 2>&1
python scripts/summarize_ncu.py "$OUT/ncu_streaming_kernels.txt" "$OUT/streaming.ncu-rep" > /dev/null 2>&1
ls -la "$OUT" | tee -a "$OUT/summary.txt"
