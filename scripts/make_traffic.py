#!/usr/bin/env python
"""Regenerates profiles/traffic.json -- measured DRAM bytes per launch of every kernel of a 4096^2 timestep -- from one ncu
run of scripts/profile_step.py (run it on the GPU box; bench.py reads the file for `roofline.traffic`).

    python scripts/make_traffic.py [--out profiles/traffic.json] [--tag r02]

Every entry names the kernel (the function name as ncu prints it), so tests/test_bench_contract.py can check that the
file still describes kernels that exist in the built library."""
import argparse
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PHASE_OF = (("diffuse_packed_kernel", "diffuse"), ("fused_sweeps_kernel", "pressure"), ("advect_color_kernel", "advect_color"),
            ("advect_kernel", "advect"), ("divergence_kernel", "divergence"), ("project_uv_kernel", "project"),
            ("project_pack_kernel", "project_stateless"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "traffic.json"))
    ap.add_argument("--tag", default="r02")
    ap.add_argument("--mode", default="ctx", choices=["ctx", "stateless"])
    a = ap.parse_args()
    env = dict(os.environ, PFS_STEP_GRAPH="0")
    cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum", "--clock-control", "none", "--csv",
           sys.executable, os.path.join(ROOT, "scripts", "profile_step.py"), "0", "2", "4096", "100", a.mode]
    raw = subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=ROOT).stdout
    rows = list(csv.reader(io.StringIO(raw[raw.index('"ID"'):])))
    hdr = rows[0]
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    launches = collections.defaultdict(lambda: collections.defaultdict(float))
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3}.get(r[ui], 1.0)   # bytes / microseconds
        launches[(r[0], r[ki])][r[mi]] = v * scale
    per_kernel = collections.defaultdict(list)
    for (_, name), m in launches.items():
        per_kernel[name].append(m)
    out = {"source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none on scripts/profile_step.py "
                     f"(2 timesteps of 4096^2, 100+100 sweeps, {a.mode} path, eager launches); {a.tag}; per launch, averaged over the launches of a kernel",
           "kernels": {}}
    for name, ms in sorted(per_kernel.items()):
        n = len(ms)
        rd = sum(m.get("dram__bytes_read.sum", 0.0) for m in ms) / n
        wr = sum(m.get("dram__bytes_write.sum", 0.0) for m in ms) / n
        us = sum(m.get("gpu__time_duration.sum", 0.0) for m in ms) / n
        out["kernels"][name] = {"launches": n, "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr, "us_per_launch_under_ncu": us}
    # phase -> the kernel with the most launches among those of the phase (the full-depth pass of the two sweep loops)
    for key, phase in PHASE_OF:
        cands = [(v["launches"], k) for k, v in out["kernels"].items() if key in k]
        if cands:
            k = max(cands)[1]
            out[phase] = {"kernel": k, "dram_bytes_per_launch": out["kernels"][k]["dram_bytes_per_launch"]}
    json.dump(out, open(a.out, "w"), indent=1)
    print("wrote", a.out, len(out["kernels"]), "kernels")


if __name__ == "__main__":
    main()
