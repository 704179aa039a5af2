#!/usr/bin/env python3
"""PFS_DIFFUSE_UNROLL=4 (opt-in main-loop unroll of the packed diffusion kernel) against the oracle, bit for
bit, plus the time of 96 sweeps at 4096^2 -- once per value of the knob, each in its own process (the
library reads the knob once).  Usage: python scripts/unroll_check.py [out.json]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import json, sys, numpy as np, torch
sys.path.insert(0, %r); sys.path.insert(0, %r)
import oracle, probabilistic_fluid_simulation_b200 as pfs
from gpu_util import to_dev, to_host
rng = np.random.default_rng(5)
ok = True
for (h, w, n) in ((200, 512, 30), (37, 256, 6), (1, 8, 5), (129, 1024, 13), (64, 128, 100)):
    a = rng.standard_normal((h, w, 4)).astype(np.float32); b = rng.standard_normal((h, w, 4)).astype(np.float32)
    fa, fb = pfs.vp_field(to_dev(a)), pfs.vp_field(to_dev(b))
    pfs.diffuse(fa, fb, 0.001, 0.1, n)
    ra, rb = oracle.Oracle().diffuse(a, b, 0.001, 0.1, n)
    same = np.array_equal(to_host(fa.data).view(np.uint32), ra.view(np.uint32)) and \
        np.array_equal(to_host(fb.data).view(np.uint32), rb.view(np.uint32))
    ok = ok and same
x = torch.rand(4096, 4096, 4, device="cuda") * 2 - 1
y = torch.zeros_like(x)
fx, fy = pfs.vp_field(x), pfs.vp_field(y)
pfs.diffuse(fx, fy, 0.001, 0.1, 96)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    pfs.diffuse(fx, fy, 0.001, 0.1, 96)
e1.record(); torch.cuda.synchronize()
print(json.dumps({"bit_identical": bool(ok), "ms_per_96_sweeps_incl_pack_unpack": e0.elapsed_time(e1) / 5}))
''' % (ROOT, os.path.join(ROOT, "tests"))

out = {}
for unroll in ("4", "2"):
    env = dict(os.environ, PFS_DIFFUSE_UNROLL=unroll)
    r = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, env=env, timeout=120)
    line = r.stdout.strip().split("\n")[-1] if r.stdout.strip() else ""
    try:
        out["unroll_" + unroll] = json.loads(line)
    except Exception:
        out["unroll_" + unroll] = {"error": (r.stdout + r.stderr)[-600:]}
    print("unroll", unroll, out["unroll_" + unroll], flush=True)
    if len(sys.argv) > 1:
        os.makedirs(os.path.dirname(sys.argv[1]), exist_ok=True)
        json.dump(out, open(sys.argv[1], "w"))
