#!/usr/bin/env python
"""pfs_compute_pressure_adaptive at the headline size: sweeps used and time against the fixed 100-sweep solve."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import probabilistic_fluid_simulation_b200 as pfs  # noqa: E402
from probabilistic_fluid_simulation_b200 import fixtures  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
vel = fixtures.smooth_velocity_bytes(size, size)
vp0, vtmp0, _, _ = fixtures.make_state(vel, None)
dt = 0.1


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out


rows = []
# a developed state: three full steps (100 + 100 sweeps), so the pressure guess is last step's p_N as in a running simulation
fv, ft = pfs.vp_field(torch.from_numpy(vp0).cuda()), pfs.vp_field(torch.from_numpy(vtmp0).cuda())
for _ in range(3):
    pfs.simulate_fluid_step(fv, ft, dt, 0.001, 100, 100)
d_vp0, d_tmp0 = fv.data.clone(), ft.data.clone()
def fixed():
    fa, fb = pfs.vp_field(d_vp0.clone()), pfs.vp_field(d_tmp0.clone())
    pfs.computePressure(fa, fb, dt, 100)
ms_fixed, _ = timed(fixed)
def clones():
    d_vp0.clone(); d_tmp0.clone()
ms_clone, _ = timed(clones)
rows.append({"mode": "fixed 100 sweeps", "ms": ms_fixed - ms_clone})
for tol in (1e-2, 3e-3, 1e-3, 3e-4):
    for every in (8, 16, 32):
        def adaptive():
            fa, fb = pfs.vp_field(d_vp0.clone()), pfs.vp_field(d_tmp0.clone())
            return pfs.computePressureAdaptive(fa, fb, dt, tol, 400, every)
        ms, (n, rms) = timed(adaptive)
        rows.append({"mode": "adaptive", "tol": tol, "check_every": every, "sweeps": n, "update_rms": rms, "ms": ms - ms_clone})
print(json.dumps({"grid": [size, size], "dt": dt, "input": "fixtures.smooth_velocity_bytes after 3 steps of 100+100 sweeps",
                  "rows": rows}))
