#!/usr/bin/env python
"""BASELINE configs[3], strong-scaling reference point: the full 16384 x 16384 grid (100 + 100 sweeps) on ONE
B200 (4 interleaved buffers = 16 GiB + 7 GiB of planes).  Inputs are built band by band so the host never
holds more than one 2048-row band.  Prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import probabilistic_fluid_simulation_b200 as pfs  # noqa: E402

W = H = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100
STEPS = int(sys.argv[3]) if len(sys.argv) > 3 else 4
BAND = 2048

bufs = [torch.empty((H, W, 4), dtype=torch.float32, device="cuda") for _ in range(4)]
t0 = time.time()
for r0 in range(0, H, BAND):
    band = bench.make_inputs(H, W, rows=(r0, min(H, r0 + BAND)))
    for dst, src in zip(bufs, band):
        dst[r0:r0 + src.shape[0]].copy_(torch.from_numpy(src))
gen_s = time.time() - t0
fv, ft, fi, fm = (pfs.vp_field(b) for b in bufs)


def step():
    pfs.simulate_fluid_step(fv, ft, bench.DT, bench.VISC, N, N)
    pfs.advect_color_step(fi, fm, fv, bench.DT)


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(STEPS):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / STEPS
norms = pfs.step_norms(fv, ft)
print(json.dumps({"workload": f"{W}x{H} grid + image, {N}+{N} sweeps, 1 GPU", "ms_per_step": ms,
                  "value_cell_updates_per_s": W * H * N / (ms * 1e-3), "input_generation_s": gen_s,
                  "gpu_mem_gib": torch.cuda.max_memory_allocated() / 2 ** 30, "norms": norms}), flush=True)
