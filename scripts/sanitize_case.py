#!/usr/bin/env python
"""Small workload for compute-sanitizer: every kernel of the step (fused + packed + basic + repair path),
the operator entry points and the in-process slab ring."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
import probabilistic_fluid_simulation_b200 as pfs
from probabilistic_fluid_simulation_b200.slab import SlabRing

h, w = 320, 384
vp, vtmp, image, itmp = bench.make_inputs(h, w)
fv, ft, fi, fm = (pfs.vp_field(torch.from_numpy(x).cuda()) for x in (vp, vtmp, image, itmp))
for s in range(2):
    pfs.simulate_fluid_step(fv, ft, 0.5, 0.002, 30, 30, sigma=0.01, seed=5, step=s)
    pfs.advect_color_step(fi, fm, fv, 0.5)
z = np.zeros_like(vp); z[..., 2:] = vp[..., 2:]
fa, fb = pfs.vp_field(torch.from_numpy(z).cuda()), pfs.vp_field(torch.from_numpy(vtmp.copy()).cuda())
pfs.diffuse(fa, fb, 0.01, 1.0, 9)          # all-zero velocities: guard + repair launch
pfs.computePressure(fa, fb, 0.3, 7)
pfs.subtractPressureGradient(fa, fb, 0.3)
pfs.advect(fa, fb, 2.0)
odd = pfs.vp_field(torch.rand(37, 29, 4, device="cuda")), pfs.vp_field(torch.rand(37, 29, 4, device="cuda"))
pfs.simulate_fluid_step(odd[0], odd[1], 1.0, 0.01, 5, 6)      # scalar (W % 4 != 0) kernels
ring = SlabRing(3, w, h, w, h)
bv, bt = ring.split(vp), ring.split(vtmp)
bi, bm = ring.split(image, image=True), ring.split(itmp, image=True)
ring.simulate_fluid_step(bv, bt, 0.5, 0.002, 20, 20)
ring.advect_color_step(bi, bm, bv, 0.5)
ring.check()
# round 2: forces fused into the last diffusion pass, persistent context (+ frame bytes from the advecting kernel), resident
# slab state, adaptive solves (single device, ring), red-black SOR
forces = torch.rand(h, w, 4, device="cuda")
pfs.simulate_fluid_step(fv, ft, 0.5, 0.002, 11, 9, forces=forces)
ctx = pfs.FluidContext(w, h, w, h)
ctx.upload(fv.data, ft.data, fi.data)
ctx.step(3, 0.5, 0.002, 12, 13)
frame = torch.zeros(h, w, 4, dtype=torch.uint8, device="cuda")
ctx.simulate_fluid_step(0.5, 0.002, 12, 13)
ctx.advect_color_step(0.5, frame_out=frame)
ctx.download()
ctx.close()
ring.upload(bv, bt, bi)
ring.step(2, 0.5, 0.002, 20, 20)
ring.download(bv, bt, bi)
ring.check()
ring.compute_pressure_adaptive(bv, bt, 0.5, 1e-30, 24, 8)
ring.check()
ga, gb = pfs.vp_field(torch.from_numpy(vp.copy()).cuda()), pfs.vp_field(torch.from_numpy(vtmp.copy()).cuda())
pfs.computePressureAdaptive(ga, gb, 0.5, 1e-30, 24, 8)
pfs.computePressureSOR(ga, gb, 0.5, 1.8, 1e-30, 12, 4)
torch.cuda.synchronize()
print("sanitize case done", pfs.kernel_launch_count())
