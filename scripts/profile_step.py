#!/usr/bin/env python
"""A few 4096^2 timesteps (100+100 sweeps) for ncu.  Usage: profile_step.py [depth] [steps] [size] [n] [ctx|stateless]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import probabilistic_fluid_simulation_b200 as pfs

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 0
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
size = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
n = int(sys.argv[4]) if len(sys.argv) > 4 else 100
mode = sys.argv[5] if len(sys.argv) > 5 else "ctx"
pfs.set_fuse_depth(depth)
vp, vtmp, image, itmp = bench.make_inputs(size, size)
fv, ft, fi, fm = (pfs.vp_field(torch.from_numpy(x).cuda()) for x in (vp, vtmp, image, itmp))
if mode == "ctx":
    ctx = pfs.FluidContext(size, size, size, size)
    ctx.upload(fv.data, ft.data, fi.data)
    ctx.step(steps, bench.DT, bench.VISC, n, n)
else:
    for _ in range(steps):
        pfs.simulate_fluid_step(fv, ft, bench.DT, bench.VISC, n, n)
        pfs.advect_color_step(fi, fm, fv, bench.DT)
torch.cuda.synchronize()
print("done", pfs.kernel_launch_count())
