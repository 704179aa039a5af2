#!/usr/bin/env python3
"""the reference's CPU path as its own CMakeLists.txt builds it (no CMAKE_BUILD_TYPE, no optimisation
flag: -O0) next to the -O2 -ffp-contract=off build used as cpu_baseline, once, at 1024^2 with 50+50 sweeps
(BASELINE configs[1]).  One host thread (fluid.cpp is single-threaded).  Both builds must produce the same bits.
Usage: python scripts/cpu_as_shipped.py [out.json]   (needs /root/reference; writes only under oracle/_ref/)"""
import ctypes
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from probabilistic_fluid_simulation_b200 import fixtures  # noqa: E402

REF = "/root/reference/src/fluid.cpp"
N, W, H, STEPS = 50, 1024, 1024, 2


def build(opt):
    out = os.path.join(ROOT, "oracle", "_ref", f"libfluid_ref_{N}_{opt.strip('-')}.so")
    cmd = [os.environ.get("CXX", "g++"), "-std=c++11", opt, "-ffp-contract=off", "-fPIC", "-w", f"-DPFS_REF_ITERS={N}",
           f"-DPFS_REF_FLUID_CPP=\"{REF}\"", "-shared", "-o", out, os.path.join(ROOT, "oracle", "ref_wrap.cpp"), "-lm"]
    subprocess.run(cmd, check=True)
    return ctypes.CDLL(out)


def run(lib):
    vel = fixtures.smooth_velocity_bytes(H, W)
    img = fixtures.random_image_bytes(H, W)
    vp, vtmp, image, itmp = fixtures.make_state(vel, img)
    pv, pi = oracle._Pair(vp, vtmp), oracle._Pair(image, itmp)
    F = ctypes.POINTER(oracle._Field)
    lib.ref_run_steps.argtypes = [F, F, F, F, ctypes.c_float, ctypes.c_float, ctypes.c_int]
    t0 = time.perf_counter()
    lib.ref_run_steps(pv.fa, pv.fb, pi.fa, pi.fb, 0.1, 0.001, STEPS)
    dt = (time.perf_counter() - t0) / STEPS
    return dt, [a.copy() for a in (*pv.resolve(), *pi.resolve())]


def main():
    res = {"workload": f"{W}x{H} grid + image, {N}+{N} sweeps, {STEPS} timesteps, 1 thread", "host_cores": os.cpu_count()}
    outs = {}
    for opt in ("-O0", "-O2"):
        t, outs[opt] = run(build(opt))
        res[opt] = {"s_per_step": t, "pressure_cell_updates_per_s": W * H * N / t}
        print(opt, res[opt], flush=True)
    res["bit_identical"] = all(np.array_equal(a.view(np.uint32), b.view(np.uint32)) for a, b in zip(outs["-O0"], outs["-O2"]))
    print("bit identical:", res["bit_identical"])
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
