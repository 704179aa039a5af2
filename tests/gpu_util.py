"""Helpers for the -m gpu parity tests: move host arrays to the device and back, drive the CUDA
path through the package (which calls the C-ABI), and compare bit for bit."""
from __future__ import annotations

import numpy as np

import probabilistic_fluid_simulation_b200 as pfs


def to_dev(a: np.ndarray):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def to_host(t) -> np.ndarray:
    return t.detach().cpu().numpy()


def gpu_run_steps(vp, vtmp, image, itmp, dt, visc, n_diffuse, n_pressure, steps):
    """main.cpp:219-240 on device buffers.  Returns host copies of (vp, vtmp, image, itmp)."""
    fv, ft = pfs.vp_field(to_dev(vp)), pfs.vp_field(to_dev(vtmp))
    fi = fm = None
    if image is not None:
        fi, fm = pfs.vp_field(to_dev(image)), pfs.vp_field(to_dev(itmp))
    for _ in range(steps):
        pfs.simulate_fluid_step(fv, ft, dt, visc, n_diffuse, n_pressure)
        if fi is not None:
            pfs.advect_color_step(fi, fm, fv, dt)
    out = [to_host(fv.data), to_host(ft.data)]
    out += [to_host(fi.data), to_host(fm.data)] if fi is not None else [None, None]
    return out
