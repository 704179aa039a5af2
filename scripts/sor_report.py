"""Sweeps to a given update tolerance: the reference's Jacobi pressure solve (parity path, pfs_compute_pressure_adaptive)
against the red-black SOR mode (pfs_compute_pressure_sor, NOT a parity path), on the smooth bench field.
usage: python scripts/sor_report.py OUT.txt [size ...]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import probabilistic_fluid_simulation_b200 as pfs  # noqa: E402
from probabilistic_fluid_simulation_b200 import fixtures  # noqa: E402


def timed(fn):
    torch.cuda.synchronize()
    t = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, (time.perf_counter() - t) * 1e3


def main():
    out_path = sys.argv[1]
    sizes = [int(a) for a in sys.argv[2:]] or [512, 1024, 4096]
    lines = ["# sweeps until rms(p_n - p_{n-1}) <= tol; Jacobi = the reference's update (bit-exact path), SOR = red-black, not parity",
             "# size  tol      omega  jacobi_sweeps jacobi_ms  sor_sweeps sor_ms  residual_jacobi residual_sor"]
    for n in sizes:
        vp, vtmp, _, _ = fixtures.make_state(fixtures.smooth_velocity_bytes(n, n), fixtures.random_image_bytes(8, 8, 1))
        dt = 1.0
        for tol in (1e-3, 1e-4):
            cap = 200000

            def jac():
                a, b = pfs.vp_field(torch.from_numpy(vp).cuda()), pfs.vp_field(torch.from_numpy(vtmp).cuda())
                k, _ = pfs.computePressureAdaptive(a, b, dt, tol, cap, 50)
                return k, (b if k % 2 else a).data
            jac()                        # warm-up: graph captures, first-use allocations
            (kj, pj), tj = timed(jac)
            for omega in (1.5, 1.8, 1.9, 1.95):
                def sor():
                    a, b = pfs.vp_field(torch.from_numpy(vp).cuda()), pfs.vp_field(torch.from_numpy(vtmp).cuda())
                    k, _ = pfs.computePressureSOR(a, b, dt, omega, tol, cap, 10)
                    return k, b.data
                (ks, ps), ts = timed(sor)

                def resid(f):
                    p, d = f[..., 2].double(), ps[..., 3].double()
                    s = p.roll(1, 0) + p.roll(-1, 0) + p.roll(1, 1) + p.roll(-1, 1) + d
                    return float((0.25 * s - p).pow(2).mean().sqrt())
                lines.append(f"{n:5d}  {tol:.0e}  {omega:4.2f}  {kj:8d} {tj:9.2f}  {ks:8d} {ts:8.2f}  {resid(pj):.3e} {resid(ps):.3e}")
    open(out_path, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
