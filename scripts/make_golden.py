#!/usr/bin/env python
"""Generate tests/golden/ from the UNMODIFIED reference (oracle/_ref/libfluid_ref_<N>.so).

Run in the build container (needs /root/reference for the compiled reference and the bundled PNGs):

    make -C oracle && python scripts/make_golden.py

Writes
  tests/golden/png_<name>.npz      decoded (libpng-exact, utils.hpp:49-66) RGBA bytes of bundled PNGs
  tests/golden/golden.json         per case: how to rebuild the inputs, the parameters, and the
                                   64-bit FNV-1a hash of every channel of vp / vtmp / image after the
                                   run, plus a few float64 sums as a coarse check.

Every case is driven exactly like src/main.cpp:143-195,219-240.  tests/ rebuild the inputs from
the recorded spec (formulas / seeds in probabilistic_fluid_simulation_b200/fixtures.py, or the npz)
and compare (a) the C restatement on CPU and (b) the CUDA path on GPU against these hashes.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))

import oracle  # noqa: E402
from probabilistic_fluid_simulation_b200 import pngio as png_exact  # noqa: E402
from probabilistic_fluid_simulation_b200 import fixtures  # noqa: E402
from tests.golden_util import build_inputs  # noqa: E402

REF = "/root/reference/inputs"
GOLD = os.path.join(ROOT, "tests", "golden")

PNGS = {
    # fixture file stem: bundled PNG (decoded libpng-exact, stored as uint8 RGBA)
    "png_baboon": "images/baboon.png",
    "png_tulips": "images/tulips.png",
    "png_perlin_t0_256": "velocity_fields/perlin/T0/PerlinRG_256.png",
    "png_perlin_t0_64": "velocity_fields/perlin/T0/PerlinRG_64.png",
    "png_voronoi_256": "velocity_fields/voronoi/T0/VoronoiRG_256.png",
    "png_circular_128": "velocity_fields/functions/T3/CircularFields_128.png",
    "png_solid_r64": "velocity_fields/solid/T0/R64.png",
    "png_peppers": "images/peppers.png",
    "png_checker16_1024": "images/Checker16_1024.png",
    "png_perlin_t1_512": "velocity_fields/perlin/T1/PerlinRG_512.png",
    "png_ramp_t0_1024": "velocity_fields/ramps/T0/RampRG_1024.png",
}


def _png(vel: str, img: str) -> dict:
    return {"kind": "png", "velocity": vel + ".npz", "image": img + ".npz"}


CASES = [
    # name, inputs, dt, viscosity, N, steps
    ("g1_1step", _png("png_perlin_t0_256", "png_baboon"), 0.1, 0.001, 30, 1),
    ("g1_10steps", _png("png_perlin_t0_256", "png_baboon"), 0.1, 0.001, 30, 10),
    ("g1_100steps", _png("png_perlin_t0_256", "png_baboon"), 0.1, 0.001, 30, 100),
    ("g2_formula", {"kind": "formula", "vel_hw": [64, 64], "img_hw": [128, 128]}, 50.0, 0.01, 30, 3),
    ("tulips_voronoi_5", _png("png_voronoi_256", "png_tulips"), 10.0, 0.001, 30, 5),
    ("tulips_voronoi_100", _png("png_voronoi_256", "png_tulips"), 10.0, 0.001, 30, 100),
    ("baboon_circular_100", _png("png_circular_128", "png_baboon"), 0.1, 0.0, 30, 100),
    ("perlin64_solid64_dt10", _png("png_perlin_t0_64", "png_solid_r64"), 10.0, 0.001, 30, 20),
    # BASELINE configs[0] (SURVEY.md 8d "Config 1"): the five bundled pairs at native resolution, 100 steps,
    # dt in {0.1, 10} x nu in {0, 0.001}, N = 30.  "heavy": checked on the GPU only (the CPU suite skips them).
    ("cfg1_baboon_perlin256_dt0.1_nu0", _png("png_perlin_t0_256", "png_baboon"), 0.1, 0.0, 30, 100),
    ("cfg1_baboon_perlin256_dt10_nu0.001", _png("png_perlin_t0_256", "png_baboon"), 10.0, 0.001, 30, 100),
    ("cfg1_baboon_perlin256_dt10_nu0", _png("png_perlin_t0_256", "png_baboon"), 10.0, 0.0, 30, 100),
    ("cfg1_peppers_perlin512_dt0.1_nu0.001", _png("png_perlin_t1_512", "png_peppers"), 0.1, 0.001, 30, 100),
    ("cfg1_peppers_perlin512_dt10_nu0", _png("png_perlin_t1_512", "png_peppers"), 10.0, 0.0, 30, 100),
    ("cfg1_checker_ramp1024_dt0.1_nu0.001", _png("png_ramp_t0_1024", "png_checker16_1024"), 0.1, 0.001, 30, 100),
    ("cfg1_checker_ramp1024_dt10_nu0", _png("png_ramp_t0_1024", "png_checker16_1024"), 10.0, 0.0, 30, 100),
    ("cfg1_tulips_voronoi_dt0.1_nu0", _png("png_voronoi_256", "png_tulips"), 0.1, 0.0, 30, 100),
    ("cfg1_baboon_circular_dt10_nu0.001", _png("png_circular_128", "png_baboon"), 10.0, 0.001, 30, 100),
    ("n1", {"kind": "formula", "vel_hw": [40, 48], "img_hw": [60, 96]}, 10.0, 0.05, 1, 4),
    ("n2", {"kind": "formula", "vel_hw": [40, 48], "img_hw": [60, 96]}, 10.0, 0.05, 2, 4),
    ("n3", {"kind": "formula", "vel_hw": [40, 48], "img_hw": [60, 96]}, 10.0, 0.05, 3, 4),
    ("n4", {"kind": "formula", "vel_hw": [40, 48], "img_hw": [60, 96]}, 10.0, 0.05, 4, 4),
    ("n5", {"kind": "formula", "vel_hw": [40, 48], "img_hw": [60, 96]}, 10.0, 0.05, 5, 4),
    ("odd_shape_n3", {"kind": "random", "vel_hw": [29, 37], "seed": 7, "img_hw": [41, 50], "img_kind": "random"}, 5.0, 0.01, 3, 3),
    ("odd_shape_n30", {"kind": "random", "vel_hw": [29, 37], "seed": 7, "img_hw": [41, 50], "img_kind": "random"}, 5.0, 0.01, 30, 3),
    ("big_dt_wrap", {"kind": "random", "vel_hw": [128, 128], "seed": 11, "img_hw": [128, 128], "img_kind": "random"}, 100.0, 0.001, 30, 2),
    ("huge_dt_wrap", {"kind": "formula", "vel_hw": [64, 96], "img_hw": [64, 96]}, 30000.0, 0.0001, 4, 2),
    ("zero_viscosity", {"kind": "formula", "vel_hw": [64, 64], "img_hw": [64, 64]}, 1.0, 0.0, 30, 3),
    ("wide_strip", {"kind": "smooth", "vel_hw": [16, 512], "img_hw": [16, 512]}, 2.0, 0.02, 30, 3),
    ("tall_strip", {"kind": "smooth", "vel_hw": [512, 16], "img_hw": [512, 16]}, 2.0, 0.02, 30, 3),
    ("n50_256", {"kind": "smooth", "vel_hw": [256, 256], "img_hw": [256, 256]}, 0.1, 0.001, 50, 3),
    ("n100_256", {"kind": "smooth", "vel_hw": [256, 256], "img_hw": [512, 512]}, 0.1, 0.001, 100, 3),
    ("n100_rand_320x192", {"kind": "random", "vel_hw": [192, 320], "seed": 5, "img_hw": [192, 320], "img_kind": "random"}, 3.0, 0.002, 100, 2),
    ("cfg2_1024_n50", {"kind": "smooth", "vel_hw": [1024, 1024], "img_hw": [1024, 1024], "img_kind": "random"}, 0.1, 0.001, 50, 2),
    ("cfg2_1024_n50_dt100", {"kind": "random", "vel_hw": [1024, 1024], "img_hw": [1024, 1024], "img_kind": "random"}, 100.0, 0.001, 50, 1),
    ("n100_1536x640", {"kind": "smooth", "vel_hw": [640, 1536], "img_hw": [640, 1536]}, 1.0, 0.001, 100, 1),
]


def run_case(spec, dt, visc, n, steps):
    vel, img = build_inputs(spec)
    vp, vtmp, image, itmp = fixtures.make_state(vel, img)
    ref = oracle.Reference(n)
    vp, vtmp, image, itmp = ref.run_steps(vp, vtmp, image, itmp, dt, visc, steps)
    out = {"vp": oracle.field_hashes(vp), "vtmp": oracle.field_hashes(vtmp)}
    sums = {"vp_u_sum": float(vp[..., 0].astype(np.float64).sum()),
            "vp_u_l2": float(np.sqrt((vp[..., 0].astype(np.float64) ** 2).sum())),
            "vtmp_p_sum": float(vtmp[..., 2].astype(np.float64).sum()),
            "div_l2": float(np.sqrt((vp[..., 3].astype(np.float64) ** 2).sum()))}
    if image is not None:
        out["image"] = oracle.field_hashes(image)
        sums["image_sum"] = float(image.astype(np.float64).sum())
    return out, sums


def main():
    os.makedirs(GOLD, exist_ok=True)
    crcs = {}
    for stem, path in PNGS.items():
        rgba = png_exact.read_rgba8(os.path.join(REF, path))
        crcs[stem] = {"png": path, "shape": list(rgba.shape), "crc32": png_exact.crc32(rgba)}
        np.savez_compressed(os.path.join(GOLD, f"{stem}.npz"), rgba=rgba)
    cases = []
    for name, spec, dt, visc, n, steps in CASES:
        t0 = time.time()
        hashes, sums = run_case(spec, dt, visc, n, steps)
        cases.append({"name": name, "inputs": spec, "dt": dt, "viscosity": visc, "n_iters": n, "steps": steps,
                      "heavy": bool(time.time() - t0 > 12.0), "hashes": hashes, "sums": sums})
        print(f"{name:24s} N={n:3d} steps={steps:3d}  {time.time() - t0:6.1f}s  vp.u={hashes['vp'][0]}")
    doc = {"generator": "scripts/make_golden.py", "source": "oracle/_ref (unmodified reference src/fluid.cpp)",
           "hash": "64-bit FNV-1a over the 32-bit words of one channel, row-major (SURVEY.md 4.4)",
           "png_inputs": crcs, "cases": cases}
    with open(os.path.join(GOLD, "golden.json"), "w") as f:
        json.dump(doc, f, indent=1)
    print("wrote", os.path.join(GOLD, "golden.json"))


if __name__ == "__main__":
    main()
