"""PNG pairs and parameters of the end-to-end driver cases (shared by scripts/make_driver_golden.py, the CPU test of the
reference's own program and the GPU tests of the two B200 drivers).  Inputs are written as 8-bit RGBA PNGs from the
committed byte arrays in tests/golden/, which libpng decodes to the same bytes."""
from __future__ import annotations

import json
import os
import subprocess

import numpy as np

from golden_util import GOLD
from probabilistic_fluid_simulation_b200 import pngio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CPU = os.path.join(ROOT, "oracle", "_ref", "fluidsim_cpu")              # reference main.cpp + fluid.cpp, unmodified
REF_MAIN_B200 = os.path.join(ROOT, "oracle", "_ref", "fluidsim_cuda_b200")  # reference main.cpp -DUSE_CUDA + libfluid_b200.so
B200 = os.path.join(ROOT, "probabilistic_fluid_simulation_b200", "host", "build", "fluidsim_b200")
GOLDEN_JSON = os.path.join(GOLD, "driver_frames.json")

CASES = {
    # name: (velocity npz, velocity crop (h, w) or None, image npz, image crop, steps, dt, viscosity)
    "perlin64_baboon_crop": ("png_perlin_t0_64.npz", None, "png_baboon.npz", (96, 160), 4, "0.5", "0.001"),
    "voronoi256_tulips": ("png_voronoi_256.npz", None, "png_tulips.npz", None, 3, "10", "0"),
    "circular128_baboon": ("png_circular_128.npz", None, "png_baboon.npz", None, 3, "0.1", "0.001"),
}


def write_inputs(name: str, directory: str):
    vel_f, vel_crop, img_f, img_crop, steps, dt, visc = CASES[name]
    vel = np.load(os.path.join(GOLD, vel_f))["rgba"]
    img = np.load(os.path.join(GOLD, img_f))["rgba"]
    if vel_crop:
        vel = vel[:vel_crop[0], :vel_crop[1]].copy()
    if img_crop:
        img = img[:img_crop[0], :img_crop[1]].copy()
    vel_png, img_png = os.path.join(directory, "vel.png"), os.path.join(directory, "img.png")
    pngio.write_rgba8(vel_png, vel)
    pngio.write_rgba8(img_png, img)
    return vel, img, vel_png, img_png, steps, dt, visc


def run_driver(exe: str, name: str, directory: str, env: dict | None = None):
    """-> (stdout lines, [crc32 of frame i])."""
    vel, img, vel_png, img_png, steps, dt, visc = write_inputs(name, directory)
    out = os.path.join(directory, "frames")
    os.makedirs(out, exist_ok=True)
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([exe, str(steps), dt, visc, img_png, vel_png, out], capture_output=True, text=True, timeout=600,
                       env=e)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    crcs = [pngio.crc32(pngio.read_rgba8(os.path.join(out, f"{i}.png"))) for i in range(steps)]
    return lines, crcs, out


def expected_lines(name: str, out: str, vel_shape):
    steps, dt = CASES[name][4], CASES[name][5]
    # main.cpp:214-215 prints the float parsed by atof with operator<< (6 significant digits)
    head = f"Simulating [{vel_shape[0]} x {vel_shape[1]}] domain for {steps} timesteps at dt={float(dt):g}..."
    return [head] + [f"[{i}] Writing to : {out}/{i}.png" for i in range(steps)]


def load_driver_golden() -> dict:
    with open(GOLDEN_JSON) as f:
        return json.load(f)
