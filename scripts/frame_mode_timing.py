#!/usr/bin/env python
"""Frame mode of the driver (fluidsim_b200 <n> ... <out_dir>): wall time of the whole run with the serial
copy-and-encode path (PFS_FRAME_WRITERS=0, what the reference's loop does) and with the threaded frame ring."""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from probabilistic_fluid_simulation_b200 import fixtures, pngio  # noqa: E402

EXE = os.path.join(ROOT, "probabilistic_fluid_simulation_b200", "host", "build", "fluidsim_b200")
size, steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, int(sys.argv[2]) if len(sys.argv) > 2 else 24
out = {"grid": [size, size], "image": [size, size], "steps": steps, "sweeps": "30+30", "runs": {}}
with tempfile.TemporaryDirectory() as d:
    pngio.write_rgba8(d + "/vel.png", fixtures.smooth_velocity_bytes(size, size))
    pngio.write_rgba8(d + "/img.png", fixtures.random_image_bytes(size, size, 4321))
    for writers in ("0", "1", "2", "4", "8"):
        os.makedirs(f"{d}/f{writers}", exist_ok=True)
        t0 = time.time()
        r = subprocess.run([EXE, str(steps), "0.1", "0.001", d + "/img.png", d + "/vel.png", f"{d}/f{writers}"],
                           capture_output=True, text=True, env=dict(os.environ, PFS_FRAME_WRITERS=writers))
        wall = time.time() - t0
        assert r.returncode == 0, r.stderr
        us = int(r.stdout.strip().splitlines()[-1].split()[3])
        crc = pngio.crc32(pngio.read_rgba8(f"{d}/f{writers}/{steps - 1}.png"))
        out["runs"][writers] = {"loop_us": us, "ms_per_frame": us / 1000 / steps, "process_wall_s": round(wall, 3),
                                "last_frame_crc32": crc}
    r = subprocess.run([EXE, str(steps), "0.1", "0.001", d + "/img.png", d + "/vel.png"], capture_output=True, text=True)
    out["timing_mode_loop_us"] = int(r.stdout.strip().splitlines()[-1].split()[3])
assert len({v["last_frame_crc32"] for v in out["runs"].values()}) == 1
print(json.dumps(out))
