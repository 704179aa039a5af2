"""Opt-in build variants of the packed diffusion kernel must stay bit-identical to the oracle.
PFS_DIFFUSE_UNROLL=4 (four stream steps per trip of the main loop; the library reads the knob once per
process, so the check runs in a child).  First run on a B200: gpurun r3e, scripts/unroll_check.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CODE = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import oracle, probabilistic_fluid_simulation_b200 as pfs
from gpu_util import to_dev, to_host
rng = np.random.default_rng(5)
for (h, w, n) in ((200, 512, 30), (37, 256, 6), (1, 8, 5), (129, 1024, 13), (64, 128, 100)):
    a = rng.standard_normal((h, w, 4)).astype(np.float32); b = rng.standard_normal((h, w, 4)).astype(np.float32)
    fa, fb = pfs.vp_field(to_dev(a)), pfs.vp_field(to_dev(b))
    pfs.diffuse(fa, fb, 0.001, 0.1, n)
    ra, rb = oracle.Oracle().diffuse(a, b, 0.001, 0.1, n)
    assert np.array_equal(to_host(fa.data).view(np.uint32), ra.view(np.uint32)), (h, w, n)
    assert np.array_equal(to_host(fb.data).view(np.uint32), rb.view(np.uint32)), (h, w, n)
# whole timesteps: the last fused pass also stores the previous iterate
vp = (rng.random((96, 256, 4)).astype(np.float32) * 2 - 1); vt = np.tile(np.float32([-1, -1, -1, 1]), (96, 256, 1))
fv, ft = pfs.vp_field(to_dev(vp)), pfs.vp_field(to_dev(vt))
for _ in range(3):
    pfs.simulate_fluid_step(fv, ft, 0.5, 0.001, 30, 30)
want = oracle.Oracle(30, 30).run_steps(vp, vt, None, None, 0.5, 0.001, 3)
assert np.array_equal(to_host(fv.data).view(np.uint32), want[0].view(np.uint32))
assert np.array_equal(to_host(ft.data).view(np.uint32), want[1].view(np.uint32))
print("unroll variant ok")
''' % (ROOT, os.path.join(ROOT, "tests"))


def test_diffusion_main_loop_unrolled_by_four():
    env = dict(os.environ, PFS_DIFFUSE_UNROLL="4")
    r = subprocess.run([sys.executable, "-c", CODE], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "unroll variant ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
