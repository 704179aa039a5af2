import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import oracle, probabilistic_fluid_simulation_b200 as pfs
from gpu_util import to_dev, to_host
h, w = 160, 384
rng = np.random.default_rng(41)
a0 = (rng.standard_normal((h, w, 4)) * 0.5).astype(np.float32)
b = rng.standard_normal((h, w, 4)).astype(np.float32)
for scale in (1e-36, 1e-30, 1e-20, 1.0):
    a = a0.copy(); a[..., :2] *= np.float32(scale)
    for depth, n in ((2, 2), (6, 6)):
        pfs.set_fuse_depth(depth)
        x, y = a.copy(), b.copy()
        fa, fb = pfs.vp_field(to_dev(x)), pfs.vp_field(to_dev(y))
        pfs.diffuse(fa, fb, 0.02, 1.5, n)
        ra, rb = oracle.Oracle().diffuse(x, y, 0.02, 1.5, n)
        g = to_host(fa.data if n % 2 == 0 else fb.data); want = ra if n % 2 == 0 else rb
        bad = (g.view(np.uint32) != want.view(np.uint32))[..., :2]
        idx = np.argwhere(bad)
        print(f"scale {scale} depth {depth} n {n}: {len(idx)} values differ", flush=True)
        for (r, c, k) in idx[:6]:
            print("   ", r, c, k, "col%112", c % 112, float(g[r, c, k]).hex(), float(want[r, c, k]).hex(), "input", float(a[r, c, k]).hex())
