import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import oracle, probabilistic_fluid_simulation_b200 as pfs
from gpu_util import to_dev, to_host
h, w = 160, 384
rng = np.random.default_rng(41)
a = (rng.standard_normal((h, w, 4)) * 0.5).astype(np.float32)
b = rng.standard_normal((h, w, 4)).astype(np.float32)
pfs.set_fuse_depth(2)
x, y = a.copy(), b.copy()
fa, fb = pfs.vp_field(to_dev(x)), pfs.vp_field(to_dev(y))
pfs.diffuse(fa, fb, 0.02, 1.5, 2)
ra, rb = oracle.Oracle().diffuse(x, y, 0.02, 1.5, 2)
g = to_host(fa.data); want = ra
ok = (g.view(np.uint32) == want.view(np.uint32))[..., 0]
print("rows fully ok:", np.nonzero(ok.all(axis=1))[0][:20], "count", int(ok.all(axis=1).sum()))
print("row 10 ok cols:", np.nonzero(ok[10])[0][:64])
print("col 20 ok rows:", np.nonzero(ok[:, 20])[0][:64])
bad = np.argwhere(~ok)[:4]
for r, c in bad:
    print(r, c, g[r, c, :2], want[r, c, :2], "input", a[r, c, :2])
