import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import oracle, probabilistic_fluid_simulation_b200 as pfs
from gpu_util import to_dev, to_host
f32 = np.float32
h, w = 160, 384
rng = np.random.default_rng(41)
a = (rng.standard_normal((h, w, 4)) * 0.5).astype(np.float32)
b = rng.standard_normal((h, w, 4)).astype(np.float32)
a[..., :2] *= f32(1e-36)
visc, dt = f32(0.02), f32(1.5)
alpha = f32(visc * dt); beta = f32(1.0 + 4.0 * float(alpha))
pfs.set_fuse_depth(2)
x, y = a.copy(), b.copy()
fa, fb = pfs.vp_field(to_dev(x)), pfs.vp_field(to_dev(y))
pfs.diffuse(fa, fb, visc, dt, 2)
g = to_host(fa.data)
x1, y1 = a.copy(), b.copy()
r1a, r1b = oracle.Oracle().diffuse(x1, y1, visc, dt, 1)      # one sweep: result in vp_out
one = r1b
x2, y2 = a.copy(), b.copy()
ra, rb = oracle.Oracle().diffuse(x2, y2, visc, dt, 2)
want = ra
bad = np.argwhere((g.view(np.uint32) != want.view(np.uint32))[..., :2])
print(len(bad), "mismatches; alpha", float(alpha).hex(), "beta", float(beta).hex())
FLT_MIN = np.finfo(np.float32).tiny
def ftz(v):
    return f32(0.0) * np.sign(v) if abs(v) < FLT_MIN else v
for (r, c, k) in bad[:5]:
    L, R = one[r, (c - 1) % w, k], one[r, (c + 1) % w, k]
    T, B = one[(r - 1) % h, c, k], one[(r + 1) % h, c, k]
    C = one[r, c, k]
    aL, aR, aT, aB = (f32(alpha * v) for v in (L, R, T, B))
    num = f32(f32(f32(f32(aL + aR) + aT) + aB) + C)
    q = f32(num / beta)
    # variants
    fL, fR, fT, fB = (ftz(v) for v in (aL, aR, aT, aB))
    num_f = f32(f32(f32(f32(fL + fR) + fT) + fB) + C)
    q_f = f32(num_f / beta)
    q_rcp = f32(num * f32(1.0 / beta))
    print(f"cell {r},{c},{k}: gpu {float(g[r,c,k]).hex()} oracle {float(want[r,c,k]).hex()} emu {float(q).hex()} | products-FTZ {float(q_f).hex()} | a*rcp {float(q_rcp).hex()} | products {[float(v).hex() for v in (aL,aR,aT,aB)]} num {float(num).hex()}")
