"""Step diagnostics (pfs_step_norms / pfs_slab_step_norms): warp-shuffle reductions on the GPU, all-reduced
over the slab ring, against float64 numpy on the oracle's post-state.  These are sums, so the comparison is
relative (1e-12: both sides accumulate in double); the maximum is exact."""
import numpy as np
import pytest

import oracle
import probabilistic_fluid_simulation_b200 as pfs
from gpu_util import to_dev
from probabilistic_fluid_simulation_b200 import fixtures
from probabilistic_fluid_simulation_b200.slab import SlabRing

pytestmark = pytest.mark.gpu


def _want(vp, vt):
    d = vp[..., 3].astype(np.float64)
    r = vt[..., 2].astype(np.float64) - vp[..., 2].astype(np.float64)
    u, v = vp[..., 0].astype(np.float64), vp[..., 1].astype(np.float64)
    return {"div_l2": np.sqrt((d * d).sum()), "pressure_update_l2": np.sqrt((r * r).sum()),
            "velocity_l2": np.sqrt((u * u + v * v).sum()), "speed_max": float(max(np.abs(vp[..., 0]).max(), np.abs(vp[..., 1]).max()))}


@pytest.mark.parametrize("shape", [(96, 128), (37, 29), (512, 768)])
def test_step_norms_match_numpy(shape):
    h, w = shape
    vel = fixtures.random_velocity_bytes(h, w, 3)
    vp, vt, _, _ = fixtures.make_state(vel)
    fv, ft = pfs.vp_field(to_dev(vp)), pfs.vp_field(to_dev(vt))
    pfs.simulate_fluid_step(fv, ft, 0.5, 0.002, 10, 10)
    got = pfs.step_norms(fv, ft)
    vp, vt = oracle.Oracle(10, 10).simulate_fluid_step(vp, vt, 0.5, 0.002)
    want = _want(vp, vt)
    for k in ("div_l2", "pressure_update_l2", "velocity_l2"):
        assert got[k] == pytest.approx(want[k], rel=1e-12), k
    assert got["speed_max"] == want["speed_max"]
    again = pfs.step_norms(fv, ft)
    assert again == got                                   # fixed-order reduction: reproducible


def test_pressure_update_shrinks_with_more_sweeps():
    h, w = 128, 128
    vp0, vt0, _, _ = fixtures.make_state(fixtures.smooth_velocity_bytes(h, w))
    res = []
    for n in (4, 16, 64, 256):
        fv, ft = pfs.vp_field(to_dev(vp0)), pfs.vp_field(to_dev(vt0))
        pfs.simulate_fluid_step(fv, ft, 0.5, 0.002, 4, n)
        res.append(pfs.step_norms(fv, ft)["pressure_update_l2"])
    assert res[0] > res[1] > res[2] > res[3] > 0


@pytest.mark.parametrize("nranks", [1, 3])
def test_ring_norms_equal_single_gpu_norms(nranks):
    h, w = 96, 128
    vel = fixtures.random_velocity_bytes(h, w, 5)
    vp, vt, _, _ = fixtures.make_state(vel)
    fv, ft = pfs.vp_field(to_dev(vp)), pfs.vp_field(to_dev(vt))
    pfs.simulate_fluid_step(fv, ft, 0.5, 0.002, 10, 10)
    single = pfs.step_norms(fv, ft)
    ring = SlabRing(nranks, w, h)
    bv, bt = ring.split(vp), ring.split(vt)
    ring.simulate_fluid_step(bv, bt, 0.5, 0.002, 10, 10)
    got = ring.step_norms(bv, bt)
    ring.close()
    for k in ("div_l2", "pressure_update_l2", "velocity_l2"):
        assert got[k] == pytest.approx(single[k], rel=1e-12), k
    assert got["speed_max"] == single["speed_max"]
