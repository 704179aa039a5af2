"""GPU parity, part 3: BASELINE.json's full sizes.

* 4096^2, 100+100 sweeps (the roofline headline): one whole timestep against the CPU oracle, bit for
  bit (the oracle needs ~35-60 s on one host core; it runs once).
* size-independent properties at the same size: independence from the temporal-blocking depth,
  exact translation covariance of the stencil phases on the periodic grid, exact mean-preservation
  structure (constant pressure with zero divergence is a fixed point), no non-finite values.
"""
import numpy as np
import pytest

import oracle
import probabilistic_fluid_simulation_b200 as pfs
from golden_util import assert_bit_equal
from gpu_util import to_dev, to_host
from probabilistic_fluid_simulation_b200 import fixtures

pytestmark = pytest.mark.gpu


def _state(h, w, seed=1234):
    vel = fixtures.smooth_velocity_bytes(h, w)
    rng = np.random.default_rng(seed)
    vel[..., :2] = np.clip(vel[..., :2].astype(np.int16) + rng.integers(-6, 7, size=(h, w, 2)), 0, 255).astype(np.uint8)
    img = fixtures.random_image_bytes(h, w, seed + 1)
    return fixtures.make_state(vel, img)


def test_4096_n100_one_step_bit_exact_vs_oracle():
    h = w = 4096
    vp, vtmp, image, itmp = _state(h, w)
    fv, ft, fi, fm = (pfs.vp_field(to_dev(x)) for x in (vp, vtmp, image, itmp))
    pfs.simulate_fluid_step(fv, ft, 0.1, 0.001, 100, 100)
    pfs.advect_color_step(fi, fm, fv, 0.1)
    got = [to_host(f.data) for f in (fv, ft, fi)]
    want = oracle.Oracle(100, 100).run_steps(vp, vtmp, image, itmp, 0.1, 0.001, 1)
    for name, g, wv in zip(("vp", "vtmp", "image"), got, want):
        assert_bit_equal(g, wv, name)


def test_4096_fuse_depth_independent_and_finite():
    import torch
    h = w = 4096
    vp, vtmp, _, _ = _state(h, w, 99)
    results = []
    old = pfs.get_fuse_depth()
    try:
        for depth in (1, 0):
            pfs.set_fuse_depth(depth)
            fv, ft = pfs.vp_field(to_dev(vp)), pfs.vp_field(to_dev(vtmp))
            for _ in range(2):
                pfs.simulate_fluid_step(fv, ft, 0.1, 0.001, 100, 100)
            assert bool(torch.isfinite(fv.data).all()) and bool(torch.isfinite(ft.data).all())
            results.append((fv.data.clone(), ft.data.clone()))
    finally:
        pfs.set_fuse_depth(old)
    assert torch.equal(results[0][0].view(torch.int32), results[1][0].view(torch.int32))
    assert torch.equal(results[0][1].view(torch.int32), results[1][1].view(torch.int32))


@pytest.mark.parametrize("shift", [(0, 4), (7, 0), (1237, 2052)])
def test_stencil_phases_commute_with_periodic_shifts(shift):
    """diffuse / computePressure / subtractPressureGradient are pure periodic stencils: rolling the
    inputs by (dy, dx) cells rolls the outputs by exactly the same amount, bit for bit."""
    import torch
    h, w = 2048, 4096
    rng = np.random.default_rng(5)
    a = rng.standard_normal((h, w, 4)).astype(np.float32)
    b = rng.standard_normal((h, w, 4)).astype(np.float32)

    def run(x, y):
        fa, fb = pfs.vp_field(to_dev(x)), pfs.vp_field(to_dev(y))
        pfs.diffuse(fa, fb, 0.01, 1.0, 20)
        pfs.computePressure(fb, fa, 0.5, 37)
        pfs.subtractPressureGradient(fa, fb, 0.5)
        return fa.data, fb.data

    r0 = run(a, b)
    r1 = run(np.roll(a, shift, axis=(0, 1)), np.roll(b, shift, axis=(0, 1)))
    for x, y in zip(r0, r1):
        assert torch.equal(torch.roll(x, shifts=shift, dims=(0, 1)).view(torch.int32), y.view(torch.int32))


def test_constant_pressure_zero_divergence_is_a_fixed_point():
    h, w = 1024, 4096
    a = np.zeros((h, w, 4), np.float32)
    a[..., 0] = 0.25       # uniform velocity: divergence exactly 0
    a[..., 1] = -0.5
    a[..., 2] = 3.0        # constant pressure: (4*3 + 0)/4 == 3 exactly
    b = np.ones((h, w, 4), np.float32)
    fa, fb = pfs.vp_field(to_dev(a)), pfs.vp_field(to_dev(b))
    pfs.computePressure(fa, fb, 0.1, 100)
    ga, gb = to_host(fa.data), to_host(fb.data)
    assert (ga[..., 2] == 3.0).all() and (gb[..., 2] == 3.0).all()
    assert (ga[..., 3] == 0.0).all() and (gb[..., 3] == 0.0).all()


def test_16384_wide_slab_runs_and_matches_oracle_rows():
    """The widest grid of BASELINE.json (16384 columns, int32 index limit of the reference) on a
    64-row slab -- small enough for the oracle."""
    h, w = 64, 16384
    vp, vtmp, image, itmp = _state(h, w, 7)
    fv, ft = pfs.vp_field(to_dev(vp)), pfs.vp_field(to_dev(vtmp))
    pfs.simulate_fluid_step(fv, ft, 0.5, 0.001, 20, 20)
    want_vp, want_vt = oracle.Oracle(20, 20).simulate_fluid_step(vp, vtmp, 0.5, 0.001)
    assert_bit_equal(to_host(fv.data), want_vp, "vp")
    assert_bit_equal(to_host(ft.data), want_vt, "vtmp")
