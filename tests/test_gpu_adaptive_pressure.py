"""pfs_compute_pressure_adaptive: the sweep count is chosen at run time, and the result is bit for bit
what the fixed-count computePressure of the reference (fluid.cpp:210-267, through the oracle) leaves for that count."""
import numpy as np
import pytest

import oracle
import probabilistic_fluid_simulation_b200 as pfs
from golden_util import assert_bit_equal
from gpu_util import to_dev, to_host

pytestmark = pytest.mark.gpu


def smooth_field(h, w, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    f = np.zeros((h, w, 4), np.float32)
    f[..., 0] = np.sin(2 * np.pi * x / w) * np.cos(2 * np.pi * y / h) + 0.05 * rng.standard_normal((h, w))
    f[..., 1] = np.cos(4 * np.pi * x / w) * np.sin(2 * np.pi * y / h) + 0.05 * rng.standard_normal((h, w))
    f[..., 2] = 0.1 * rng.standard_normal((h, w))
    f[..., 3] = rng.standard_normal((h, w))
    return f


def oracle_rms(a, b, dt, n):
    ra, rb = oracle.Oracle().compute_pressure(a.copy(), b.copy(), dt, n)
    d = rb[..., 2].astype(np.float64) - ra[..., 2].astype(np.float64)   # p_N - p_{N-1}
    return ra, rb, float(np.sqrt(np.mean(d * d)))


@pytest.mark.parametrize("shape,check_every,max_sweeps", [((64, 96), 8, 200), ((130, 260), 5, 120), ((48, 40), 16, 400),
                                                          ((29, 37), 3, 64)])
def test_adaptive_stops_at_the_first_batch_under_tol(shape, check_every, max_sweeps):
    h, w = shape
    a, b = smooth_field(h, w, 21), smooth_field(h, w, 22)
    dt = 0.37
    # choose the tolerance from the oracle so that the solve must stop strictly inside (0, max_sweeps)
    target = 4 * check_every
    _, _, rms_t = oracle_rms(a, b, dt, target)
    _, _, rms_prev = oracle_rms(a, b, dt, target - check_every)
    assert rms_t < rms_prev
    tol = 0.5 * (rms_t + rms_prev)
    fa, fb = pfs.vp_field(to_dev(a)), pfs.vp_field(to_dev(b))
    n, rms = pfs.computePressureAdaptive(fa, fb, dt, tol, max_sweeps, check_every)
    assert n % check_every == 0 and 0 < n <= target
    ra, rb, want_rms = oracle_rms(a, b, dt, n)
    assert_bit_equal(to_host(fa.data), ra, "adaptive vp")
    assert_bit_equal(to_host(fb.data), rb, "adaptive vp_out")
    assert rms <= tol and abs(rms - want_rms) <= 1e-12 * max(1.0, want_rms)
    if n > check_every:     # minimal: the batch before was still above the tolerance
        assert oracle_rms(a, b, dt, n - check_every)[2] > tol
    # and it is exactly the fixed-count operator at that count
    ga, gb = pfs.vp_field(to_dev(a)), pfs.vp_field(to_dev(b))
    pfs.computePressure(ga, gb, dt, n)
    assert_bit_equal(to_host(fa.data), to_host(ga.data), "vs fixed count vp")
    assert_bit_equal(to_host(fb.data), to_host(gb.data), "vs fixed count vp_out")


@pytest.mark.parametrize("max_sweeps,check_every", [(1, 4), (2, 4), (17, 4), (33, 8), (30, 30), (31, 30)])
def test_adaptive_with_zero_tolerance_runs_to_the_cap(max_sweeps, check_every):
    a, b = smooth_field(40, 56, 31), smooth_field(40, 56, 32)
    fa, fb = pfs.vp_field(to_dev(a)), pfs.vp_field(to_dev(b))
    n, rms = pfs.computePressureAdaptive(fa, fb, 0.2, 0.0, max_sweeps, check_every)
    assert n == max_sweeps
    ra, rb, want_rms = oracle_rms(a, b, 0.2, n)
    assert_bit_equal(to_host(fa.data), ra, "vp")
    assert_bit_equal(to_host(fb.data), rb, "vp_out")
    assert abs(rms - want_rms) <= 1e-12 * max(1.0, want_rms)


def test_adaptive_rejects_bad_arguments():
    a = smooth_field(16, 16, 1)
    fa, fb = pfs.vp_field(to_dev(a)), pfs.vp_field(to_dev(a))
    with pytest.raises(pfs.PfsError):
        pfs.computePressureAdaptive(fa, fb, 0.1, 1e-3, 100, 1)
    with pytest.raises(pfs.PfsError):
        pfs.computePressureAdaptive(fa, fb, 0.1, -1.0, 100, 8)
    with pytest.raises(pfs.PfsError):
        pfs.computePressureAdaptive(fa, fb, 0.1, float("nan"), 100, 8)


@pytest.mark.parametrize("nranks", [1, 2, 3])
@pytest.mark.parametrize("check_every,max_sweeps", [(8, 200), (5, 64)])
def test_adaptive_on_a_ring_of_slabs(nranks, check_every, max_sweeps):
    """pfs_slab_compute_pressure_adaptive: the rms is folded over the whole ring after every batch, so the ring stops at the
    count the single-GPU solve stops at, and the bands hold the reference's computePressure at that count, bit for bit."""
    from probabilistic_fluid_simulation_b200.slab import SlabRing
    h, w = 96, 128
    a, b = smooth_field(h, w, 41), smooth_field(h, w, 42)
    dt = 0.37
    target = 3 * check_every
    tol = 0.5 * (oracle_rms(a, b, dt, target)[2] + oracle_rms(a, b, dt, target - check_every)[2])
    fa, fb = pfs.vp_field(to_dev(a)), pfs.vp_field(to_dev(b))
    n1, rms1 = pfs.computePressureAdaptive(fa, fb, dt, tol, max_sweeps, check_every)
    ring = SlabRing(nranks, w, h)
    ba, bb = ring.split(a), ring.split(b)
    n, rms = ring.compute_pressure_adaptive(ba, bb, dt, tol, max_sweeps, check_every)
    ring.check()
    assert n == n1 and abs(rms - rms1) <= 1e-12 * max(1.0, rms1)
    ra, rb, _ = oracle_rms(a, b, dt, n)
    assert_bit_equal(ring.gather(ba), ra, f"ring vp (R={nranks})")
    assert_bit_equal(ring.gather(bb), rb, f"ring vp_out (R={nranks})")
    ring.close()


def test_sor_reaches_the_tolerance_in_fewer_sweeps_than_jacobi():
    """Red-black SOR is NOT a parity path: it is checked for what it claims -- the same discrete equation solved to the same
    update tolerance in fewer sweeps.  Its result must satisfy the reference's own Jacobi fixed-point equation as well as
    the Jacobi iterate at that tolerance does."""
    from probabilistic_fluid_simulation_b200 import fixtures
    h, w = 256, 256
    a, b, _, _ = fixtures.make_state(fixtures.smooth_velocity_bytes(h, w), fixtures.random_image_bytes(8, 8, 1))
    dt, tol = 1.0, 1e-3
    fa, fb = pfs.vp_field(to_dev(a)), pfs.vp_field(to_dev(b))
    n_jac, rms_jac = pfs.computePressureAdaptive(fa, fb, dt, tol, 20000, 8)
    p_jac = to_host(fb.data if n_jac % 2 else fa.data)[..., 2].astype(np.float64)
    ga, gb = pfs.vp_field(to_dev(a)), pfs.vp_field(to_dev(b))
    n_sor, rms_sor = pfs.computePressureSOR(ga, gb, dt, 1.9, tol, 20000, 4)
    out = to_host(gb.data)
    p_sor, div = out[..., 2].astype(np.float64), out[..., 3].astype(np.float64)
    assert rms_jac <= tol and rms_sor <= tol
    assert n_sor * 3 < n_jac, (n_sor, n_jac)

    def fixed_point_residual(p):
        s = np.roll(p, 1, 1) + np.roll(p, -1, 1) + np.roll(p, 1, 0) + np.roll(p, -1, 0) + div
        return float(np.sqrt(np.mean((0.25 * s - p) ** 2)))
    assert fixed_point_residual(p_sor) <= 2.0 * max(fixed_point_residual(p_jac), tol)
    # divergence channel as the reference forms it, in both buffers
    ra, rb = oracle.Oracle().compute_pressure(a.copy(), b.copy(), dt, 1)
    assert np.array_equal(out[..., 3].view(np.uint32), rb[..., 3].view(np.uint32))
    assert np.array_equal(to_host(ga.data)[..., 3].view(np.uint32), ra[..., 3].view(np.uint32))
