"""GPU parity, part 4: the row-slab (multi-GPU) path.  R slabs of one periodic grid, driven inside one
process on ONE GPU (direct-copy transport), must reproduce the CPU oracle bit for bit -- the same check
as the single-GPU path, so it also proves slab result == single-GPU result.  The NCCL transport is
exercised by test_ring_of_processes (2, 4 and 8 processes, one GPU each) when that many GPUs are visible."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
from golden_util import assert_bit_equal
from probabilistic_fluid_simulation_b200 import fixtures
from probabilistic_fluid_simulation_b200.slab import SlabRing

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _state(h, w, ih, iw, seed):
    vel = fixtures.smooth_velocity_bytes(h, w)
    rng = np.random.default_rng(seed)
    vel[..., :2] = np.clip(vel[..., :2].astype(np.int16) + rng.integers(-9, 10, size=(h, w, 2)), 0, 255).astype(np.uint8)
    img = fixtures.random_image_bytes(ih, iw, seed + 1)
    return fixtures.make_state(vel, img)


def _run_ring(nranks, state, dt, visc, nd, npr, steps):
    vp, vtmp, image, itmp = state
    h, w = vp.shape[:2]
    ih, iw = image.shape[:2]
    ring = SlabRing(nranks, w, h, iw, ih)
    bv, bt = ring.split(vp), ring.split(vtmp)
    bi, bm = ring.split(image, image=True), ring.split(itmp, image=True)
    for _ in range(steps):
        ring.simulate_fluid_step(bv, bt, dt, visc, nd, npr)
        ring.advect_color_step(bi, bm, bv, dt)
    ring.check()
    out = ring.gather(bv), ring.gather(bt), ring.gather(bi)
    ring.close()
    return out


@pytest.mark.parametrize("nranks", [1, 2, 3, 4])
@pytest.mark.parametrize("dt,visc,nd,npr", [(0.5, 0.003, 30, 30), (40.0, 0.01, 7, 10), (3.0, 0.0, 3, 4)])
def test_ring_matches_oracle(nranks, dt, visc, nd, npr):
    h, w, ih, iw = 96, 128, 96, 128
    state = _state(h, w, ih, iw, 3)
    got = _run_ring(nranks, [x.copy() for x in state], dt, visc, nd, npr, 3)
    want = oracle.Oracle(nd, npr).run_steps(*state, dt, visc, 3)
    for name, g, wv in zip(("vp", "vtmp", "image"), got, want):
        assert_bit_equal(g, wv, f"{name} (R={nranks})")


@pytest.mark.parametrize("nranks", [2, 4])
def test_ring_large_displacement_uses_all_gather(nranks):
    """dt so large that departure points leave the neighbouring slabs: the gather source becomes the whole
    field (all-gather path)."""
    h, w, ih, iw = 64, 96, 64, 96
    state = _state(h, w, ih, iw, 5)
    dt = 3000.0
    got = _run_ring(nranks, [x.copy() for x in state], dt, 0.001, 4, 4, 2)
    want = oracle.Oracle(4, 4).run_steps(*state, dt, 0.001, 2)
    for name, g, wv in zip(("vp", "vtmp", "image"), got, want):
        assert_bit_equal(g, wv, f"{name} (R={nranks})")


@pytest.mark.parametrize("nranks", [2, 3])
def test_ring_image_ratio_not_integer(nranks):
    """768x512-style image on a 256-wide grid: viw = 1/3 is inexact in binary32;
    the image bands follow the float look-up of fluid.cpp:89-90."""
    h, w, ih, iw = 72, 64, 96, 192
    state = _state(h, w, ih, iw, 7)
    got = _run_ring(nranks, [x.copy() for x in state], 10.0, 0.001, 30, 30, 2)
    want = oracle.Oracle(30, 30).run_steps(*state, 10.0, 0.001, 2)
    for name, g, wv in zip(("vp", "vtmp", "image"), got, want):
        assert_bit_equal(g, wv, f"{name} (R={nranks})")


def test_ring_wide_grid_fused_depth():
    """A 2048-wide grid over 4 slabs with 100+100 sweeps (the fused passes and their halo depth 8)."""
    h, w = 128, 2048
    state = _state(h, w, h, w, 9)
    got = _run_ring(4, [x.copy() for x in state], 0.1, 0.001, 100, 100, 1)
    want = oracle.Oracle(100, 100).run_steps(*state, 0.1, 0.001, 1)
    for name, g, wv in zip(("vp", "vtmp", "image"), got, want):
        assert_bit_equal(g, wv, name)


@pytest.mark.parametrize("nranks", [2, 4])
def test_ring_speculative_gather_depth_reruns_when_the_field_jumps(nranks):
    """From the second step on, the gather depth of the velocity advection is a guess from the previous step's max|v|
    and a device flag says whether it was enough.  Here the caller replaces the velocities between two steps by ones
    100x larger (the API allows that: it is what addForces is for), so the guess is too shallow: the step must notice,
    rerun itself with the exact bound and still match the oracle bit for bit."""
    from probabilistic_fluid_simulation_b200 import _cabi
    h, w = 256, 64
    vp, vtmp, image, itmp = _state(h, w, h, w, 13)
    big = vp[..., :2].copy()
    vp[..., :2] *= np.float32(0.01)
    dt, visc, nd, npr = 4000.0, 0.001, 4, 4
    orc = oracle.Oracle(nd, npr)
    ring = SlabRing(nranks, w, h, w, h)
    bv, bt = ring.split(vp), ring.split(vtmp)
    L = _cabi.lib()
    c0 = L.pfs_kernel_launch_count()
    ring.simulate_fluid_step(bv, bt, dt, visc, nd, npr)           # measured bound (first step)
    c1 = L.pfs_kernel_launch_count()
    ring.simulate_fluid_step(bv, bt, dt, visc, nd, npr)           # guessed bound, sufficient
    c2 = L.pfs_kernel_launch_count()
    want_v, want_t = orc.simulate_fluid_step(vp, vtmp, dt, visc)
    want_v, want_t = orc.simulate_fluid_step(want_v, want_t, dt, visc)
    assert_bit_equal(ring.gather(bv), want_v, "vp after two steps")
    hv = ring.gather(bv)
    hv[..., :2] = big                                               # the field jumps
    want_v = want_v.copy()
    want_v[..., :2] = big
    bv = ring.split(hv)
    ring.simulate_fluid_step(bv, bt, dt, visc, nd, npr)           # guessed bound too shallow -> rerun
    c3 = L.pfs_kernel_launch_count()
    want_v, want_t = orc.simulate_fluid_step(want_v, want_t, dt, visc)
    assert_bit_equal(ring.gather(bv), want_v, "vp after the jump")
    assert_bit_equal(ring.gather(bt), want_t, "vtmp after the jump")
    ring.check()
    ring.close()
    assert (c3 - c2) > 1.5 * (c2 - c1), (c1 - c0, c2 - c1, c3 - c2)   # the third step really ran twice


def _run_ring_resident(nranks, state, dt, visc, nd, npr, steps, split_steps=False):
    """The same run through pfs_slab_upload / pfs_slab_step / pfs_slab_download (state resident in the slabs' planes)."""
    import torch
    vp, vtmp, image, itmp = state
    h, w = vp.shape[:2]
    ih, iw = image.shape[:2]
    ring = SlabRing(nranks, w, h, iw, ih)
    bv, bt, bi = ring.split(vp), ring.split(vtmp), ring.split(image, image=True)
    ring.upload(bv, bt, bi)
    if split_steps:
        for _ in range(steps):
            ring.step(1, dt, visc, nd, npr)
    else:
        ring.step(steps, dt, visc, nd, npr)
    ov, ot, oi = [torch.empty_like(t) for t in bv], [torch.empty_like(t) for t in bt], [torch.empty_like(t) for t in bi]
    ring.download(ov, ot, oi)
    ring.check()
    out = ring.gather(ov), ring.gather(ot), ring.gather(oi)
    ring.close()
    return out


@pytest.mark.parametrize("nranks", [1, 2, 3, 4])
@pytest.mark.parametrize("dt,visc,nd,npr", [(0.5, 0.003, 30, 30), (40.0, 0.01, 7, 10), (3.0, 0.0, 3, 4), (2.0, 0.002, 1, 1), (2.0, 0.002, 2, 5)])
def test_resident_ring_matches_oracle(nranks, dt, visc, nd, npr):
    h, w, ih, iw = 96, 128, 96, 128
    state = _state(h, w, ih, iw, 3)
    got = _run_ring_resident(nranks, [x.copy() for x in state], dt, visc, nd, npr, 4)
    want = oracle.Oracle(nd, npr).run_steps(*state, dt, visc, 4)
    for name, g, wv in zip(("vp", "vtmp", "image"), got, want):
        assert_bit_equal(g, wv, f"{name} (R={nranks})")


@pytest.mark.parametrize("nranks", [2, 3])
def test_resident_ring_image_ratio_and_large_displacement(nranks):
    h, w, ih, iw = 72, 64, 96, 192
    state = _state(h, w, ih, iw, 7)
    for dt in (10.0, 3000.0):
        got = _run_ring_resident(nranks, [x.copy() for x in state], dt, 0.001, 6, 8, 3, split_steps=True)
        want = oracle.Oracle(6, 8).run_steps(*[x.copy() for x in state], dt, 0.001, 3)
        for name, g, wv in zip(("vp", "vtmp", "image"), got, want):
            assert_bit_equal(g, wv, f"{name} (R={nranks}, dt={dt})")


def test_resident_ring_repairs_a_wrong_gather_guess():
    """Resident state: both gather depths of step k are guessed from the field step k-1 started from.  Here the field is
    replaced (a new upload keeps the old bound out of the picture) ... and, harder, the time step grows 100x between two
    calls on the SAME state, so the guesses of the second call are far too shallow for the velocity advection and for the
    image advection: both must be noticed and repaired, bit for bit."""
    import torch
    h, w = 256, 64
    state = _state(h, w, h, w, 13)
    vp, vtmp, image, itmp = [x.copy() for x in state]
    ring = SlabRing(4, w, h, w, h)
    bv, bt, bi = ring.split(vp), ring.split(vtmp), ring.split(image, image=True)
    ring.upload(bv, bt, bi)
    nd, npr, visc = 4, 4, 0.001
    ring.step(2, 40.0, visc, nd, npr)
    ring.step(2, 4000.0, visc, nd, npr)
    ov, ot, oi = [torch.empty_like(t) for t in bv], [torch.empty_like(t) for t in bt], [torch.empty_like(t) for t in bi]
    ring.download(ov, ot, oi)
    ring.check()
    orc = oracle.Oracle(nd, npr)
    want = orc.run_steps(vp, vtmp, image, itmp, 40.0, visc, 2)
    want = orc.run_steps(*want, 4000.0, visc, 2)
    for name, g, wv in zip(("vp", "vtmp", "image"), (ring.gather(ov), ring.gather(ot), ring.gather(oi)), want):
        assert_bit_equal(g, wv, name)
    ring.close()


@pytest.mark.parametrize("nranks", [1, 2, 3])
@pytest.mark.parametrize("nd,npr", [(30, 30), (7, 10), (1, 2), (6, 6)])
def test_ring_forced_step(nranks, nd, npr):
    """The external force of the addForces slot on slabs: every rank gets its band of the force field; the rows a fused
    pass recomputes outside its band get no force, so the halo is exchanged again before the divergence."""
    h, w = 96, 128
    vp, vtmp, _, _ = _state(h, w, 8, 8, 21)
    force = (np.random.default_rng(22).standard_normal((h, w, 4)) * 0.1).astype(np.float32)
    ring = SlabRing(nranks, w, h)
    bv, bt, bf = ring.split(vp), ring.split(vtmp), ring.split(force)
    orc = oracle.Oracle(nd, npr)
    for _ in range(3):
        ring.simulate_fluid_step(bv, bt, 0.5, 0.003, nd, npr, forces=bf)
        vp, vtmp = orc.simulate_fluid_step_forced(vp, vtmp, 0.5, 0.003, force)
    ring.check()
    assert_bit_equal(ring.gather(bv), vp, f"vp (R={nranks})")
    assert_bit_equal(ring.gather(bt), vtmp, f"vtmp (R={nranks})")
    ring.close()


@pytest.mark.parametrize("transport", ["p2p", "nccl"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_ring_of_processes(tmp_path, world, transport):
    """One process per GPU, `world` of them.  "p2p": halo rows stored into the neighbours' memory through CUDA IPC mappings
    (the default when every rank can map its neighbours); "nccl": send/recv.  Both must reproduce the oracle bit for bit
    (tests/nccl_ring_check.py: bands of different heights, a 25-row gather, a gather deeper than the peer halos, the
    whole-field path, the rerun after a field jump).  Skipped only when fewer than `world` GPUs are visible."""
    _ring_of_processes(world, transport, {})


def test_ring_of_two_processes_without_sweep_graphs():
    """The resident ring's sweep segment launched kernel by kernel (PFS_SLAB_GRAPH=0) and without programmatic dependent
    launch: the same checks as test_ring_of_processes[2-p2p], whose resident case replays captured graphs."""
    _ring_of_processes(2, "p2p", {"PFS_SLAB_GRAPH": "0", "PFS_PDL": "0"}, port_offset=40)


def _ring_of_processes(world, transport, env_extra, port_offset=0):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, {torch.cuda.device_count()} visible")
    env = dict(os.environ, **env_extra)
    env.pop("PFS_SLAB_TRANSPORT", None)
    if transport == "nccl":
        env["PFS_SLAB_TRANSPORT"] = "nccl"
    port = 29541 + 2 * world + (0 if transport == "p2p" else 1) + port_offset
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "nccl_ring_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ring matches oracle") == 4, r.stdout[-3000:]
    assert "jump case matches oracle" in r.stdout, r.stdout[-3000:]
    assert "adaptive case matches oracle" in r.stdout, r.stdout[-3000:]
    assert r.stdout.count(f"transport {transport}") == 4, r.stdout[-3000:]
