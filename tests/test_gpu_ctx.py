"""GPU parity, part 5: persistent-state contexts (pfs_ctx_*).  The state lives in the library's planar
layout between steps; what n context steps leave behind must equal, bit for bit, what the CPU oracle's n timesteps leave
in its two buffers -- for every sweep-count parity (the reference's pointer choreography decides which iterates survive),
with and without step graphs, after partial re-uploads, and with the forcing variants."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
import probabilistic_fluid_simulation_b200 as pfs
from golden_util import assert_bit_equal, case_state, load_golden
from gpu_util import to_dev, to_host
from probabilistic_fluid_simulation_b200 import fixtures

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rand_field(h, w, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((h, w, 4)) * scale).astype(np.float32)


def _ctx_run(state, dt, visc, nd, npr, steps, split=False):
    vp, vtmp, image, itmp = state
    h, w = vp.shape[:2]
    ih, iw = image.shape[:2] if image is not None else (0, 0)
    ctx = pfs.FluidContext(w, h, iw, ih)
    ctx.upload(to_dev(vp), to_dev(vtmp), None if image is None else to_dev(image))
    if split:
        for _ in range(steps):
            ctx.simulate_fluid_step(dt, visc, nd, npr)
            if image is not None:
                ctx.advect_color_step(dt)
    else:
        ctx.step(steps, dt, visc, nd, npr)
    out = ctx.download()
    ctx.close()
    return [None if t is None else to_host(t) for t in out]


@pytest.mark.parametrize("nd,npr", [(1, 1), (1, 2), (2, 1), (2, 2), (3, 3), (3, 4), (4, 3), (5, 8), (30, 30), (7, 30),
                                     (30, 7), (100, 100)])
def test_ctx_steps_with_any_sweep_counts(nd, npr):
    h, w = 36, 52
    vp, vt = rand_field(h, w, 21, 0.8), rand_field(h, w, 22, 0.5)
    img = np.random.default_rng(5).random((48, 40, 4)).astype(np.float32)
    state = (vp, vt, img, np.zeros_like(img))
    got = _ctx_run([x.copy() for x in state], 0.4, 0.02, nd, npr, 5)
    want = oracle.Oracle(nd, npr).run_steps(*[x.copy() for x in state], 0.4, 0.02, 5)
    for name, g, wv in zip(("vp", "vtmp", "image"), got, want):
        assert_bit_equal(g, wv, f"{name} ({nd}+{npr})")


@pytest.mark.parametrize("shape", [(16, 16), (29, 37), (64, 4), (130, 260), (1, 8), (8, 1), (257, 512)])
def test_ctx_shapes(shape):
    h, w = shape
    vp, vt = rand_field(h, w, 31, 0.8), rand_field(h, w, 32, 0.5)
    img = np.random.default_rng(6).random((h + 3, 2 * w + 1, 4)).astype(np.float32)
    state = (vp, vt, img, np.zeros_like(img))
    got = _ctx_run([x.copy() for x in state], 2.5, 0.003, 7, 10, 3, split=True)
    want = oracle.Oracle(7, 10).run_steps(*[x.copy() for x in state], 2.5, 0.003, 3)
    for name, g, wv in zip(("vp", "vtmp", "image"), got, want):
        assert_bit_equal(g, wv, f"{name} {shape}")


def test_ctx_upload_download_round_trip_and_partial_upload():
    h, w = 24, 40
    vp, vt = rand_field(h, w, 41), rand_field(h, w, 42)
    img = np.random.default_rng(7).random((h, w, 4)).astype(np.float32)
    ctx = pfs.FluidContext(w, h, w, h)
    ctx.upload(to_dev(vp), to_dev(vt), to_dev(img))
    g = [to_host(t) for t in ctx.download()]
    assert_bit_equal(g[0], vp, "vp round trip"); assert_bit_equal(g[1], vt, "vtmp round trip"); assert_bit_equal(g[2], img, "image")
    orc = oracle.Oracle(4, 6)
    ctx.step(2, 0.3, 0.01, 4, 6)
    w_vp, w_vt, w_img, w_itmp = orc.run_steps(vp.copy(), vt.copy(), img.copy(), np.zeros_like(img), 0.3, 0.01, 2)
    # the caller replaces vp only (what addForces-style edits between two steps amount to); vtmp's state must survive
    new_vp = rand_field(h, w, 43, 0.6)
    ctx.upload(to_dev(new_vp), None, None)
    g = [to_host(t) for t in ctx.download()]
    assert_bit_equal(g[0], new_vp, "replaced vp"); assert_bit_equal(g[1], w_vt, "vtmp kept")
    ctx.step(3, 0.3, 0.01, 4, 6)
    w_vp, w_vt, w_img, w_itmp = orc.run_steps(new_vp.copy(), w_vt.copy(), w_img.copy(), w_itmp.copy(), 0.3, 0.01, 3)
    g = [to_host(t) for t in ctx.download()]
    for name, a, b in zip(("vp", "vtmp", "image"), g, (w_vp, w_vt, w_img)):
        assert_bit_equal(a, b, name)
    ctx.close()


@pytest.mark.parametrize("name", ["tulips_voronoi_5", "cfg1_baboon_perlin256_dt10_nu0.001", "odd_shape_n30", "huge_dt_wrap", "n100_256"])
def test_ctx_on_golden_cases(name):
    """The committed hashes were generated from the compiled reference (scripts/make_golden.py)."""
    case = next(c for c in load_golden()["cases"] if c["name"] == name)
    state = case_state(case)
    n = int(case["n_iters"])
    got = _ctx_run(state, np.float32(case["dt"]), np.float32(case["viscosity"]), n, n, int(case["steps"]))
    for fname, arr in zip(("vp", "vtmp", "image"), got):
        if fname in case["hashes"] and arr is not None:
            assert oracle.field_hashes(arr) == case["hashes"][fname], f"{name}: {fname}"


def test_ctx_forced_and_stochastic_steps():
    h, w = 40, 256
    vp, vt, f = rand_field(h, w, 51, 0.8), rand_field(h, w, 52, 0.5), rand_field(h, w, 53, 0.2)
    orc = oracle.Oracle(13, 8)
    ctx = pfs.FluidContext(w, h)
    ctx.upload(to_dev(vp), to_dev(vt))
    df = to_dev(f)
    for k in range(3):
        ctx.simulate_fluid_step(0.4, 0.02, 13, 8, forces=df)
        vp, vt = orc.simulate_fluid_step_forced(vp, vt, 0.4, 0.02, f)
    for k in range(2):
        ctx.simulate_fluid_step(0.4, 0.02, 13, 8, sigma=0.01, seed=99, step=k)
        vp, vt = orc.simulate_fluid_step_stochastic(vp, vt, 0.4, 0.02, 0.01, 99, k)
    g = ctx.download()
    assert_bit_equal(to_host(g[0]), vp, "vp"); assert_bit_equal(to_host(g[1]), vt, "vtmp")
    ctx.close()


@pytest.mark.parametrize("env_extra", [{"PFS_STEP_GRAPH": "0"}, {"PFS_STEP_GRAPH": "0", "PFS_PDL": "0"}, {"PFS_PDL": "0"}],
                         ids=["eager+pdl", "eager", "graphs-no-pdl"])
def test_ctx_without_step_graphs_in_a_child_process(env_extra):
    """PFS_STEP_GRAPH / PFS_PDL are read once per process: eager launches (with and without programmatic dependent launch of
    the fused passes) must give the same bits as graph replays."""
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import oracle, probabilistic_fluid_simulation_b200 as pfs
from gpu_util import to_dev, to_host
rng = np.random.default_rng(3)
vp = (rng.standard_normal((64, 128, 4)) * 0.7).astype(np.float32); vt = (rng.standard_normal((64, 128, 4)) * 0.4).astype(np.float32)
img = rng.random((64, 128, 4)).astype(np.float32)
ctx = pfs.FluidContext(128, 64, 128, 64)
ctx.upload(to_dev(vp), to_dev(vt), to_dev(img))
ctx.step(6, 0.5, 0.004, 30, 30)
want = oracle.Oracle(30, 30).run_steps(vp, vt, img, np.zeros_like(img), 0.5, 0.004, 6)
for g, w in zip(ctx.download(), want):
    assert np.array_equal(to_host(g).view(np.uint32), w.view(np.uint32))
print("eager ctx ok")
''' % (ROOT, os.path.join(ROOT, "tests"))
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "eager ctx ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_ctx_full_size_matches_stateless_path():
    """4096^2, 100+100 sweeps, two steps: the context path against the stateless entry points (which
    tests/test_gpu_fullsize.py pins against the oracle at this size)."""
    import torch
    n = 4096
    vel = fixtures.smooth_velocity_bytes(n, n)
    noise = fixtures.hash_bytes(n, n, 2, 1234).astype(np.int16) % 13 - 6
    vel[..., :2] = np.clip(vel[..., :2].astype(np.int16) + noise, 0, 255).astype(np.uint8)
    img = fixtures.hash_bytes(n, n, 4, 4321)
    vp, vtmp, image, itmp = fixtures.make_state(vel, img)
    fv, ft, fi, fm = (pfs.vp_field(to_dev(x)) for x in (vp, vtmp, image, itmp))
    ctx = pfs.FluidContext(n, n, n, n)
    ctx.upload(fv.data, ft.data, fi.data)
    for _ in range(2):
        pfs.simulate_fluid_step(fv, ft, 0.1, 0.001, 100, 100)
        pfs.advect_color_step(fi, fm, fv, 0.1)
    ctx.step(2, 0.1, 0.001, 100, 100)
    g = ctx.download()
    for name, a, b in zip(("vp", "vtmp", "image"), g, (fv.data, ft.data, fi.data)):
        assert torch.equal(a.view(torch.int32), b.view(torch.int32)), name
    ctx.close()


@pytest.mark.parametrize("ratio", ["same", "odd"])
def test_frame_bytes_from_the_advecting_kernel(ratio):
    """pfs_ctx_advect_color_step_rgba8 / pfs_advect_color_step_rgba8: the kernel that advects the image also stores the frame;
    the bytes must be the reference writer's (png_byte)(x*255.0) (utils.hpp:129-131, through oracle.unit_float_to_bytes) of the
    oracle's new image -- including values one ulp either side of every k/255 -- and the image itself is unchanged by it."""
    import torch
    h, w = 40, 64
    ih, iw = (h, w) if ratio == "same" else (57, 131)
    vp, vt = rand_field(h, w, 71, 0.8), rand_field(h, w, 72, 0.5)
    vp[..., :2] = 0.0                       # zero velocity: the advected image is the input image, boundary values survive
    rng = np.random.default_rng(73)
    img = rng.random((ih, iw, 4)).astype(np.float32)
    k255 = fixtures.bytes_to_unit_float(np.arange(256, dtype=np.uint8))
    flat = img.reshape(-1)
    flat[:256] = k255
    flat[256:512] = np.nextafter(k255, np.float32(0))
    flat[512:768] = np.nextafter(k255, np.float32(2))
    flat[768:772] = [-0.25, 1.5, -0.0, 0.9999998]
    _, want_img = oracle.Oracle().advect_color(img.copy(), np.zeros_like(img), vp, 0.3)     # (image, itmp): the result is in itmp
    want_bytes = oracle.unit_float_to_bytes(np.clip(want_img, 0.0, np.float32(255.99 / 255.0)))
    in_range = (want_img >= 0) & (want_img < np.float32(256.0 / 255.0))
    # context
    ctx = pfs.FluidContext(w, h, iw, ih)
    ctx.upload(to_dev(vp), to_dev(vt), to_dev(img))
    frame = torch.zeros((ih, iw, 4), dtype=torch.uint8, device="cuda")
    ctx.advect_color_step(0.3, frame_out=frame)
    got_img = to_host(ctx.download(image_only=True)[2])
    ctx.close()
    assert_bit_equal(got_img, want_img, "ctx image")
    got = frame.cpu().numpy()
    assert np.array_equal(got[in_range], want_bytes[in_range])
    assert np.array_equal(got, pfs.image_to_rgba8(pfs.vp_field(to_dev(want_img))).cpu().numpy())     # saturation included
    # stateless entry point
    fi, fm, fv = pfs.vp_field(to_dev(img)), pfs.vp_field(to_dev(np.zeros_like(img))), pfs.vp_field(to_dev(vp))
    frame2 = torch.zeros_like(frame)
    pfs.advect_color_step(fi, fm, fv, 0.3, frame_out=frame2)
    assert_bit_equal(to_host(fi.data), want_img, "stateless image")
    assert torch.equal(frame2, frame)


def test_frame_bytes_after_real_steps_equal_the_separate_pack():
    h, w = 96, 160
    vel = fixtures.smooth_velocity_bytes(h, w)
    vp, vtmp, image, itmp = fixtures.make_state(vel, fixtures.random_image_bytes(2 * h, 2 * w, 3))
    import torch
    ctx = pfs.FluidContext(w, h, 2 * w, 2 * h)
    ctx.upload(to_dev(vp), to_dev(vtmp), to_dev(image))
    frame = torch.zeros((2 * h, 2 * w, 4), dtype=torch.uint8, device="cuda")
    for _ in range(4):
        ctx.simulate_fluid_step(50.0, 0.001, 9, 12)
        ctx.advect_color_step(50.0, frame_out=frame)
    got_img = ctx.download(image_only=True)[2]
    ctx.close()
    want = oracle.Oracle(9, 12).run_steps(vp, vtmp, image, itmp, 50.0, 0.001, 4)
    assert_bit_equal(to_host(got_img), want[2], "image after 4 steps")
    assert np.array_equal(frame.cpu().numpy(), oracle.unit_float_to_bytes(want[2]))
