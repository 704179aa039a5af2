"""GPU parity, part 2: every operator of includes/fluid.hpp, CUDA (through the C-ABI) vs the CPU
oracle on the same seeded inputs -- values bit for bit, untouched channels untouched, and the
data-pointer exchanges of the reference reproduced."""
import numpy as np
import pytest

import oracle
import probabilistic_fluid_simulation_b200 as pfs
from golden_util import assert_bit_equal
from gpu_util import to_dev, to_host

pytestmark = pytest.mark.gpu

SHAPES = [(16, 16), (29, 37), (48, 40), (8, 128), (64, 4), (130, 260), (1, 8), (8, 1), (257, 512)]


def rand_field(h, w, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((h, w, 4)) * scale).astype(np.float32)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("dt", [0.1, 7.5, 1000.0])
def test_advect(shape, dt):
    h, w = shape
    a, b = rand_field(h, w, 1), rand_field(h, w, 2)
    fa, fb = pfs.vp_field(to_dev(a)), pfs.vp_field(to_dev(b))
    pfs.advect(fa, fb, dt)
    oracle.Oracle().advect(a, b, dt)
    assert_bit_equal(to_host(fb.data), b, "advect out (ch2,3 must be untouched)")
    assert_bit_equal(to_host(fa.data), a, "advect in")


@pytest.mark.parametrize("n", [1, 2, 3, 4, 7, 30, 31])
@pytest.mark.parametrize("shape", [(16, 16), (29, 37), (48, 40), (130, 260), (64, 4), (3, 5)])
def test_diffuse(n, shape):
    h, w = shape
    a, b = rand_field(h, w, 3), rand_field(h, w, 4)
    fa, fb = pfs.vp_field(to_dev(a)), pfs.vp_field(to_dev(b))
    pa, pb = fa.data.data_ptr(), fb.data.data_ptr()
    pfs.diffuse(fa, fb, 0.013, 2.5, n)
    ra, rb = oracle.Oracle().diffuse(a, b, 0.013, 2.5, n)
    assert_bit_equal(to_host(fa.data), ra, "diffuse vp")
    assert_bit_equal(to_host(fb.data), rb, "diffuse vp_out")
    swapped = ra is not a
    assert (fa.data.data_ptr() == pb) == swapped and (fb.data.data_ptr() == pa) == swapped


@pytest.mark.parametrize("visc,dt", [(0.0, 1.0), (0.5, 4.0), (1e-6, 1e-3), (25.0, 10.0)])
def test_diffuse_parameter_range(visc, dt):
    a, b = rand_field(40, 64, 13, 3.0), rand_field(40, 64, 14)
    fa, fb = pfs.vp_field(to_dev(a)), pfs.vp_field(to_dev(b))
    pfs.diffuse(fa, fb, visc, dt, 6)
    ra, rb = oracle.Oracle().diffuse(a, b, visc, dt, 6)
    assert_bit_equal(to_host(fa.data), ra, "diffuse vp")
    assert_bit_equal(to_host(fb.data), rb, "diffuse vp_out")


@pytest.mark.parametrize("beta_bits", ["3ff4291f", "406b65c7", "40d1684b", "3f800d1b", "3fb33333", "40000000"])
def test_diffuse_with_either_constant_division(beta_bits):
    """Divisors for which the fused passes use the verified two-instruction division (the last three) and divisors
    for which that form would be wrong for one numerator, so the three-instruction one is used (the first three;
    tests/test_exact_division.py checks the decision itself)."""
    from test_exact_division import DIV2_WRONG, _beta_to_params
    from probabilistic_fluid_simulation_b200 import _cabi
    visc, dt = _beta_to_params(beta_bits)
    assert _cabi.lib().pfs_diffuse_division_ops(visc, dt) == (3 if beta_bits in DIV2_WRONG else 2)
    h, w = 72, 256
    a, b = rand_field(h, w, 41), rand_field(h, w, 42)
    fa, fb = pfs.vp_field(to_dev(a)), pfs.vp_field(to_dev(b))
    pfs.diffuse(fa, fb, visc, dt, 25)
    ra, rb = oracle.Oracle().diffuse(a, b, visc, dt, 25)
    assert_bit_equal(to_host(fa.data), ra, "diffuse vp")
    assert_bit_equal(to_host(fb.data), rb, "diffuse vp_out")


@pytest.mark.parametrize("n", [1, 2, 3, 4, 7, 30, 31])
@pytest.mark.parametrize("shape", [(16, 16), (29, 37), (48, 40), (130, 260), (64, 4), (3, 5)])
def test_compute_pressure(n, shape):
    h, w = shape
    a, b = rand_field(h, w, 5), rand_field(h, w, 6)
    fa, fb = pfs.vp_field(to_dev(a)), pfs.vp_field(to_dev(b))
    pfs.computePressure(fa, fb, 0.37, n)
    ra, rb = oracle.Oracle().compute_pressure(a, b, 0.37, n)
    assert_bit_equal(to_host(fa.data), ra, "pressure vp")
    assert_bit_equal(to_host(fb.data), rb, "pressure vp_out")


@pytest.mark.parametrize("shape", SHAPES)
def test_subtract_pressure_gradient(shape):
    h, w = shape
    a, b = rand_field(h, w, 7), rand_field(h, w, 8)
    fa, fb = pfs.vp_field(to_dev(a)), pfs.vp_field(to_dev(b))
    pfs.subtractPressureGradient(fa, fb, 0.9)
    oracle.Oracle().subtract_pressure_gradient(a, b, 0.9)
    assert_bit_equal(to_host(fb.data), b, "subtract out")


@pytest.mark.parametrize("ishape,vshape", [((32, 48), (16, 16)), ((41, 50), (29, 37)), ((64, 96), (32, 32)),
                                            ((16, 16), (64, 64)), ((512, 768), (256, 256)), ((100, 7), (3, 300))])
@pytest.mark.parametrize("dt", [0.1, 250.0])
def test_advect_color(ishape, vshape, dt):
    rng = np.random.default_rng(9)
    img = rng.random((ishape[0], ishape[1], 4)).astype(np.float32)
    out = np.zeros_like(img)
    vp = rand_field(vshape[0], vshape[1], 10)
    fi, fo, fv = pfs.vp_field(to_dev(img)), pfs.vp_field(to_dev(out)), pfs.vp_field(to_dev(vp))
    pfs.advect_color(fi, fo, fv, dt)
    oracle.Oracle().advect_color(img, out, vp, dt)
    assert_bit_equal(to_host(fo.data), out, "advect_color out")


def test_add_forces_without_a_force_field_is_the_reference_noop():
    a = rand_field(16, 16, 15)
    fa = pfs.vp_field(to_dev(a))
    pfs.addForces(fa, None)
    assert_bit_equal(to_host(fa.data), a, "addForces")


@pytest.mark.parametrize("shape", [(16, 16), (29, 37), (130, 260), (1, 8)])
def test_add_forces_with_a_force_field(shape):
    """addForces(vp, forces): the reference body is an empty loop (fluid.cpp:198-208) -- parity unpinned by construction;
    the library's definition (velocity channels += force channels 0,1) against its restatement in the oracle."""
    h, w = shape
    a, f = rand_field(h, w, 15), rand_field(h, w, 16, 0.3)
    fa = pfs.vp_field(to_dev(a))
    pfs.addForces(fa, to_dev(f))
    want = oracle.Oracle().add_forces(a.copy(), f)
    assert_bit_equal(to_host(fa.data), want, "addForces")
    assert_bit_equal(want[..., 2:], a[..., 2:], "pressure / divergence channels untouched")


@pytest.mark.parametrize("nd,npr", [(1, 1), (1, 2), (2, 1), (3, 4), (5, 8), (30, 30), (31, 30), (100, 100)])
@pytest.mark.parametrize("shape", [(36, 52), (40, 256), (29, 37)])
def test_forced_step(nd, npr, shape):
    """simulate_fluid_step with addForces at its slot (fluid.cpp:302): the force is added to diffusion iterate n as the
    last fused pass stores it (or by a separate kernel on the unfused paths); iterate n-1 stays unforced, which the
    projection reads when the sweep-count parities say so."""
    h, w = shape
    vp, vt, f = rand_field(h, w, 21, 0.8), rand_field(h, w, 22, 0.5), rand_field(h, w, 23, 0.2)
    fv, ft = pfs.vp_field(to_dev(vp)), pfs.vp_field(to_dev(vt))
    df = to_dev(f)
    orc = oracle.Oracle(nd, npr)
    for _ in range(3):
        pfs.simulate_fluid_step(fv, ft, 0.4, 0.02, nd, npr, forces=df)
        vp, vt = orc.simulate_fluid_step_forced(vp, vt, 0.4, 0.02, f)
    assert_bit_equal(to_host(fv.data), vp, "vp")
    assert_bit_equal(to_host(ft.data), vt, "vtmp")


@pytest.mark.parametrize("nd,npr", [(1, 1), (1, 2), (2, 1), (2, 2), (3, 3), (3, 4), (4, 3), (5, 8), (30, 30), (7, 30),
                                     (30, 7), (100, 100)])
def test_step_with_any_sweep_counts(nd, npr):
    """Independent diffusion / pressure counts: the pointer choreography decides which diffusion
    iterate the projection uses and whether the caller's buffers end up exchanged."""
    h, w = 36, 52
    vp, vt = rand_field(h, w, 21, 0.8), rand_field(h, w, 22, 0.5)
    fv, ft = pfs.vp_field(to_dev(vp)), pfs.vp_field(to_dev(vt))
    p_vp = fv.data.data_ptr()
    orc = oracle.Oracle(nd, npr)
    for _ in range(3):
        pfs.simulate_fluid_step(fv, ft, 1.7, 0.02, nd, npr)
        vp2, vt2 = orc.simulate_fluid_step(vp, vt, 1.7, 0.02)
        exchanged = vp2 is not vp
        vp, vt = vp2, vt2
        assert (fv.data.data_ptr() != p_vp) == exchanged
        p_vp = fv.data.data_ptr()
        assert_bit_equal(to_host(fv.data), vp, "vp")
        assert_bit_equal(to_host(ft.data), vt, "tmp")


@pytest.mark.parametrize("depth", [1, 2, 3, 4, 5, 8, 0])
def test_fuse_depth_is_invisible(depth):
    """Temporal blocking must not change a single bit, whatever the depth."""
    h, w = 96, 256
    vp, vt = rand_field(h, w, 31, 0.8), rand_field(h, w, 32, 0.5)
    want_vp, want_vt = oracle.Oracle(13, 21).simulate_fluid_step(vp.copy(), vt.copy(), 0.9, 0.01)
    old = pfs.get_fuse_depth()
    try:
        pfs.set_fuse_depth(depth)
        fv, ft = pfs.vp_field(to_dev(vp)), pfs.vp_field(to_dev(vt))
        pfs.simulate_fluid_step(fv, ft, 0.9, 0.01, 13, 21)
        assert_bit_equal(to_host(fv.data), want_vp, f"vp depth={depth}")
        assert_bit_equal(to_host(ft.data), want_vt, f"tmp depth={depth}")
    finally:
        pfs.set_fuse_depth(old)


@pytest.mark.parametrize("kind", ["zeros", "neg_zeros", "tiny", "huge", "mixed_patch", "denormal"])
def test_diffuse_zero_and_tiny_fields(kind):
    """Inputs outside the safe range of the 3-instruction FMA division (exact zeros, -0, numerators
    below 2^-96, values above 2^60) must take the repair path and still match fluid.cpp bit for bit."""
    h, w = 160, 384
    rng = np.random.default_rng(41)
    a = (rng.standard_normal((h, w, 4)) * 0.5).astype(np.float32)
    if kind == "zeros":
        a[..., :2] = 0.0
    elif kind == "neg_zeros":
        a[..., :2] = -0.0
    elif kind == "tiny":
        a[..., :2] *= np.float32(1e-36)
    elif kind == "huge":
        a[..., :2] *= np.float32(1e25)
    elif kind == "mixed_patch":
        a[40:90, 100:300, :2] = 0.0          # a still region inside a moving field
        a[120:130, 10:20, 0] = -0.0
    elif kind == "denormal":
        a[..., :2] *= np.float32(1e-42)
    b = rng.standard_normal((h, w, 4)).astype(np.float32)
    for visc, dt, n in ((0.02, 1.5, 19), (0.0, 1.0, 9)):
        x, y = a.copy(), b.copy()
        fa, fb = pfs.vp_field(to_dev(x)), pfs.vp_field(to_dev(y))
        pfs.diffuse(fa, fb, visc, dt, n)
        ra, rb = oracle.Oracle().diffuse(x, y, visc, dt, n)
        assert_bit_equal(to_host(fa.data), ra, f"diffuse vp ({kind}, visc={visc})")
        assert_bit_equal(to_host(fb.data), rb, f"diffuse vp_out ({kind}, visc={visc})")


def test_diffuse_exact_path_alone_in_a_child_process():
    """PFS_DIFFUSE_FORCE_REPAIR=1 (read once per process) sends every work item of the packed passes through the
    out-of-line exact-division path ONLY.  On ordinary fields the fast path is already right, so a broken exact path hides
    behind it unless it runs alone -- that is how a miscompiled store in that path was found (tests/test_sass_stores.py)."""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import oracle, probabilistic_fluid_simulation_b200 as pfs
from gpu_util import to_dev, to_host
rng = np.random.default_rng(5)
for (h, w), n, depth in (((160, 384), 19, 0), ((72, 256), 6, 6), ((40, 128), 2, 2), ((64, 512), 7, 3), ((33, 260), 9, 4)):
    pfs.set_fuse_depth(depth)
    for scale in (1.0, 1e-36):
        a = (rng.standard_normal((h, w, 4)) * 0.5).astype(np.float32); a[..., :2] *= np.float32(scale)
        b = rng.standard_normal((h, w, 4)).astype(np.float32)
        fa, fb = pfs.vp_field(to_dev(a)), pfs.vp_field(to_dev(b))
        pfs.diffuse(fa, fb, 0.02, 1.5, n)
        ra, rb = oracle.Oracle().diffuse(a, b, 0.02, 1.5, n)
        assert np.array_equal(to_host(fa.data).view(np.uint32), ra.view(np.uint32)), (h, w, n, scale)
        assert np.array_equal(to_host(fb.data).view(np.uint32), rb.view(np.uint32)), (h, w, n, scale)
print("exact path ok")
""" % (root, os.path.join(root, "tests"))
    env = dict(os.environ, PFS_DIFFUSE_FORCE_REPAIR="1")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "exact path ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_diffuse_negative_viscosity_takes_the_exact_path():
    """alpha < 0 breaks the convexity argument of the guard: the library must not use the packed
    fast path, and must still match the reference."""
    a, b = rand_field(64, 128, 51), rand_field(64, 128, 52)
    fa, fb = pfs.vp_field(to_dev(a)), pfs.vp_field(to_dev(b))
    pfs.diffuse(fa, fb, -0.01, 1.0, 7)
    ra, rb = oracle.Oracle().diffuse(a, b, -0.01, 1.0, 7)
    assert_bit_equal(to_host(fa.data), ra, "diffuse vp")
    assert_bit_equal(to_host(fb.data), rb, "diffuse vp_out")


def test_step_is_cuda_graph_capturable():
    """The device-pointer step only enqueues work on the caller's stream (no hidden synchronisation, no
    allocation once the scratch planes exist), so a whole timestep can be captured into a CUDA graph; a
    replay must give exactly what the eager calls give."""
    import torch
    h, w = 128, 256
    vp, vt = rand_field(h, w, 61, 0.8), rand_field(h, w, 62, 0.5)
    img = np.random.default_rng(63).random((h, w, 4)).astype(np.float32)

    def fields():
        return (pfs.vp_field(to_dev(vp)), pfs.vp_field(to_dev(vt)), pfs.vp_field(to_dev(img)),
                pfs.vp_field(to_dev(np.zeros_like(img))))

    ev, et, ei, em = fields()                       # eager (also sizes the library's scratch planes)
    pfs.simulate_fluid_step(ev, et, 0.7, 0.01, 30, 30)
    pfs.advect_color_step(ei, em, ev, 0.7)
    torch.cuda.synchronize()

    gv, gt, gi, gm = fields()                       # captured: nothing runs until replay
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            pfs.simulate_fluid_step(gv, gt, 0.7, 0.01, 30, 30)
            pfs.advect_color_step(gi, gm, gv, 0.7)
    graph.replay()
    torch.cuda.synchronize()
    assert_bit_equal(to_host(gv.data), to_host(ev.data), "graph vp")
    assert_bit_equal(to_host(gt.data), to_host(et.data), "graph vtmp")
    assert_bit_equal(to_host(gi.data), to_host(ei.data), "graph image")


def test_step_graph_cache_survives_resizes_and_shutdown():
    """The library captures the 2nd identical step call into a CUDA graph and replays it afterwards.  Alternate
    between grids (forcing a scratch reallocation in between), reuse buffers, shut the library down in the
    middle -- every step must still equal the oracle's."""
    from probabilistic_fluid_simulation_b200 import _cabi
    shapes = [(64, 96), (160, 256), (64, 96), (32, 512)]
    states = {}
    for rnd in range(3):
        for idx, (h, w) in enumerate(shapes):
            key = (idx, h, w)
            if key not in states:
                vp, vt = rand_field(h, w, 100 + idx, 0.8), rand_field(h, w, 200 + idx, 0.5)
                states[key] = [vp, vt, pfs.vp_field(to_dev(vp)), pfs.vp_field(to_dev(vt))]
            vp, vt, fv, ft = states[key]
            for _ in range(4):                       # eager, capture, replay, replay
                pfs.simulate_fluid_step(fv, ft, 0.9, 0.01, 6, 8)
                vp, vt = oracle.Oracle(6, 8).simulate_fluid_step(vp, vt, 0.9, 0.01)
            states[key][0], states[key][1] = vp, vt
            assert_bit_equal(to_host(fv.data), vp, f"vp round {rnd} shape {h}x{w}")
            assert_bit_equal(to_host(ft.data), vt, f"vtmp round {rnd} shape {h}x{w}")
        if rnd == 1:
            import torch
            torch.cuda.synchronize()
            assert _cabi.lib().pfs_shutdown() == 0   # drops scratch, graphs, streams; the next call rebuilds them


def test_changing_parameters_never_replays_a_stale_graph():
    h, w = 48, 64
    vp, vt = rand_field(h, w, 301, 0.8), rand_field(h, w, 302, 0.5)
    fv, ft = pfs.vp_field(to_dev(vp)), pfs.vp_field(to_dev(vt))
    for dt, visc, nd, npr in [(0.5, 0.01, 4, 4)] * 3 + [(0.7, 0.01, 4, 4)] * 3 + [(0.7, 0.02, 4, 4)] * 3 + [(0.7, 0.02, 6, 4)] * 3 + \
                             [(0.7, 0.02, 6, 2)] * 3 + [(0.5, 0.01, 4, 4)] * 3:
        pfs.simulate_fluid_step(fv, ft, dt, visc, nd, npr)
        vp, vt = oracle.Oracle(nd, npr).simulate_fluid_step(vp, vt, dt, visc)
        assert_bit_equal(to_host(fv.data), vp, f"vp dt={dt} visc={visc} nd={nd} np={npr}")
        assert_bit_equal(to_host(ft.data), vt, "vtmp")


def test_image_to_rgba8_matches_reference_writer():
    """(png_byte)(x * 255.0), utils.hpp:129-131: double multiply + truncation, every k/255 boundary included."""
    import torch
    rng = np.random.default_rng(91)
    img = rng.random((67, 129, 4)).astype(np.float32)
    img.reshape(-1)[:256] = fixtures_bytes_as_floats()
    img.reshape(-1)[256:512] = np.nextafter(fixtures_bytes_as_floats(), np.float32(0))     # one ulp below k/255
    img[0, 0, 3] = np.float32(0.9999998)                                                   # alpha 254
    got = pfs.image_to_rgba8(pfs.vp_field(to_dev(img))).cpu().numpy()
    want = oracle.unit_float_to_bytes(img)
    assert np.array_equal(got, want)
    assert got[0, 0, 3] == 254


def fixtures_bytes_as_floats():
    from probabilistic_fluid_simulation_b200 import fixtures
    return fixtures.bytes_to_unit_float(np.arange(256, dtype=np.uint8))
