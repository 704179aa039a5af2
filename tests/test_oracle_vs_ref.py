"""Operator-by-operator, bit-for-bit: the C restatement vs the UNMODIFIED reference compiled from
/root/reference/src/fluid.cpp (oracle/_ref, built by oracle/Makefile).  Skipped where the compiled
reference is not present (it is git-ignored; it travels with gpurun snapshots)."""
import numpy as np
import pytest

import oracle
from golden_util import assert_bit_equal

pytestmark = pytest.mark.skipif(not oracle.Reference.available(30), reason="oracle/_ref not built")

NS = [n for n in (1, 2, 3, 4, 5, 30) if oracle.Reference.available(n)]
SHAPES = [(16, 16), (29, 37), (48, 40), (8, 128), (64, 4)]


def rand_field(h, w, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((h, w, 4)) * scale).astype(np.float32)


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("dt", [0.1, 7.5, 1000.0])
def test_advect(shape, dt):
    h, w = shape
    a, b = rand_field(h, w, 1), rand_field(h, w, 2)
    a2, b2 = a.copy(), b.copy()
    oracle.Oracle(30).advect(a, b, dt)
    oracle.Reference(30).advect(a2, b2, dt)
    assert_bit_equal(b, b2, "advect out")
    assert_bit_equal(a, a2, "advect in")


@pytest.mark.parametrize("n", NS)
@pytest.mark.parametrize("shape", SHAPES[:3])
def test_diffuse(n, shape):
    h, w = shape
    a, b = rand_field(h, w, 3), rand_field(h, w, 4)
    a2, b2 = a.copy(), b.copy()
    ra, rb = oracle.Oracle(n).diffuse(a, b, 0.013, 2.5)
    qa, qb = oracle.Reference(n).diffuse(a2, b2, 0.013, 2.5)
    assert_bit_equal(ra, qa, "diffuse vp")
    assert_bit_equal(rb, qb, "diffuse vp_out")
    # pointer outcome: n-1 swaps
    assert (ra is a) == (qa is a2)


@pytest.mark.parametrize("n", NS)
@pytest.mark.parametrize("shape", SHAPES[:3])
def test_compute_pressure(n, shape):
    h, w = shape
    a, b = rand_field(h, w, 5), rand_field(h, w, 6)
    a2, b2 = a.copy(), b.copy()
    ra, rb = oracle.Oracle(n).compute_pressure(a, b, 0.37)
    qa, qb = oracle.Reference(n).compute_pressure(a2, b2, 0.37)
    assert_bit_equal(ra, qa, "pressure vp")
    assert_bit_equal(rb, qb, "pressure vp_out")
    assert (ra is a) == (qa is a2)


@pytest.mark.parametrize("shape", SHAPES)
def test_subtract_pressure_gradient(shape):
    h, w = shape
    a, b = rand_field(h, w, 7), rand_field(h, w, 8)
    a2, b2 = a.copy(), b.copy()
    oracle.Oracle(30).subtract_pressure_gradient(a, b, 0.9)
    oracle.Reference(30).subtract_pressure_gradient(a2, b2, 0.9)
    assert_bit_equal(b, b2, "subtract out")


@pytest.mark.parametrize("ishape,vshape", [((32, 48), (16, 16)), ((41, 50), (29, 37)), ((64, 96), (32, 32)),
                                            ((16, 16), (64, 64))])
@pytest.mark.parametrize("dt", [0.1, 250.0])
def test_advect_color(ishape, vshape, dt):
    rng = np.random.default_rng(9)
    img = rng.random((ishape[0], ishape[1], 4)).astype(np.float32)
    out = np.zeros_like(img)
    vp = rand_field(vshape[0], vshape[1], 10)
    img2, out2 = img.copy(), out.copy()
    oracle.Oracle(30).advect_color(img, out, vp, dt)
    oracle.Reference(30).advect_color(img2, out2, vp.copy(), dt)
    assert_bit_equal(out, out2, "advect_color out")


@pytest.mark.parametrize("n", NS)
def test_whole_steps(n):
    h, w = 40, 56
    vp, vt = rand_field(h, w, 11, 0.7), oracle.initial_vtmp(h, w)
    rng = np.random.default_rng(12)
    img = rng.random((60, 84, 4)).astype(np.float32)
    it = np.zeros_like(img)
    args2 = [x.copy() for x in (vp, vt, img, it)]
    r = oracle.Oracle(n).run_steps(vp, vt, img, it, 3.0, 0.004, 5)
    q = oracle.Reference(n).run_steps(*args2, 3.0, 0.004, 5)
    for name, x, y in zip(("vp", "vtmp", "image", "itmp"), r, q):
        assert_bit_equal(x, y, name)


def test_reference_rejects_mixed_counts():
    with pytest.raises(ValueError):
        oracle.Reference(30, 50)
