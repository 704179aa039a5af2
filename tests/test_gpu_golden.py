"""GPU parity, part 1: the CUDA path (through the C-ABI) against the golden hashes produced by the
unmodified reference.  Bit-exact: FP32 with no FMA contraction on either side (tolerance = 0)."""
import numpy as np
import pytest

import oracle
import probabilistic_fluid_simulation_b200 as pfs
from golden_util import case_state, load_golden
from gpu_util import gpu_run_steps

pytestmark = pytest.mark.gpu

CASES = {c["name"]: c for c in load_golden()["cases"]}


@pytest.mark.parametrize("name", list(CASES))
def test_device_api_matches_reference_hashes(name):
    c = CASES[name]
    vp, vtmp, image, itmp = case_state(c)
    n = c["n_iters"]
    before = pfs.kernel_launch_count()
    vp, vtmp, image, itmp = gpu_run_steps(vp, vtmp, image, itmp, c["dt"], c["viscosity"], n, n, c["steps"])
    assert pfs.kernel_launch_count() > before          # the CUDA kernels really ran
    assert oracle.field_hashes(vp) == c["hashes"]["vp"]
    assert oracle.field_hashes(vtmp) == c["hashes"]["vtmp"]
    if image is not None:
        assert oracle.field_hashes(image) == c["hashes"]["image"]


@pytest.mark.parametrize("name", ["g2_formula", "n3", "odd_shape_n30", "g1_10steps"])
def test_host_api_matches_reference_hashes(name):
    """The reference's CPU-build signatures (vp_field structs with host pointers)."""
    c = CASES[name]
    vp, vtmp, image, itmp = case_state(c)
    fv, ft, fi, fm = (pfs.vp_field(x) for x in (vp, vtmp, image, itmp))
    n = c["n_iters"]
    for _ in range(c["steps"]):
        pfs.simulate_fluid_step(fv, ft, c["dt"], c["viscosity"], n, n)
        pfs.advect_color_step(fi, fm, fv, c["dt"])
    assert oracle.field_hashes(fv.data) == c["hashes"]["vp"]
    assert oracle.field_hashes(ft.data) == c["hashes"]["vtmp"]
    assert oracle.field_hashes(fi.data) == c["hashes"]["image"]


@pytest.mark.parametrize("name", ["g2_formula", "n4", "tulips_voronoi_5"])
def test_timestep_host_matches_reference_hashes(name):
    c = CASES[name]
    state = case_state(c)
    pinned = []
    for a in state:
        p = pfs.pinned_empty(a.shape)
        p[...] = a
        pinned.append(p)
    fv, ft, fi, fm = (pfs.vp_field(x) for x in pinned)
    n = c["n_iters"]
    for _ in range(c["steps"]):
        pfs.timestep_host(fv, ft, fi, fm, c["dt"], c["viscosity"], n, n)
    assert oracle.field_hashes(fv.data) == c["hashes"]["vp"]
    assert oracle.field_hashes(ft.data) == c["hashes"]["vtmp"]
    assert oracle.field_hashes(fi.data) == c["hashes"]["image"]
    for p in pinned:
        pfs.pinned_free(p)
