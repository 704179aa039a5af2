"""The C restatement (oracle/fluid_oracle.c) against the golden vectors produced by the UNMODIFIED
reference (tests/golden/golden.json, scripts/make_golden.py) -- this is what pins the oracle."""
import numpy as np
import pytest

import oracle
from golden_util import case_state, load_golden

GOLDEN = load_golden()
CASES = {c["name"]: c for c in GOLDEN["cases"]}

# SURVEY.md 4.4 hashes, produced independently during the survey (G1 1 and 10 steps, G2).
SURVEY_PINS = {
    "g1_1step": {"vp": ["3281edb618eafad4", "5afdadb36e774292", "16733404d77448cb", "aeba972a6e81e41f"],
                 "vtmp": ["334c27a533313430", "837fa56ac0baeb65", "1bfdfa1c1832d434", "aeba972a6e81e41f"],
                 "image": ["2e602272ebc40475", "efd921d83213087f", "4461a64500183e0a", "0c500398b8cd0383"]},
    "g1_10steps": {"vp": ["fd9b072d16fe29c5", "05918d0d818c780f", "2d7b1ca82dbca390", "449988e5881c5338"],
                   "vtmp": ["066d5388b1b4b202", "5a5368964a916e11", "4ddbcfb87f9d74cd", "449988e5881c5338"],
                   "image": ["98c397005da02687", "76b60ff51abfa427", "4839a98d45fe5233", "0c500398b8cd0383"]},
    "g2_formula": {"vp": ["df6086e65c19164c", "ba5eeaaa3faa2123", "66edc2a5d2dca914", "1a85de9a45e6bc83"],
                   "vtmp": ["daf544d3f06333b5", "233b619e46be5b9c", "cd80d93a4ed4eb52", "1a85de9a45e6bc83"],
                   "image": ["303811c4c0a9718e", "5d8fc331c1bf78fa", "d72f8c59322742dc", "d51556cd26bceacf"]},
}


def test_golden_file_agrees_with_survey_pins():
    for name, pins in SURVEY_PINS.items():
        assert CASES[name]["hashes"] == pins, name


# "heavy" cases (>12 s of single-core CPU time when generated) are checked against the reference hashes by
# the GPU suite only; the CPU suite keeps to the light ones so that it finishes within a few minutes.
@pytest.mark.parametrize("name", [n for n, c in CASES.items() if not c.get("heavy")])
def test_oracle_reproduces_reference_hashes(name):
    c = CASES[name]
    vp, vtmp, image, itmp = case_state(c)
    orc = oracle.Oracle(c["n_iters"])
    vp, vtmp, image, itmp = orc.run_steps(vp, vtmp, image, itmp, c["dt"], c["viscosity"], c["steps"])
    assert oracle.field_hashes(vp) == c["hashes"]["vp"]
    assert oracle.field_hashes(vtmp) == c["hashes"]["vtmp"]
    if image is not None:
        assert oracle.field_hashes(image) == c["hashes"]["image"]
    assert np.isclose(float(vp[..., 0].astype(np.float64).sum()), c["sums"]["vp_u_sum"], rtol=0, atol=1e-6)


def test_jacobi_preserves_pressure_mean_g1():
    # after 10 steps of G1 the pressure sum stays -1 * 256^2 (mean-preserving sweeps)
    c = CASES["g1_10steps"]
    assert abs(c["sums"]["vtmp_p_sum"] + 256 * 256) < 1.0


def test_png_fixture_crcs_match_survey():
    import zlib
    import os
    from golden_util import GOLD
    want = {"png_baboon": "694c777d", "png_tulips": "a114f8ff", "png_perlin_t0_256": "a31d1ab6",
            "png_perlin_t0_64": "297b9608", "png_voronoi_256": "26ac43e3", "png_circular_128": "a13cdee0",
            "png_solid_r64": "603429f6"}
    for stem, crc in want.items():
        rgba = np.load(os.path.join(GOLD, stem + ".npz"))["rgba"]
        assert f"{zlib.crc32(rgba.tobytes()) & 0xffffffff:08x}" == crc, stem
        assert GOLDEN["png_inputs"][stem]["crc32"] == crc
