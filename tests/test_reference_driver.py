"""The reference's OWN program, unmodified (oracle/_ref/fluidsim_cpu: src/main.cpp + src/fluid.cpp compiled where they
lie, oracle/Makefile), pins the end-to-end expectations of the driver tests: its frames must equal the frames the
CPU oracle computes and the CRCs committed in tests/golden/driver_frames.json (which the GPU tests compare with)."""
import os

import numpy as np
import pytest

import driver_cases as dc
import oracle
from probabilistic_fluid_simulation_b200 import fixtures, pngio


def test_driver_golden_covers_every_case():
    gold = dc.load_driver_golden()["cases"]
    assert set(gold) == set(dc.CASES)
    for name, g in gold.items():
        assert len(g["frames_crc32"]) == dc.CASES[name][4]


@pytest.mark.parametrize("name", sorted(dc.CASES))
def test_oracle_frames_match_committed_crcs(name, tmp_path):
    """The oracle alone (no reference tree needed) reproduces the committed frame CRCs of the reference program."""
    vel, img, _, _, steps, dt, visc = dc.write_inputs(name, str(tmp_path))
    gold = dc.load_driver_golden()["cases"][name]
    vp, vtmp, image, itmp = fixtures.make_state(vel, img)
    orc = oracle.Oracle(30)
    for i in range(steps):
        # main.cpp:115-116: delta_t and viscosity are atof() results narrowed to float
        vp, vtmp, image, itmp = orc.run_steps(vp, vtmp, image, itmp, np.float32(float(dt)), np.float32(float(visc)), 1)
        assert pngio.crc32(fixtures.unit_float_to_bytes(image)) == gold["frames_crc32"][i], f"frame {i}"


@pytest.mark.skipif(not os.path.exists(dc.REF_CPU), reason="oracle/_ref/fluidsim_cpu not built (no reference tree)")
@pytest.mark.parametrize("name", sorted(dc.CASES))
def test_reference_program_matches_golden(name, tmp_path):
    lines, crcs, out = dc.run_driver(dc.REF_CPU, name, str(tmp_path))
    gold = dc.load_driver_golden()["cases"][name]
    assert crcs == gold["frames_crc32"]
    vel = pngio.read_rgba8(str(tmp_path / "vel.png"))
    want = dc.expected_lines(name, out, vel.shape)
    assert lines[:len(want)] == want
    assert lines[-1].startswith(f"{dc.CASES[name][4]} timesteps took ") and lines[-1].endswith(" us.")
