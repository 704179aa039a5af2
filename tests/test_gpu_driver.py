"""The main.cpp-compatible driver (host/main.cpp -> build/fluidsim_b200): same CLI, stdout lines and
output frames as the reference driver, checked against frames computed by the CPU oracle."""
import os
import subprocess

import numpy as np
import pytest

import oracle
from golden_util import GOLD
from probabilistic_fluid_simulation_b200 import fixtures, pngio

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "probabilistic_fluid_simulation_b200", "host", "build", "fluidsim_b200")


@pytest.fixture(scope="module")
def pngs(tmp_path_factory):
    d = tmp_path_factory.mktemp("pngs")
    vel = np.load(os.path.join(GOLD, "png_perlin_t0_64.npz"))["rgba"]
    img = np.load(os.path.join(GOLD, "png_baboon.npz"))["rgba"][:96, :160].copy()
    pngio.write_rgba8(str(d / "vel.png"), vel)
    pngio.write_rgba8(str(d / "img.png"), img)
    # 8-bit RGBA files decode to the bytes that were written
    assert np.array_equal(pngio.read_rgba8(str(d / "vel.png")), vel)
    return d, vel, img


def test_driver_frames_match_oracle(pngs, tmp_path):
    d, vel, img = pngs
    assert os.path.exists(EXE), "driver not built (run __graft_entry__.build())"
    out = tmp_path / "frames"
    out.mkdir()
    steps, dt, visc = 4, 0.5, 0.001
    r = subprocess.run([EXE, str(steps), str(dt), str(visc), str(d / "img.png"), str(d / "vel.png"), str(out)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert lines[0] == f"Simulating [64 x 64] domain for {steps} timesteps at dt={dt}..."      # main.cpp:214-215
    assert lines[-1].startswith(f"{steps} timesteps took ") and lines[-1].endswith(" us.")       # main.cpp:250
    vp, vtmp, image, itmp = fixtures.make_state(vel, img)
    orc = oracle.Oracle(30)
    for i in range(steps):
        vp, vtmp, image, itmp = orc.run_steps(vp, vtmp, image, itmp, np.float32(dt), np.float32(visc), 1)
        assert lines[1 + i] == f"[{i}] Writing to : {out}/{i}.png"                                # main.cpp:67
        frame = pngio.read_rgba8(str(out / f"{i}.png"))
        assert np.array_equal(frame, fixtures.unit_float_to_bytes(image)), f"frame {i}"          # utils.hpp:129-131


def test_driver_timing_mode_and_errors(pngs):
    d, _, _ = pngs
    r = subprocess.run([EXE, "2", "0.1", "0.001", str(d / "img.png"), str(d / "vel.png")], capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0 and "Writing to" not in r.stdout and "2 timesteps took" in r.stdout
    assert subprocess.run([EXE], capture_output=True).returncode == 1
    r = subprocess.run([EXE, "0", "0.1", "0.001", "a", "b"], capture_output=True, text=True)
    assert r.returncode == 1 and "Timesteps must be greater than 0." in r.stderr
    r = subprocess.run([EXE, "3", "0.1", "0.001", "/nonexistent.png", "b"], capture_output=True, text=True)
    assert r.returncode == 1 and "Something went wrong reading the input image..." in r.stderr


# ---- the committed frames of the reference's own program (tests/golden/driver_frames.json) --------------------------
import driver_cases as dc  # noqa: E402


@pytest.mark.parametrize("name", sorted(dc.CASES))
def test_b200_driver_matches_reference_program(name, tmp_path):
    lines, crcs, out = dc.run_driver(dc.B200, name, str(tmp_path))
    gold = dc.load_driver_golden()["cases"][name]
    assert crcs == gold["frames_crc32"]
    assert lines[0] == gold["stdout_head"]
    want = dc.expected_lines(name, out, pngio.read_rgba8(str(tmp_path / "vel.png")).shape)
    assert lines[:len(want)] == want


@pytest.mark.skipif(not os.path.exists(dc.REF_MAIN_B200), reason="oracle/_ref/fluidsim_cuda_b200 not built")
@pytest.mark.parametrize("name", sorted(dc.CASES))
def test_unmodified_reference_main_on_b200_backend(name, tmp_path):
    """src/main.cpp -DUSE_CUDA of the reference, unmodified, linked against libfluid_b200.so instead of fluid.cu
    (INTEGRATION.md 3.1).  PFS_SHIM_INIT_TMP=1 supplies the initial vtmp the reference's CUDA driver forgets to upload."""
    lines, crcs, out = dc.run_driver(dc.REF_MAIN_B200, name, str(tmp_path), env={"PFS_SHIM_INIT_TMP": "1"})
    gold = dc.load_driver_golden()["cases"][name]
    assert crcs == gold["frames_crc32"]
    assert lines[0] == gold["stdout_head"]


@pytest.mark.parametrize("writers", ["0", "1", "3"])
def test_b200_driver_frame_writer_modes(writers, tmp_path):
    """Serial copy-and-encode (0) and the threaded frame ring (host/frame_writer.hpp) write the same files."""
    name = "voronoi256_tulips"
    lines, crcs, out = dc.run_driver(dc.B200, name, str(tmp_path), env={"PFS_FRAME_WRITERS": writers})
    assert crcs == dc.load_driver_golden()["cases"][name]["frames_crc32"]


@pytest.mark.parametrize("writers", ["0", "2"])
def test_b200_driver_unwritable_output_dir_is_not_fatal(writers, pngs, tmp_path):
    """The reference ignores the return value of write_png_from_array (main.cpp:68): a run whose frames cannot be written
    still prints every "Writing to" line and the timing line and exits 0.  Both writer modes of the B200 driver do the
    same, plus one warning per frame on stderr."""
    d, _, _ = pngs
    env = dict(os.environ)
    env["PFS_FRAME_WRITERS"] = writers
    missing = tmp_path / "does" / "not" / "exist"
    r = subprocess.run([EXE, "3", "0.5", "0.001", str(d / "img.png"), str(d / "vel.png"), str(missing)],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, (r.stdout, r.stderr)
    lines = r.stdout.strip().splitlines()
    assert [ln for ln in lines if "Writing to" in ln] == [f"[{i}] Writing to : {missing}/{i}.png" for i in range(3)]
    assert lines[-1].startswith("3 timesteps took ") and lines[-1].endswith(" us.")
    assert r.stderr.count("warning: cannot write") == 3, r.stderr


def test_b200_driver_on_the_stateless_entry_points(tmp_path):
    """PFS_DRIVER_STATELESS=1: the driver loop calls the fluid.hpp entry points on caller-owned interleaved buffers, as
    the reference's loop does (main.cpp:222,225), instead of a persistent context; same frames."""
    name = "voronoi256_tulips"
    lines, crcs, out = dc.run_driver(dc.B200, name, str(tmp_path), env={"PFS_DRIVER_STATELESS": "1"})
    assert crcs == dc.load_driver_golden()["cases"][name]["frames_crc32"]
