"""Argument handling of the main.cpp-compatible driver (host/main.cpp -> host/build/fluidsim_b200) needs no GPU:
every rejected command line must end like the reference's own program (src/main.cpp:105-131, compiled unmodified
into oracle/_ref/fluidsim_cpu where the reference tree exists) -- same exit code, same stdout, same stderr up to
argv[0]."""
import os
import subprocess

import pytest

import driver_cases as dc

USAGE = "Usage: {exe} <n_timesteps> <delta_t> <viscosity> <input_image> <velocity_field> [output_dir]\n"   # main.cpp usage()

# (arguments, stderr the reference prints -- main.cpp:105-131)
CASES = {
    "no_args": ([], USAGE),
    "one_arg": (["5"], USAGE),
    "four_args": (["5", "0.1", "0.001", "a.png"], USAGE),
    "seven_args": (["1", "2", "3", "4", "5", "6", "7"], USAGE),
    "zero_steps": (["0", "0.1", "0.001", "a.png", "b.png"], "Timesteps must be greater than 0.\n"),
    "negative_steps": (["-3", "0.1", "0.001", "a.png", "b.png"], "Timesteps must be greater than 0.\n"),
    "steps_not_a_number": (["abc", "0.1", "0.001", "a.png", "b.png"], "Timesteps must be greater than 0.\n"),
    "zero_dt": (["3", "0", "0.001", "a.png", "b.png"], "Delta T must be greater than 0.\n"),
    "negative_dt": (["3", "-1", "0.001", "a.png", "b.png", "out"], "Delta T must be greater than 0.\n"),
    "dt_not_a_number": (["3", "x", "0.001", "a.png", "b.png"], "Delta T must be greater than 0.\n"),
    "missing_image": (["3", "0.1", "0.001", "/nonexistent/a.png", "/nonexistent/b.png"],
                      "Something went wrong reading the input image...\n"),
    "image_not_a_png": (["3", "0.1", "0.001", "{notpng}", "{notpng}"],
                        "Something went wrong reading the input image...\n"),
}


def _run(exe, args):
    r = subprocess.run([exe] + args, capture_output=True, text=True, timeout=120)
    return r.returncode, r.stdout, r.stderr.replace(exe, "{exe}")


@pytest.fixture(scope="module")
def notpng(tmp_path_factory):
    p = tmp_path_factory.mktemp("cli") / "notpng.png"
    p.write_text("this is not a PNG file\n")
    return str(p)


@pytest.mark.skipif(not os.path.exists(dc.B200), reason="host/build/fluidsim_b200 not built")
@pytest.mark.parametrize("name", sorted(CASES))
def test_rejected_command_lines(name, notpng):
    args, want_err = CASES[name]
    args = [a.format(notpng=notpng) for a in args]
    rc, out, err = _run(dc.B200, args)
    assert rc == 1 and out == "" and err == want_err, (rc, out, err)
    if os.path.exists(dc.REF_CPU):                      # the reference's own program says the same
        assert _run(dc.REF_CPU, args) == (rc, out, err)
