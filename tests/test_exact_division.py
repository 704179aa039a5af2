"""The 3-instruction FMA division used by the packed diffusion kernel (sweeps_packed.cu,
div_const_fast2) is correctly rounded: exhaustive worst-case enumeration (tests/exact_div_check.c)
plus random and near-midpoint quotients, all on the CPU with fmaf()."""
import os
import subprocess


HERE = os.path.dirname(os.path.abspath(__file__))


def _has_fma():
    try:
        return " fma " in open("/proc/cpuinfo").read()
    except OSError:
        return False


def test_worst_case_enumeration(tmp_path):
    exe = tmp_path / "exact_div_check"
    flags = ["-O2", "-ffp-contract=off"] + (["-mfma"] if _has_fma() else [])
    subprocess.run(["gcc", *flags, "-o", str(exe), os.path.join(HERE, "exact_div_check.c"), "-lm"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True, timeout=600).stdout
    # "worst-case quotients tested N  one-iteration mismatches A  two-iteration mismatches B"
    words = out.strip().split()
    tested, one_iter_bad = int(words[3]), int(words[6])
    assert tested > 20_000_000
    assert one_iter_bad == 0, out
    last = out.strip().splitlines()[-1].split()     # "random quotients tested N  mismatches M"
    assert int(last[3]) > 10_000_000 and int(last[5]) == 0, out




# Divisors beta for which the two-instruction division is wrong for one numerator (found with `div2_check 600 1`), and
# a few for which it is right for all of them (the benchmark's beta among them).
DIV2_WRONG = ["3ff4291f", "406b65c7", "40d1684b", "40d8a5eb", "40f8d583", "40e68797"]
DIV2_RIGHT = ["3f800d1b", "3f800000", "40000000", "3fb33333", "3f8a3d70"]


def _beta_to_params(bits_hex):
    """viscosity, dt with (float)(1.0 + 4.0*(double)(viscosity*dt)) == beta exactly (fluid.cpp:144-145)."""
    import numpy as np
    beta = np.array([int(bits_hex, 16)], dtype=np.uint32).view(np.float32)[0]
    visc = np.float32((np.float32(beta) - np.float32(1.0)) / np.float32(4.0))      # exact: Sterbenz, then a power of two
    assert np.float32(1.0 + 4.0 * float(visc)) == beta
    return float(visc), 1.0


def test_two_instruction_division_is_used_only_where_it_is_exact(tmp_path):
    """The library decides per divisor (div2_constants in sweeps_packed.cu) whether RN(a*zh + RN(a*zl)) may replace
    the 3-instruction division; tests/div2_check.c takes the same decision independently.  They must agree, on
    divisors where the short form is wrong for some numerator and on divisors where it never is."""
    import numpy as np
    from probabilistic_fluid_simulation_b200 import _cabi
    exe = tmp_path / "div2_check"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", str(exe), os.path.join(HERE, "div2_check.c"), "-lm"], check=True)
    headline = np.float32(1.0 + 4.0 * float(np.float32(0.001) * np.float32(0.1)))
    assert f"{headline.view(np.uint32):08x}" == DIV2_RIGHT[0]
    out = subprocess.run([str(exe), "bits", *DIV2_WRONG, *DIV2_RIGHT], check=True, capture_output=True, text=True,
                         timeout=600).stdout
    verdict = {ln.split()[0]: ln.split()[1] == "1" for ln in out.strip().splitlines()}
    assert all(not verdict[b] for b in DIV2_WRONG) and all(verdict[b] for b in DIV2_RIGHT), out
    L = _cabi.lib()
    for b, ok in verdict.items():
        visc, dt = _beta_to_params(b)
        assert L.pfs_diffuse_division_ops(visc, dt) == (2 if ok else 3), b
    # random divisors: same decision as the independent checker
    out = subprocess.run([str(exe), "24", "5"], check=True, capture_output=True, text=True, timeout=600).stdout
    for ln in out.strip().splitlines():
        bits, ok = ln.split()[0], ln.split()[1] == "1"
        visc, dt = _beta_to_params(bits)
        assert L.pfs_diffuse_division_ops(visc, dt) == (2 if ok else 3), ln
    assert L.pfs_diffuse_division_ops(-0.5, 1.0) == 0          # negative alpha: scalar kernels, exact IEEE division


def test_fast_periodic_wrap_equals_double_fmod(tmp_path):
    """wrap_coord() in pfs_internal.cuh skips fmodf where the result is known in closed form (Sterbenz);
    tests/wrap_check.c compares it with fmod(fmod(x, W) + W, W) on 52 M values incl. every edge case."""
    exe = tmp_path / "wrap_check"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", str(exe), os.path.join(HERE, "wrap_check.c"), "-lm"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True, timeout=600).stdout
    words = out.strip().splitlines()[-1].split()
    assert int(words[3]) > 50_000_000 and int(words[5]) == 0, out
