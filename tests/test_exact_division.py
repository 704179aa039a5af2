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




def test_fast_periodic_wrap_equals_double_fmod(tmp_path):
    """wrap_coord() in pfs_internal.cuh skips fmodf where the result is known in closed form (Sterbenz);
    tests/wrap_check.c compares it with fmod(fmod(x, W) + W, W) on 52 M values incl. every edge case."""
    exe = tmp_path / "wrap_check"
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", str(exe), os.path.join(HERE, "wrap_check.c"), "-lm"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True, timeout=600).stdout
    words = out.strip().splitlines()[-1].split()
    assert int(words[3]) > 50_000_000 and int(words[5]) == 0, out
