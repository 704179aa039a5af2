"""The C-ABI library loads and exports every symbol include/pfs_b200.h declares; the Python binding
lists exactly those symbols; with no GPU the product path fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import probabilistic_fluid_simulation_b200 as pfs
from probabilistic_fluid_simulation_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pfs_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pfs_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_reference_entry_points():
    names = declared_functions()
    for need in ("pfs_simulate_fluid_step", "pfs_advect_color_step", "pfs_advect", "pfs_advect_color", "pfs_diffuse",
                 "pfs_add_forces", "pfs_compute_pressure", "pfs_subtract_pressure_gradient",
                 "pfs_simulate_fluid_step_host", "pfs_advect_color_step_host"):
        assert need in names


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_cabi.LIB_PATH), "libpfs_b200.so not built (run __graft_entry__.build())"
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in pfs_b200.h but not exported"


def test_binding_covers_exactly_the_header():
    assert sorted(_cabi.SIGNATURES) == declared_functions()


def test_version_and_knobs_without_gpu():
    L = _cabi.lib()
    assert L.pfs_version() == 200
    assert L.pfs_set_fuse_depth(-1) == 1          # PFS_EINVAL
    assert b"depth" in L.pfs_last_error()
    assert L.pfs_set_fuse_depth(0) == 0
    assert L.pfs_kernel_launch_count() >= 0


def test_struct_layout_matches_vp_field():
    # includes/fluid.hpp:17-22: three ints then a pointer
    assert ctypes.sizeof(_cabi.Field) == 24
    assert _cabi.Field.data.offset == 16


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_without_gpu():
    a = np.zeros((8, 8, 4), np.float32)
    b = np.zeros((8, 8, 4), np.float32)
    with pytest.raises(pfs.PfsError) as e:
        pfs.simulate_fluid_step(pfs.vp_field(a), pfs.vp_field(b), 0.1, 0.001)
    assert e.value.status == 4   # PFS_ENODEVICE


def test_argument_validation():
    L = _cabi.lib()
    null = ctypes.c_void_p(0)
    assert L.pfs_advect(null, null, 0.1, 8, 8, 4, None) == 1
    assert L.pfs_advect(ctypes.c_void_p(16), ctypes.c_void_p(32), 0.1, 8, 8, 3, None) == 1      # channels != 4
    assert L.pfs_advect(ctypes.c_void_p(16), ctypes.c_void_p(36), 0.1, 8, 8, 4, None) == 1      # misaligned
    assert L.pfs_advect(ctypes.c_void_p(16), ctypes.c_void_p(32), 0.1, 0, 8, 4, None) == 1      # empty grid
    a, b = ctypes.c_void_p(16), ctypes.c_void_p(32)
    assert L.pfs_simulate_fluid_step(ctypes.byref(a), ctypes.byref(b), 0.1, 0.0, 8, 8, 4, 0, 30, None) == 1
    assert L.pfs_simulate_fluid_step(ctypes.byref(a), ctypes.byref(a), 0.1, 0.0, 8, 8, 4, 30, 30, None) == 1
    assert L.pfs_simulate_fluid_step(ctypes.byref(a), ctypes.byref(b), 0.1, 0.0, 20000, 20000, 4, 30, 30, None) == 1
